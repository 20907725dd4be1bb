"""Is the tile engine bound by chip-wide L2->SM feed or by a per-SM cost? Same GEMM on a restricted persistent grid."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L
lib = L.lib()
g = torch.Generator(device="cuda").manual_seed(1)
def bench(M, K, N, bn, reps=10):
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    out = torch.empty(M, N, device="cuda", dtype=torch.half)
    epi = L.TcEpilogue(kind=L.EPI_F16, act=0, out=out.data_ptr(), ldc=N)
    def run():
        L.check(lib.cvb_op_gemm_f16(L.ptr(A), M, K, C.c_longlong(K), L.ptr(W), N, C.c_longlong(K), bn, C.byref(epi), L.stream_ptr()), "g")
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 2.0 * M * N * K / us / 1e6
for (M, K, N) in [(16384, 1280, 5120), (16384, 5120, 5120)]:
    for pair in (1, 0):
        lib.cvb_tc_set_pair_mode(pair)
        for ctas in (148, 112, 74, 36, 16):
            lib.cvb_tc_set_max_ctas(ctas)
            us, tf = bench(M, K, N, 256)
            print(f"M{M} K{K} N{N} pair={pair} ctas={ctas:3d} {us:8.1f} us {tf:7.0f} TFLOP/s  per-CTA {tf/ctas:6.2f}", flush=True)
lib.cvb_tc_set_max_ctas(0); lib.cvb_tc_set_pair_mode(1)
