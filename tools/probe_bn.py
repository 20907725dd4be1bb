"""Tile-engine block_n sweep on the encoder / ConvT GEMM shapes (B=4 tiles): wave quantisation vs tile width."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L
lib = L.lib()
g = torch.Generator(device="cuda").manual_seed(1)
def bench(M, K, N, kind, act, bn, reps=20):
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    if kind == L.EPI_F16:
        out = torch.empty(M, N, device="cuda", dtype=torch.half)
        epi = L.TcEpilogue(kind=kind, act=act, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
    else:
        out = torch.zeros(M, N, device="cuda")
        epi = L.TcEpilogue(kind=kind, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N, res=out.data_ptr(), ldres=N)
    def run():
        L.check(lib.cvb_op_gemm_f16(L.ptr(A), M, K, C.c_longlong(K), L.ptr(W), N, C.c_longlong(K), bn, C.byref(epi), L.stream_ptr()), "g")
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 2.0 * M * N * K / us / 1e6
shapes = [("qkv", 19600, 1280, 3840, L.EPI_F16, 0), ("qkv-glob", 16384, 1280, 3840, L.EPI_F16, 0), ("proj f16", 19600, 1280, 1280, L.EPI_F16, 0),
          ("fc1 gelu", 16384, 1280, 5120, L.EPI_F16, 2), ("fc1 none", 16384, 1280, 5120, L.EPI_F16, 0), ("fc2 res", 16384, 5120, 1280, L.EPI_RES_F32, 0)]
for name, M, K, N, kind, act in shapes:
    for pair in (1, 0):
        lib.cvb_tc_set_pair_mode(pair)
        for bn in (64, 96, 128, 160, 192, 224, 256):
            if N % bn: continue
            if pair and bn % 32: continue
            try:
                us, tf = bench(M, K, N, kind, act, bn)
                print(f"{name:10s} pair={pair} bn={bn:3d} {us:8.1f} us {tf:7.0f} TFLOP/s", flush=True)
            except Exception as ex:
                print(name, pair, bn, "ERR", str(ex)[:80])
lib.cvb_tc_set_pair_mode(1)
