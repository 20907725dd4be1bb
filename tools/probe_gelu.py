"""Does the GELU epilogue cost time because of issue slots on the SM or because of chip power? Same GEMM, act none/relu/gelu,
on the full and on a restricted persistent grid (a quarter of the SMs never hits the power cap)."""
import ctypes as C, sys, subprocess, torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L
lib = L.lib()
g = torch.Generator(device="cuda").manual_seed(1)
M, K, N = 16384, 1280, 5120
A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
shift = torch.randn(N, device="cuda", generator=g) * 0.1
out = torch.empty(M, N, device="cuda", dtype=torch.half)
def bench(act, reps=30):
    epi = L.TcEpilogue(kind=L.EPI_F16, act=act, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
    def run():
        L.check(lib.cvb_op_gemm_f16(L.ptr(A), M, K, C.c_longlong(K), L.ptr(W), N, C.c_longlong(K), 256, C.byref(epi), L.stream_ptr()), "g")
    for _ in range(5): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    return e0.elapsed_time(e1) / reps * 1e3, clk
for ctas in (148, 36):
    lib.cvb_tc_set_max_ctas(ctas)
    for name, act in (("none", 0), ("relu", 1), ("gelu", 2)):
        us, clk = bench(act, 200 if ctas == 148 else 60)
        print(f"ctas={ctas:3d} act={name} {us:8.1f} us  {2.0*M*N*K/us/1e6:7.0f} TFLOP/s   [{clk}]", flush=True)
lib.cvb_tc_set_max_ctas(0)
