"""One fc1-shaped GEMM (GELU epilogue) for an ncu --set full capture."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L
lib = L.lib()
act = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator(device="cuda").manual_seed(1)
M, K, N = 16384, 1280, 5120
A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
shift = torch.randn(N, device="cuda", generator=g) * 0.1
out = torch.empty(M, N, device="cuda", dtype=torch.half)
epi = L.TcEpilogue(kind=L.EPI_F16, act=act, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
for _ in range(3):
    L.check(lib.cvb_op_gemm_f16(L.ptr(A), M, K, C.c_longlong(K), L.ptr(W), N, C.c_longlong(K), 256, C.byref(epi), L.stream_ptr()), "g")
torch.cuda.synchronize()
