"""Probe: patch-resident conv (descriptor base-offset semantics) vs fp32 reference, timing vs the k-block conv."""
import ctypes as C, sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(5)
lib = L.lib()
def run(mode, s0, s1, Wp, N, scale, shift, B, H, W, C0, C1, reps=1):
    lib.cvb_tc_set_conv_patch_mode(mode)
    out = torch.full((B, H, W, N), float("nan"), device="cuda", dtype=torch.half)
    epi = L.TcEpilogue(kind=L.EPI_F16, act=L.ACT_RELU, scale=scale.data_ptr(), shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        L.check(lib.cvb_op_conv3x3_f16(L.ptr(s0), C0, L.ptr(s1), C1, B, H, W, L.ptr(Wp), N, N, C.byref(epi), L.stream_ptr()), "conv")
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps
for (B, H, W, C0, C1, N) in [(1, 4, 128, 64, 0, 64), (2, 64, 256, 64, 64, 64), (2, 128, 128, 128, 128, 128), (4, 1024, 1024, 64, 0, 64), (4, 1024, 1024, 64, 64, 64), (4, 512, 512, 128, 128, 128), (4, 512, 512, 128, 0, 128), (4, 256, 256, 256, 256, 256)]:
    s0 = (torch.randn(B, H, W, C0, device="cuda", generator=g) * 0.5).half()
    s1 = (torch.randn(B, H, W, C1, device="cuda", generator=g) * 0.5).half() if C1 else None
    w = torch.randn(N, C0 + C1, 3, 3, device="cuda", generator=g) * 0.03
    Wp = w.permute(0, 2, 3, 1).reshape(N, -1).contiguous().half()
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    big = H * W * B > 2e6
    o0, t0 = run(0, s0, s1, Wp, N, scale, shift, B, H, W, C0, C1, 3 if big else 1)
    res = [f"{B}x{H}x{W} C={C0}+{C1} N={N}: kblock {t0*1e3:.0f}us"]
    for mode in (1,):
        o, t = run(mode, s0, s1, Wp, N, scale, shift, B, H, W, C0, C1, 3 if big else 1)
        err = (o.float() - o0.float()).abs().max().item()
        res.append(f"mode{mode}: err_vs_kblock {err:.3e} nan {int(torch.isnan(o.float()).sum())} {t*1e3:.0f}us")
    if not big:
        xin = torch.cat([s0, s1], -1) if C1 else s0
        ref = F.relu(F.conv2d(xin.permute(0, 3, 1, 2).float(), Wp.view(N, 3, 3, C0 + C1).permute(0, 3, 1, 2).float(), padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
        res.append(f"kblock_vs_fp32 {(o0.permute(0,3,1,2).float()-ref).abs().max().item():.3e}")
    print(" | ".join(res), flush=True)
lib.cvb_tc_set_conv_patch_mode(1)
