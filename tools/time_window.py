"""Times one window-attention implementation at the BASELINE shape (SAM-H, B = 4: 100 windows x 16 heads, 196 tokens,
head dim 80) with CUDA events, and checks it against the torch expression of image_encoder.py:235-260,354-392.
usage: python tools/time_window.py <variant>     0 = mma.sync window kernel, 1 = tcgen05 kernel (window_tc.cu), 2 = the same with dynamic item scheduling; optional third argument `zero` = zero rel-pos tables"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cellvit_b200 import _lib as L  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_items = int(sys.argv[2]) if len(sys.argv) > 2 else 100
heads, hd, S, gh, gw = 16, 80, 196, 14, 14
D = heads * hd
g = torch.Generator(device="cuda").manual_seed(13)
qkv = torch.randn(n_items * S, 3 * D, device="cuda", generator=g).half()
Rh = (torch.randn(2 * gh - 1, hd, device="cuda", generator=g) * 0.2).half()
Rw = (torch.randn(2 * gw - 1, hd, device="cuda", generator=g) * 0.2).half()
relcat = torch.zeros(64, hd, device="cuda", dtype=torch.half)
relcat[:27] = Rh
relcat[32:59] = Rw
out = torch.full((n_items * S, D), float("nan"), device="cuda", dtype=torch.half)
scale = hd ** -0.5
lib = L.lib()
import os
if os.environ.get("CVB_WSKEW") is not None:
    lib.cvb_debug_window_skew(int(os.environ["CVB_WSKEW"]))
ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
if len(sys.argv) > 3 and sys.argv[3] == "zero":
    relcat.zero_(); Rh.zero_(); Rw.zero_()


def run():
    if variant == 0:
        L.check(lib.cvb_op_attention(L.ptr(qkv), n_items, S, heads, hd, C.c_float(scale), L.ptr(Rh), L.ptr(Rw), gh, gw, L.ptr(out),
                                     L.stream_ptr()), "attention")
    else:
        if variant == 2:   # dynamic item scheduling
            ctr.zero_()
        L.check(lib.cvb_op_window_attention_tc(L.ptr(qkv), n_items, heads, hd, C.c_float(scale), L.ptr(relcat), L.ptr(out),
                                               L.ptr(ctr) if variant == 2 else None, L.stream_ptr()), "window_tc")


run()
torch.cuda.synchronize()
# reference on a subset of items (fp32 torch)
chk = min(n_items, 8)
q, k, v = qkv[:chk * S].float().view(chk, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
att = (q * scale) @ k.transpose(-1, -2)
idx = torch.arange(gh, device="cuda")[:, None] - torch.arange(gh, device="cuda")[None, :] + gh - 1
rq = q.reshape(chk, heads, gh, gw, hd)
rel_h = torch.einsum("bnhwc,hkc->bnhwk", rq, Rh.float()[idx])
rel_w = torch.einsum("bnhwc,wkc->bnhwk", rq, Rw.float()[idx])
att = (att.view(chk, heads, gh, gw, gh, gw) + rel_h[..., :, None] + rel_w[..., None, :]).view(chk, heads, S, S)
ref = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(chk * S, D)
err = (out[:chk * S].float() - ref).abs().max().item()
finite = bool(torch.isfinite(out.float()).all())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(30):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts = sorted(ts[5:])
print(f"variant {variant} items {n_items}: max err {err:.2e} finite {finite} median {ts[len(ts) // 2]:.1f} us min {ts[0]:.1f} us")
if len(sys.argv) > 4:
    d = (out[:chk * S].float() - ref).abs()
    i = int(d.argmax()); r, c = divmod(i, D)
    print("mean err %.3e; worst at row %d (q %d of item %d) col %d (head %d dim %d): got %.5f want %.5f" % (
        d.mean().item(), r, r % S, r // S, c, c // hd, c % hd, out[r, c].item(), ref[r, c].item()))
    print("err by query group: rows<128 %.3e rows>=128 %.3e" % (d.view(chk, S, D)[:, :128].max().item(), d.view(chk, S, D)[:, 128:].max().item()))
    print("err by col block: <48 %.3e >=48 %.3e" % (d.view(chk * S, heads, hd)[:, :, :48].max().item(), d.view(chk * S, heads, hd)[:, :, 48:].max().item()))
    # same reference with the bias quantised the way the kernel carries it (fp16 of rel / scale)
    qh_ = lambda t: (t / scale).half().float() * scale
    att2 = ((q * scale) @ k.transpose(-1, -2)).view(chk, heads, gh, gw, gh, gw) + qh_(rel_h)[..., :, None] + qh_(rel_w)[..., None, :]
    ref2 = (att2.view(chk, heads, S, S).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(chk * S, D)
    d2 = (out[:chk * S].float() - ref2).abs()
    print("vs the reference with fp16-quantised bias/scale: max %.3e mean %.3e" % (d2.max().item(), d2.mean().item()))
