"""Per-event timeline (clock64 of SM of CTA 0) of the first items of window_tc_kernel: python tools/trace_window.py [n_items]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L  # noqa: E402
n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 100
heads, hd, S = 16, 80, 196
D = heads * hd
g = torch.Generator(device="cuda").manual_seed(13)
qkv = torch.randn(n_items * S, 3 * D, device="cuda", generator=g).half()
relcat = (torch.randn(64, hd, device="cuda", generator=g) * 0.2).half()
out = torch.empty(n_items * S, D, device="cuda", dtype=torch.half)
lib = L.lib()
import os
if os.environ.get("CVB_WSKEW") is not None:
    lib.cvb_debug_window_skew(int(os.environ["CVB_WSKEW"]))
def run():
    L.check(lib.cvb_op_window_attention_tc(L.ptr(qkv), n_items, heads, hd, C.c_float(hd ** -0.5), L.ptr(relcat), L.ptr(out), None, L.stream_ptr()), "w")
for _ in range(3): run()
torch.cuda.synchronize()
tr = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
lib.cvb_debug_window_trace(C.c_void_p(tr.data_ptr()))
run(); torch.cuda.synchronize()
lib.cvb_debug_window_trace(None)
t = tr.cpu().view(8, 64)
t0 = int(t[t > 0].min())
names = {0: "tma.qk_free", 1: "tma.v_free"}
for g_ in (0, 1):
    b = 8 + 8 * g_
    names.update({b: f"iss{g_}.qk_full(next)", b + 1: f"iss{g_}.G(first)", b + 2: f"iss{g_}.qg_ready->S", b + 3: f"iss{g_}.v_full", b + 4: f"iss{g_}.p_full->PV", b + 5: f"iss{g_}.o_full->G(next)"})
for g_ in (0, 1):
    for h in (0, 1):
        b = 24 + 8 * g_ + 16 * h
        names.update({b: f"thr{g_}{h}.g_full", b + 1: f"thr{g_}{h}.gsel_done", b + 2: f"thr{g_}{h}.s_full", b + 3: f"thr{g_}{h}.pass1_done", b + 4: f"thr{g_}{h}.p_done",
                      b + 5: f"thr{g_}{h}.o_full", b + 6: f"thr{g_}{h}.out_done"})
for it in range(2, 6):
    ev = sorted((int(t[it, e]) - t0, names[e]) for e in names if t[it, e] > 0)
    print(f"--- item {it}")
    for c, n in ev:
        print(f"{c:8d}  {n}")
