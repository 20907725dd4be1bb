"""Forward-only A/B of an engine option: python tools/ab_forward.py <option> <v0> <v1> [arch] -- per value: graph-replayed forward
time (CUDA events, 10 replays) and the tile engine's summed launch time / executed TFLOP/s from cvb_tc_profile."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from cellvit_b200 import _lib as L, synth  # noqa: E402
from cellvit_b200.cellvit import CellViT256, CellViTSAM  # noqa: E402
opt, vals = sys.argv[1], [int(v) for v in sys.argv[2:4]]
arch = sys.argv[4] if len(sys.argv) > 4 else "SAM-H"
B = 4 if arch != "ViT256" else 8
torch.manual_seed(0)
m = (CellViTSAM(None, 6, 19, arch) if arch != "ViT256" else CellViT256(None, 6, 19)).eval().cuda()
x = torch.from_numpy(synth.synthetic_tiles(B, 1024, seed=1)).cuda()
lib = L.lib()
for rep in range(2):
    for v in vals:
        m.set_engine_option(opt, v)
        with torch.no_grad():
            for _ in range(3):
                m.forward_graphed(x, True, 0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                m.forward_graphed(x, True, 0)
            b.record()
            torch.cuda.synchronize()
            ms, n, fl = C.c_double(), C.c_int(), C.c_double()
            L.check(lib.cvb_tc_profile_begin(4096), "begin")
            m(x, retrieve_tokens=True)
            L.check(lib.cvb_tc_profile_end(C.byref(ms), C.byref(n), C.byref(fl)), "end")
        print(f"{opt}={v}: forward {a.elapsed_time(b) / 10:.2f} ms; tile engine {ms.value:.2f} ms in {n.value} launches = {fl.value / ms.value / 1e9:.0f} TFLOP/s executed")
