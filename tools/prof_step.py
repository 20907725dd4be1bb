import os
import torch, sys, numpy as np
sys.path.insert(0, ".")
from cellvit_b200.cellvit import CellViT256, CellViTSAM
from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor
from cellvit_b200 import synth
from cellvit_b200 import _lib as _L
torch.manual_seed(0)
ARCH = os.environ.get("CVB_ARCH", "SAM-H")
B = 8 if ARCH == "ViT256" else 4
m = (CellViT256(None, 6, 19) if ARCH == "ViT256" else CellViTSAM(None, 6, 19, ARCH)).eval().cuda()
if len(sys.argv) > 1: m.set_engine_option("attention_tc", int(sys.argv[1]))
if os.environ.get("CVB_DYN") is not None: m.set_engine_option("dynamic_tiles", int(os.environ["CVB_DYN"]))   # attention mode bits (see csrc/model.cu)
x = torch.from_numpy(synth.synthetic_tiles(B, 1024, seed=1)).cuda()
nuc = [synth.synthetic_nuclei(1024, 700, seed=i) for i in range(B)]
lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
npd = torch.from_numpy(np.stack([l[0] for l in lg])).cuda(); ntd = torch.from_numpy(np.stack([l[1] for l in lg])).cuda()
hvd = torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()
p = DetectionCellPostProcessor(6, 40)
with torch.no_grad():
    if os.environ.get("CVB_POST", "float") == "argmax":
        # the product pipeline's path (K12): the head epilogue writes the uint8 arg-max planes, cvb_postproc_argmax consumes planes
        # (here: the planes of the injected synthetic-nuclei maps), contours and cell tokens follow
        out = m(x, retrieve_tokens=True, argmax_maps=True)
        p.launch_argmax(npd.argmax(1).to(torch.uint8), hvd, ntd.argmax(1).to(torch.uint8), slot=0, tokens=out["tokens"], patch_size=16)
        torch.cuda.synchronize()
        p.collect(0, None, with_tokens=True)
    else:
        m(x, retrieve_tokens=True); p.run_float(npd, hvd, ntd)
torch.cuda.synchronize()
