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
    m(x, retrieve_tokens=True); p.run_float(npd, hvd, ntd)
torch.cuda.synchronize()
