"""One post-processing pass (cvb_postproc + cvb_contours) at the BASELINE shape (4 tiles of 1024^2, 700 synthetic nuclei each)
for ncu: python tools/prof_post.py [n_passes]. Warm-up pass first (profile with `ncu -s <launches of one pass>` or take the
last pass of the log)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cellvit_b200 import synth  # noqa: E402
from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 4
nuc = [synth.synthetic_nuclei(1024, 700, seed=i) for i in range(B)]
lg = [synth.head_logits_from_maps(x["np_bin"], x["nt"], 6) for x in nuc]
npd = torch.from_numpy(np.stack([l[0] for l in lg])).cuda()
ntd = torch.from_numpy(np.stack([l[1] for l in lg])).cuda()
hvd = torch.from_numpy(np.stack([x["hv"] for x in nuc])).cuda()
p = DetectionCellPostProcessor(6, 40)
for _ in range(n):
    p.launch_float(npd, hvd, ntd, slot=0)
    torch.cuda.synchronize()
print("instances per tile:", p._wsp.host[0]["counts"].tolist())
