"""Small end-to-end pass (CellViT-256 + SAM-B forward on 256^2 tiles, post-processing, contours) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cellvit_b200 import synth, weights
from cellvit_b200.cellvit import CellViT256, CellViTSAM
from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor

for arch in ("ViT256", "SAM-B"):
    m = (CellViT256(None, 6, 19) if arch == "ViT256" else CellViTSAM(None, 6, 19, arch))
    m.load_state_dict(weights.synth_state_dict(arch, 6, 19, seed=3))
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(torch.from_numpy(synth.synthetic_tiles(1, 256, seed=5)).cuda(), retrieve_tokens=True)
    torch.cuda.synchronize()
    print(arch, "forward ok", float(out["hv_map"].abs().max()))
t = synth.synthetic_nuclei(256, 40, 0)
npl, ntl = synth.head_logits_from_maps(t["np_bin"], t["nt"], 6)
proc = DetectionCellPostProcessor(6, 40)
labels, dicts = proc.post_process_batch(torch.from_numpy(npl)[None].cuda(), torch.from_numpy(t["hv"])[None].cuda(), torch.from_numpy(ntl)[None].cuda())
torch.cuda.synchronize()
print("postproc ok", len(dicts[0]))
