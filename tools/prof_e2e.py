"""Host-side timeline of CellSegmentationInference.process_tiles (where does e2e lose against the device-resident value?)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from cellvit_b200.cellvit import CellViTSAM
from cellvit_b200.cell_detection import CellSegmentationInference
from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor
from cellvit_b200 import synth
import torch.nn.functional as F
torch.manual_seed(0)
B = 4
m = CellViTSAM(None, 6, 19, "SAM-H").eval().cuda()
tiles_host = torch.from_numpy(synth.synthetic_tiles(B, 1024, seed=1)).pin_memory()
nuc = [synth.synthetic_nuclei(1024, 700, seed=i) for i in range(B)]
lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
ov = {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(), "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
      "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()}
inf = CellSegmentationInference.from_model(m, 0)
inf.process_tiles([tiles_host] * 2, 40, head_override=ov)
torch.cuda.synchronize()
proc = DetectionCellPostProcessor(6, 40)
T = lambda: time.perf_counter()
log = []
with torch.no_grad():
    pending = None
    t_start = T()
    for k in range(10):
        t0 = T(); p = tiles_host.to("cuda", non_blocking=True)
        t1 = T(); pred = m.forward(p, retrieve_tokens=True); pred.update(ov)
        t2 = T(); a = F.softmax(pred["nuclei_binary_map"], 1); b = F.softmax(pred["nuclei_type_map"], 1)
        t3 = T(); proc.launch_float(a, pred["hv_map"], b, slot=k & 1)
        t4 = T()
        tw = tb = 0.0
        if pending is not None:
            h = proc._wsp.host[pending]; h["event"].synchronize(); t5 = T(); proc.collect(pending); t6 = T(); tw, tb = t5 - t4, t6 - t5
        pending = k & 1
        log.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, tw, tb))
    proc.collect(pending)
    torch.cuda.synchronize()
    tot = T() - t_start
print("per batch ms: h2d-enq fwd-enq softmax-enq post-enq wait-event build-dicts")
for r in log: print(" ".join(f"{1e3*v:8.2f}" for v in r))
print(f"total {1e3*tot/10:.2f} ms/batch -> {B*10/tot:.1f} tiles/s")
# device-only timings of the e2e extras
def dev_ms(fn, n=10):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    print("h2d ms", dev_ms(lambda: tiles_host.to("cuda", non_blocking=True)))
    print("fwd ms", dev_ms(lambda: m.forward(p, retrieve_tokens=True)))
    print("softmax ms", dev_ms(lambda: (F.softmax(ov["nuclei_binary_map"], 1), F.softmax(ov["nuclei_type_map"], 1))))
    w = proc._wsp
    print("post+contours+d2h ms", dev_ms(lambda: proc.launch_float(a, ov["hv_map"], b, slot=0)))
    print("d2h only ms", dev_ms(lambda: w.copy_to_host(0, 2048)))
    print("contours ms", dev_ms(lambda: w.launch_contours(B, 1024, 1024, proc.max_rows)))
