"""GPU timeline of one forward + post-processing step via the torch profiler (CUPTI): kernel time vs idle gaps."""
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from cellvit_b200.cellvit import CellViTSAM
from cellvit_b200 import synth
torch.manual_seed(0)
m = CellViTSAM(None, 6, 19, "SAM-H").eval().cuda()
x = torch.from_numpy(synth.synthetic_tiles(4, 1024, seed=1)).cuda()
with torch.no_grad():
    for _ in range(3): m(x, retrieve_tokens=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2): m(x, retrieve_tokens=True)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev], key=lambda t: t[0])
# second forward only
half = len(ks) // 2
ks = ks[half:]
busy = sum(e - s for s, e, _ in ks)
span = ks[-1][1] - ks[0][0]
gaps = [ks[i + 1][0] - ks[i][1] for i in range(len(ks) - 1)]
print(f"kernels {len(ks)} span {span/1e3:.2f} ms busy {busy/1e3:.2f} ms idle {(span-busy)/1e3:.2f} ms")
g = np.array(gaps)
print("gap us: mean %.2f median %.2f p90 %.2f max %.2f; >5us: %d" % (g.mean(), np.median(g), np.percentile(g, 90), g.max(), (g > 5).sum()))
big = sorted([(gaps[i], ks[i][2][:40], ks[i + 1][2][:40]) for i in range(len(gaps))], reverse=True)[:12]
for b in big: print("  %.1f us  after %s  before %s" % b)
import collections
agg = collections.Counter()
for s, e, n in ks: agg[n.split("(")[0][-40:]] += e - s
for n, t in agg.most_common(12): print(f"{t/1e3:8.3f} ms {n}")
