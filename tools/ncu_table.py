"""Per-kernel table from an `ncu --csv --metrics ...` log (long format: one row per launch and metric).

usage: python tools/ncu_table.py <log.csv> [--per-launch] [--hbm GBs]
Prints, per kernel name: launches, mean duration, mean DRAM read / write bytes, achieved DRAM GB/s (and its fraction of
--hbm, default MEASURED_PEAKS.json), mean tensor-pipe %."""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
per_launch = "--per-launch" in sys.argv
hbm = None
if "--hbm" in sys.argv:
    hbm = float(sys.argv[sys.argv.index("--hbm") + 1])
else:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = json.load(open(p)).get("hbm_gbs", 6450.6) if os.path.exists(p) else 6650.0
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    v = float(r[vi].replace(",", "") or 0)
    unit = r[ui]
    if r[mi].startswith("gpu__time_duration"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)           # -> us
    elif "bytes" in r[mi]:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)  # -> bytes
    launch.setdefault((int(r[ii]), name), {})[r[mi]] = v


def fmt(name, n, d):
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / t / 1e3 if t > 0 else 0.0
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    s = f"{name[:48]:48s} n={n:4d} {t:9.1f} us  rd {rd / 1e6:8.2f} MB  wr {wr / 1e6:8.2f} MB  {gbs:7.0f} GB/s = {gbs / hbm:5.3f} of {hbm:.0f}"
    if tp is not None:
        s += f"  tensor {tp:5.1f}%"
    return s


if per_launch:
    for (i, name), d in launch.items():
        print(fmt(name, 1, d))
else:
    agg = collections.OrderedDict()
    for (i, name), d in launch.items():
        a = agg.setdefault(name, [0, collections.defaultdict(float)])
        a[0] += 1
        for k, v in d.items():
            a[1][k] += v
    tot = sum(a[1].get("gpu__time_duration.sum", 0.0) for a in agg.values())
    for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1].get("gpu__time_duration.sum", 0.0)):
        print(fmt(name, n, {k: v / n for k, v in s.items()}) + f"  total {s.get('gpu__time_duration.sum', 0.0):9.1f} us")
    print(f"sum of launch durations: {tot:.1f} us")
