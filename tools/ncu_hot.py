"""Hot instructions of an `ncu --page source --csv` export: python tools/ncu_hot.py <src.csv> [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
isrc, iss, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
body = rows[2:]
tot = sum(int(r[iss] or 0) for r in body)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
print("total samples", tot, "instructions", len(body))
for i, r in enumerate(body):
    s = int(r[iss] or 0)
    if s >= tot * thr / 100:
        print(f"{i:5d} {s:5d} {100 * s / tot:5.1f}% ex={r[iex]:>8s} {r[isrc][:120]}")
