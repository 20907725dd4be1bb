"""profiles/r2_postproc_traffic_k12.json from an ncu launch list of ``CVB_POST=argmax python tools/prof_step.py`` (the product
pipeline's post-processing: cvb_postproc_argmax + cvb_contours + cvb_cell_tokens over 4 tiles of 1024^2, 700 nuclei each).

usage: python tools/postproc_traffic.py <launches.csv> <out.json>
Compulsory bytes per SURVEY.md section 8d for the arg-max entry: u8 NP + u8 NT + 2 x fp32 HV in, int32 labels out = 14 B/px."""
import collections
import csv
import json
import os
import re
import sys

POST = ("watershed_kernel", "ccl_merge_kernel", "ccl_compress_kernel", "ccl_init_kernel", "sobel_kernel", "morph5_kernel", "prep_float_kernel",
        "prep_argmax_kernel", "table_accum_kernel", "count_kernel", "blob_scatter_kernel", "blur_kernel", "energy_kernel", "minmax_f32_kernel",
        "marker_kernel", "table_init_kernel", "fill_kernel", "scan_apply_kernel", "scan_reduce_kernel", "scan_bsums_kernel", "blb_kernel",
        "blob_queue_kernel", "root_flag_kernel", "table_finalize_kernel", "border_flag_kernel", "contour_kernel", "cell_tokens_kernel")
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, mi, ui, vi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"[(<].*", "", name).split("::")[-1]
    if name not in POST:
        continue
    v = float(r[vi].replace(",", "") or 0)
    d = per.setdefault(name, {"launches": 0, "us": 0.0, "bytes": 0.0})
    if r[mi].startswith("gpu__time_duration"):
        d["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        d["launches"] += 1
    elif "bytes" in r[mi]:
        d["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
tiles, px = 4, 1024 * 1024
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = float(peaks.get("hbm_gbs", 6550.1))
flood = per.get("watershed_kernel", {"us": 0.0})["us"]
extra = sum(per.get(k, {"us": 0.0})["us"] for k in ("contour_kernel", "cell_tokens_kernel"))
total_us = sum(d["us"] for d in per.values())
total_b = sum(d["bytes"] for d in per.values())
map_us = total_us - flood - extra
map_b = total_b - sum(per.get(k, {"bytes": 0.0})["bytes"] for k in ("watershed_kernel", "contour_kernel", "cell_tokens_kernel"))
comp = 14.0 * px * tiles + tiles * 700 * 88
out = {"source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; one cvb_postproc_argmax "
                 "+ cvb_contours + cvb_cell_tokens pass over 4 tiles of 1024^2 with 700 nuclei each -- the product pipeline's path after the K12 "
                 "fusion; launches are serialised and cold-cache under ncu)",
       "tiles": tiles, "compulsory_bytes": comp, "compulsory": "14 B/px (u8 NP + u8 NT + 2 x fp32 HV in, int32 labels out, SURVEY 8d) + instance tables",
       "measured_dram_bytes": total_b, "traffic_over_compulsory": total_b / comp, "kernel_us_total": total_us, "kernel_us_flood": flood,
       "kernel_us_contours_and_tokens": extra, "kernel_us_map_stages": map_us, "map_stages_dram_bytes": map_b,
       "map_stages_GBps": map_b / map_us / 1e3 if map_us else 0.0, "hbm_peak_GBps": hbm,
       "frac_of_hbm_peak": (map_b / map_us / 1e3) / hbm if map_us else 0.0,
       "note": "dram__bytes counts HBM traffic only: the planes of one batch largely stay in the 126 MB L2 between consecutive kernels, so the map stages "
               "are bound by L2 latency / atomics / launch count rather than by HBM; the flood is serial per blob and latency-bound (no roofline "
               "fraction). Against the float entry (r2_postproc_traffic.json: 713.7 MB, prep_float_kernel 137.6 MB) the arg-max entry reads 14 "
               "instead of 44 B/px of input.",
       "per_kernel": {k: {"launches": d["launches"], "us": round(d["us"], 1), "dram_MB": round(d["bytes"] / 1e6, 2),
                          "GBps": round(d["bytes"] / d["us"] / 1e3, 1) if d["us"] else 0.0}
                      for k, d in sorted(per.items(), key=lambda kv: -kv[1]["us"])}}
json.dump(out, open(dst, "w"), indent=1)
print({k: v for k, v in out.items() if k not in ("per_kernel", "source", "note")})
