"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline needs."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__grid_size", "smsp__cycles_active.avg", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {k: hdr.index(k) for k in KEYS if k in hdr}
tens = [h for h in hdr if "tensor" in h and "pct" in h]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:40]
    out = [name]
    for k, i in idx.items():
        out.append(f"{k.split('.')[0].replace('__','.')[-28:]}={r[i]}{units[i]}")
    print(" | ".join(out))
if len(sys.argv) > 2:
    print([h for h in hdr if sys.argv[2] in h])
