"""Compact per-launch summary of an ncu --set full report (run here, on the CPU box):
    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/X.csv;  python profiles/summarize_ncu.py /tmp/X.csv"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "dur"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__cycles_active.avg", "cycles")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(" | ".join(["kernel".ljust(34)] + [k[1].rjust(9) for k in KEYS]))
for d in data:
    name = d[col["Kernel Name"]][:34].ljust(34)
    vals = []
    for key, _ in KEYS:
        i = col.get(key)
        v = d[i] if i is not None else "-"
        u = units[i] if i is not None else ""
        try:
            vals.append(("%.4g%s" % (float(v), {"Mbyte": "MB", "Gbyte": "GB", "Kbyte": "KB", "ms": "ms", "us": "us"}.get(u, ""))).rjust(9))
        except ValueError:
            vals.append(v[:9].rjust(9))
    print(" | ".join([name] + vals))
