"""Summarise `ncu -i rep --page raw --csv` rows: one line per profiled launch with time, DRAM bytes and throughputs."""
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
col = {k: i for i, k in enumerate(h)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "% dram"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "% sm"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "% l1"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "% l2"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "% occ"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
for r in rows[2 if units else 1:]:
    name = r[col["Kernel Name"]][:60] if "Kernel Name" in col else "?"
    parts = []
    for k, lab in want:
        if k in col:
            u = units[col[k]] if units else ""
            parts.append("%s %s%s" % (r[col[k]], (u + " ") if u and lab in ("us", "MB rd", "MB wr") else "", lab))
    print("%-60s | %s" % (name, " | ".join(parts)))
