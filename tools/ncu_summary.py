"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a compact per-kernel table for profiles/."""
import csv, subprocess, sys, collections

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    seen = collections.OrderedDict()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].replace("<unnamed>::", "")[:90]
        seen.setdefault(name, []).append(d)
    for name, ds in seen.items():
        d = ds[len(ds) // 2]
        print("%s   (%d captured launches; median one shown)" % (name, len(ds)))
        for k, short in KEYS:
            if k in d and d[k] != "":
                print("    %-16s %14s %s" % (short, d[k], units.get(k, "")))

if __name__ == "__main__":
    main(sys.argv[1])
