"""Split an ncu launch list of ONE static-graph training step (tools/profile_static.py) into phases by landmark
kernels and print, per phase, the launch count, the kernel time and the top kernels."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"(void |at::native::|<unnamed>::|at::)", "", r[ki])
    seq.append((re.sub(r"std::array<char \*.*", "", n)[:100], v / 1e3))
first = lambda pat, lo=0: next(i for i in range(lo, len(seq)) if pat in seq[i][0])      # noqa: E731
last = lambda pat: max(i for i in range(len(seq)) if pat in seq[i][0])                    # noqa: E731
a = first("upsample_ce_kernel") + 1          # end of backbone + FPN + semantic head forward
b = first("roi_align_fwd_kernel")            # RPN forward ends where the first ROI pooling starts (approx.)
c = last("crop_resize_masks_kernel") + 1     # ROI heads forward end
d = last("stem_wgrad_kernel") + 1            # backward end
phases = [("backbone + FPN + semantic head forward", 0, a), ("RPN forward (losses, decode, top-k, NMS)", a, b),
          ("ROI heads forward (3 cascade stages + mask)", b, c), ("backward", c, d),
          ("gradient gather, all-reduce, clip, SGD", d, len(seq))]
tot = sum(v for _, v in seq)
print("%d launches, %.2f ms of kernel time (serialised, cold cache)" % (len(seq), tot / 1e3))
for label, lo, hi in phases:
    sub = seq[lo:hi]
    t = sum(v for _, v in sub)
    print("== %-48s %4d launches %6.2f ms (%4.1f %%)" % (label, len(sub), t / 1e3, 100 * t / tot))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in sub:
        agg[n][0] += 1
        agg[n][1] += v
    for n, (cnt, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
        print("     %7.1f us %4d  %s" % (v, cnt, n[:90]))
