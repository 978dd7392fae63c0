"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0        # launches are divided by this many steps
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"std::array<char \*.*", "", r[ki]).replace("native::", "").replace("void ", "").replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n) if n.startswith(("bn_", "roi_", "nms_", "conv_tc", "kmeans", "paste", "crop", "iou", "match")) else n
    agg[n[:150]][0] += 1
    agg[n[:150]][1] += v
    tot += v
OWN = ("bn_", "gn_", "roi_", "nms_", "conv_tc", "conv2_", "conv_wgrad", "wgrad2_", "stem_", "kmeans", "knn_", "paste", "crop", "iou", "match",
       "assign", "gather_sorted", "upsample_", "mask_loss", "maxpool3x3", "sum2x2", "preprocess_u8", "rpn_", "box_losses", "cascade_",
       "sgd_segments", "flip_", "deconv")
own = sum(v for n, (c, v) in agg.items() if n.startswith(OWN))
print("total %.3f ms/step over %d launches/step; libu2b200 kernels %.1f%% of kernel time" % (tot / per / 1e6, sum(c for c, _ in agg.values()) / per, 100 * own / tot))
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    print("%8.3f ms %6.1f/step %5.1f%% avg %6.1f us  %s" % (v / per / 1e6, c / per, 100 * v / tot, v / c / 1e3, n))
