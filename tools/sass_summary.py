"""Developer tool: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / cluster use in libu2b200.so
(cuobjdump -sass; see B200_PROFILING.md). Usage: python tools/sass_summary.py > profiles/rNN_sass_tcgen05_kernels.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "u2seg_b200/libu2b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR_ARV", "REDG", "ATOMG"]
counts, variants, name = collections.defaultdict(collections.Counter), collections.defaultdict(set), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        counts[name][op.split(".")[0]] += 1
        counts[name]["_total"] += 1
        if op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR"):
            variants[name].add(op)
demangled = {}
try:
    out = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(counts, out))
except Exception:
    pass
print("# kernels of %s that issue tcgen05 (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc)" % so)
print("# or TMA (UTMALDG = cp.async.bulk.tensor load, UTMASTG = store); SYNCS = mbarrier ops; UCGABAR = cluster barrier")
print("%-78s %7s " % ("kernel", "instrs") + " ".join("%10s" % k for k in KEYS))
for n in sorted(counts, key=lambda k: demangled.get(k, k)):
    c = counts[n]
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"]):
        continue
    d = re.sub(r"\(anonymous namespace\)::", "", demangled.get(n, n))
    d = d.split(">(")[0] + ">" if ">(" in d else re.sub(r"\(.*", "", d)
    d = d.replace("void ", "").replace("<unnamed>::", "")
    print("%-78s %7d " % (d[:78], c["_total"]) + " ".join("%10d" % c[k] for k in KEYS))
    print("      variants: " + ", ".join(sorted(variants[n])))
