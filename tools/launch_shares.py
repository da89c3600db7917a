"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, n = defaultdict(lambda: [0, 0.0]), 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
    n += 1
tot = sum(t for _, t in agg.values())
print("launches: %d   total %.3f ms (cold-cache, serialised under ncu: compare SHARES)" % (n, tot / 1e6))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-32s n=%4d total=%10.3f ms share=%.4f avg=%9.1f us" % (k[:32], c, t / 1e6, t / tot, t / c / 1e3))
