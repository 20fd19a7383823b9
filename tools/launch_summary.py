"""launch_summary.py <ncu launches csv> -- per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum` log."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
    agg[r[kn][:110]][0] += 1
    agg[r[kn][:110]][1] += v
    tot += v
print("# %s: %d launches, %.1f us total (cold-cache, serialised under ncu: compare shares)" % (sys.argv[1], sum(a[0] for a in agg.values()), tot))
print("%10s %6s %5s %9s  kernel" % ("total_us", "share", "n", "avg_us"))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%10.1f %5.1f%% %5d %9.1f  %s" % (t, 100 * t / tot, c, t / c, k))
