"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total time, share."""
import csv, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    k = r[ik]
    t = float(r[iv].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {n:8d} {t/1e3:12.1f} {t/1e3/n:10.2f} {100*t/tot:6.1f}%")
