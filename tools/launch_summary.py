"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, average, share."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    k = re.sub(r"\(.*", "", name)[:80]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{sum(a[0] for a in agg.values())} launches, {tot:.1f} us")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:82s} n={a[0]:5d} total={a[1]:10.1f}us avg={a[1] / a[0]:8.1f}us share={100 * a[1] / tot:5.1f}%")
