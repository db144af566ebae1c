"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
n = 0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    t = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    key = (name, row["Grid Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
    n += 1
print(f"{n} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-62s %-16s n=%4d tot=%9.1f us avg=%8.1f us %5.1f%%" % (k[0][:62], k[1], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
