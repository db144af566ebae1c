"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel / grid."""
import collections
import csv
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        key = (row["Kernel Name"].split("(")[0][-78:], row["Grid Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    own = sum(t for (k, _), (_, t) in agg.items() if "fdfd::" in k)
    print(f"{n} launches, {tot:.1f} us total (cold-cache, serialised: compare shares); fdfd:: kernels {100 * own / tot:.1f} % of the device time")
    for (k, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:80s} {g:16s} n={c:5d} tot={t:10.1f} us avg={t / c:9.1f} us {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
