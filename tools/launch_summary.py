"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_summary.py <file.csv>"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg, total, n = collections.OrderedDict(), 0.0, 0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    us = v / 1000.0 if row["Metric Unit"].startswith("n") else v
    k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("echo::<unnamed>::", "")[:64]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
print(f"launches {n}  total {total / 1000:.3f} ms (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':66s} {'n':>4s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} {c:4d} {t:10.1f} {100 * t / total:6.1f}% {t / c:8.2f}")
