"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^.*::", "", name)
    name = re.sub(r"<.*", "", name)
    tot[name] += us
    cnt[name] += 1
all_us = sum(tot.values())
print(f"launches {sum(cnt.values())}  total {all_us/1000:.3f} ms (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':44s} {'n':>5s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:44s} {cnt[k]:5d} {tot[k]:10.1f} {100*tot[k]/all_us:6.1f}% {tot[k]/cnt[k]:9.2f}")
