"""Join an ECHO_TRACE=1 log of one profiled step with the ncu launch list of the same run: per-contraction time,
TFLOP/s and tile-wave utilisation.  Usage: python tools/gemm_efficiency.py launches.csv trace.log"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
times = []
for r in rows[1:]:
    if "gemm_tc_kernel" in r[ki]:
        v = float(r[vi].replace(",", ""))
        times.append(v / 1000.0 if r[ui] in ("ns", "nsecond") else v)
tr = [l for l in open(sys.argv[2]) if l.startswith("[echo-trace] gemm_tc")]
tr = tr[-len(times):]
assert len(tr) == len(times), (len(tr), len(times))
agg = defaultdict(lambda: [0, 0.0, 0.0])
for l, us in zip(tr, times):
    d = dict(kv.split("=") for kv in l.split()[2:])
    rows_, cin, cout, k = int(d["rows"]), int(d["cin"]), int(d["cout"]), int(d["k"])
    fl = 2.0 * rows_ * cin * cout * k ** 3
    if int(d.get("up", 0)):
        fl = 2.0 * rows_ * cin * cout * 27   # algorithmic FLOPs of the reference's upsample + conv (the folded form does 12/27 of them)
    key = (rows_, cin, cout, k, int(d["stride"]), int(d["epi"]), int(d.get("up", 0)), int(d.get("bn", 0)), int(d.get("msub", 0)), int(d.get("splitk", 0)))
    a = agg[key]
    a[0] += 1; a[1] += us; a[2] += fl
tot_us = sum(a[1] for a in agg.values()); tot_fl = sum(a[2] for a in agg.values())
print(f"{'rows':>6s} {'cin':>5s} {'cout':>5s} k s e u  bn m sk {'n':>3s} {'us/launch':>9s} {'share':>6s} {'TFLOP/s':>8s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{key[0]:6d} {key[1]:5d} {key[2]:5d} {key[3]} {key[4]} {key[5]} {key[6]} {key[7]:3d} {key[8]} {key[9]:2d} {a[0]:3d} {a[1]/a[0]:9.1f} {100*a[1]/tot_us:5.1f}% {a[2]/a[1]/1e6:8.0f}")
print(f"total {tot_us/1000:.3f} ms, {tot_fl/1e12:.3f} TFLOP, {tot_fl/tot_us/1e6:.0f} TFLOP/s average")
