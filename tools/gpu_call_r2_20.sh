#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -s 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gcn_train_launches.csv python tools/time_gcn_train.py --scenes 64 --reps 1 --out gpurun_out/gcn_train_timing_ncu.json > gpurun_out/gcn_train_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = []
with open("gpurun_out/gcn_train_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    try:
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]))
    except Exception:
        pass
# the last my_step + my_fwd are at the end; find the last occurrence window: take the last 400 launches
tail = rows[-400:]
agg = collections.OrderedDict()
for name, v, u in tail:
    k = name.split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v / (1000.0 if u.startswith("n") else 1.0)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={n:4d} total_us={t:9.1f}")
PY
