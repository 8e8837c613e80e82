#!/bin/bash
# 8 GPUs: the driver's scaling run, once, with a tight timeout (a hang costs 8x)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench N=8 rc=$?"; tail -c 300 gpurun_out/bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','launch_mode','e2e','config4_scene_sharded','strong_scaling_one_scene'): print(k, d.get(k))
except Exception as e: print('parse failed', e)
PY
