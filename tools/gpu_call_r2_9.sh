#!/bin/bash
# 2-GPU: does the graph-replayed sharded bench exit cleanly now?  (tight timeouts: a hang costs 2x GPU time)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench N=2 rc=$?"; tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','launch_mode'): print(k, d.get(k))
except Exception as e: print('parse failed', e)
PY
