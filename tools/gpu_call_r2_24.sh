#!/bin/bash
# final 2-GPU sanity of the sharded bench (the driver's launch line), tight timeout
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
echo "bench N=2 rc=$?"; tail -c 300 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','launch_mode','gpu_launches','e2e','config4_scene_sharded','strong_scaling_one_scene'): print(k, d.get(k))
except Exception as e: print('parse failed', e)
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 400
