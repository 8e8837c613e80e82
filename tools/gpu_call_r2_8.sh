#!/bin/bash
# Round 2: the sharded bench at 2 GPUs: graph replay with NCCL inside, strong scaling and scene-sharded lines; gloo-free
set -u
mkdir -p gpurun_out
for mode in "" "--no-graph"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 $mode > gpurun_out/bench_n2$mode.json 2> gpurun_out/bench_n2$mode.err
  echo "bench N=2 $mode rc=$?"; tail -c 400 gpurun_out/bench_n2$mode.err
  python - "$mode" <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_n2{m}.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','launch_mode','gpu_launches','e2e','config4_scene_sharded','strong_scaling_one_scene'): print(k, d.get(k))
except Exception as e: print('parse failed', e)
PY
done
