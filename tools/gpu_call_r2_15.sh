#!/bin/bash
# Round 2: full verification on one GPU -- smoke(), the whole -m gpu suite, the default bench line, the reference arm
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','steps','launch_mode','gpu_launches','e2e','clocks','cpu_baseline'): print(k, d.get(k))
print('roofline frac', d['roofline'].get('frac'), 'step', d['roofline']['step']['frac'])
for k in ('layout_branch','layout_branch_batched_64_scenes','sdf_to_mesh','full_chain_seconds_per_scene_batched','parity_mode_x3','config3_n32_s250','config4_scene_sharded','gpu_eager_baseline','scene_encode','vqvae_decode','full_chain_seconds_per_scene'): print(k, d.get(k))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -c 700
