#!/bin/bash
# Round 2: the graph-replayed bench step at N = 1, full GPU test suite
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_graph.json 2> gpurun_out/bench_n1_graph.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1_graph.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_graph.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','launch_mode','gpu_launches','e2e'): print(k, d.get(k))
print('roofline frac', d['roofline'].get('frac'), 'step', d['roofline']['step']['frac'])
print('layout', d.get('layout_branch',{}).get('ms_per_step'), 'x3', d.get('parity_mode_x3'), 'config4', d.get('config4_scene_sharded'))
print('full chain', d.get('full_chain_seconds_per_scene'))
PY
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-graph', d['value'], d['ms_per_step'], d['launch_mode'])"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/gpu_tests.log
