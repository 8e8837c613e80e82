#!/bin/bash
# First gpurun call of a round: the GPU tests that have never run, the once-per-scene timings, a short bench line and its
# launch list.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
python -m pytest tests/test_zzz_vqenc_gpu.py tests/test_zz_scene_gpu.py -m gpu -x -q > gpurun_out/new_gpu_tests.log 2>&1
echo "new GPU tests rc=$?"; tail -5 gpurun_out/new_gpu_tests.log
python tools/time_scene.py > gpurun_out/time_scene.log 2>&1; cat gpurun_out/time_scene.log
python tools/sample_scene.py > gpurun_out/sample_scene.log 2>&1; tail -5 gpurun_out/sample_scene.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_scene.csv \
    python tools/time_scene.py 16 64 1 > gpurun_out/ncu_scene.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_scene.csv > gpurun_out/launches_scene_summary.txt 2>&1 || true
python -m pytest tests -m gpu -x -q > gpurun_out/all_gpu_tests.log 2>&1
echo "all GPU tests rc=$?"; tail -3 gpurun_out/all_gpu_tests.log
