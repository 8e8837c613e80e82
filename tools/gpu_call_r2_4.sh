#!/bin/bash
# Round 2, GPU call 4: persistent layout executor v2 (16 warps, 3xTF32 mma.sync contraction, GEGLU epilogue): tests, timing, ncu.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layout_mk_gpu.py -m gpu -x -q > gpurun_out/mk_tests.log 2>&1
echo "mk tests rc=$?"; tail -25 gpurun_out/mk_tests.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "layout" > gpurun_out/layout_tests.log 2>&1
echo "layout tests rc=$?"; tail -5 gpurun_out/layout_tests.log
timeout 200 python tools/time_layout.py fp32 2>&1 | tail -2
raw() { ncu -i "$1" --page raw --csv > "$2" 2>/dev/null; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'layout_mk' -c 1 \
   -o gpurun_out/r2_layout_mk -f python tools/profile_step.py --branch layout > gpurun_out/ncu_mk.log 2>&1; tail -2 gpurun_out/ncu_mk.log
raw gpurun_out/r2_layout_mk.ncu-rep gpurun_out/r2_ncu_layout_mk.csv
ncu -i gpurun_out/r2_layout_mk.ncu-rep --page source --csv > gpurun_out/r2_ncu_layout_mk_source.csv 2>/dev/null
rm -f gpurun_out/r2_layout_mk.ncu-rep
ls -la gpurun_out
