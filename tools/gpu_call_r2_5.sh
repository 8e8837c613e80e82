#!/bin/bash
# Round 2, GPU call 5: persistent layout executor v2b: tests, timing, per-stage timeline; shape step after the GroupNorm prologue change.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layout_mk_gpu.py -m gpu -x -q > gpurun_out/mk_tests.log 2>&1
echo "mk tests rc=$?"; tail -8 gpurun_out/mk_tests.log
timeout 200 python tools/time_layout.py fp32 2>&1 | tail -1
ECHO_MK_TIMELINE=gpurun_out/mk_timeline.bin timeout 200 python tools/profile_step.py --branch layout 2>&1 | tail -1
timeout 300 python tools/time_step.py --precision bf16 --steps 50 2>&1 | tail -1
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -x -q > gpurun_out/model_tests.log 2>&1
echo "model+ops tests rc=$?"; tail -5 gpurun_out/model_tests.log
ls -la gpurun_out
