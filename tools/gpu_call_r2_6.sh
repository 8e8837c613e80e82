#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layout_mk_gpu.py -m gpu -x -q 2>&1 | tail -3
ECHO_MK_TIMELINE=gpurun_out/mk_timeline_v8.bin timeout 200 python tools/profile_step.py --branch layout 2>&1 | tail -1
timeout 200 python tools/time_layout.py fp32 2>&1 | tail -1
ECHO_MK_NO_BALANCE=1 timeout 200 python tools/time_layout.py fp32 2>&1 | tail -1
