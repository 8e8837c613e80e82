#!/bin/bash
set -u
mkdir -p gpurun_out
ECHO_MK_TIMELINE=gpurun_out/mk_timeline.bin timeout 200 python tools/profile_step.py --branch layout 2>&1 | tail -1
timeout 200 python tools/time_layout.py fp32 2>&1 | tail -1
