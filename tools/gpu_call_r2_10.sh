#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zzzzz_refbind_gpu.py -m gpu -x -q -s > gpurun_out/refbind.log 2>&1; echo "refbind rc=$?"; grep -v "it/s\|it \[" gpurun_out/refbind.log | tail -25
