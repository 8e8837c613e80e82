#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_zz_scene_gpu.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/time_step.py --precision bf16 --steps 50 2>&1 | tail -1
