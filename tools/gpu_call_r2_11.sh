#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/time_optimizer.py 2>&1 | tail -3 | tee gpurun_out/optimizer.json
