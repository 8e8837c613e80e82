#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_metrics_gpu.py -m gpu -x -q 2>&1 | tail -15
