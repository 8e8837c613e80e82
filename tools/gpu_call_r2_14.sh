#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "sharded or shard or components" 2>&1 | tail -8
