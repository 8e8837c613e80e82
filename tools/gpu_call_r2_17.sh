#!/bin/bash
# GCN backward tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -30
