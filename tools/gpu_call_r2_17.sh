#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "gcn or GCN or train" 2>&1 | tail -12
