#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_zz_scene_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -8
timeout 500 python tools/time_layout_batch.py 1 4 16 64 2>&1 | tail -5
