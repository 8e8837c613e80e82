#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -s 2>&1 | tail -8
timeout 300 python tools/time_gcn_train.py --scenes 64 --out gpurun_out/gcn_train_timing_s64.json 2>&1 | grep -E "ms|speedup|nodes|triples"
timeout 300 python tools/time_gcn_train.py --scenes 8 --out gpurun_out/gcn_train_timing_s8.json 2>&1 | grep -E "ms|speedup|nodes|triples"
timeout 300 python tools/time_gcn_train.py --scenes 1 --out gpurun_out/gcn_train_timing_s1.json 2>&1 | grep -E "ms|speedup|nodes|triples"
