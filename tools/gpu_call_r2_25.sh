#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -k "cuda_graph or accumulates" 2>&1 | tail -8
for s in 1 8 64; do timeout 300 python tools/time_gcn_train.py --scenes $s --out gpurun_out/gcn_train_timing_s$s.json 2>&1 | grep -E "ms|speedup|nodes"; done
