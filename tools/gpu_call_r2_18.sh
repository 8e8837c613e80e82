#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_gcn_train.py --scenes 64 --out gpurun_out/gcn_train_timing_s64.json 2>&1 | tail -16
timeout 300 python tools/time_gcn_train.py --scenes 8 --out gpurun_out/gcn_train_timing_s8.json 2>&1 | tail -16
