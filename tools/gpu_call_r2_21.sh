#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gcn_train_launches.csv python tools/time_gcn_train.py --scenes 64 --profile-one-step > gpurun_out/gcn_train_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/gcn_train_launches.csv | tee gpurun_out/gcn_train_launch_summary.txt
