#!/bin/bash
# launch list of one GCN training iteration (64 scenes) after the optimisations + ncu --set full of the 3xTF32 GEMM launches in it
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gcn_train_launches.csv python tools/time_gcn_train.py --scenes 64 --profile-one-step > gpurun_out/gcn_train_ncu.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/gcn_train_launches.csv > gpurun_out/gcn_train_launch_summary.txt 2>&1; head -30 gpurun_out/gcn_train_launch_summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:sgemm_x3 -c 18 -o gpurun_out/sgemm_x3_full -f python tools/time_gcn_train.py --scenes 64 --profile-one-step > gpurun_out/sgemm_x3_ncu.log 2>&1
echo "set full rc=$?"
ncu -i gpurun_out/sgemm_x3_full.ncu-rep --page raw --csv > gpurun_out/sgemm_x3_full.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/sgemm_x3_full.csv > gpurun_out/sgemm_x3_full_summary.txt 2>&1; head -40 gpurun_out/sgemm_x3_full_summary.txt
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/sgemm_x3_full.csv")))
hdr = rows[0]
want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) for w in want if w in hdr]
print([hdr[i] for i in idx])
for r in rows[2:8]:
    print([r[i][:40] for i in idx])
PY
rm -f gpurun_out/sgemm_x3_full.ncu-rep
