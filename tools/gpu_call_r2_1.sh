#!/bin/bash
# Round 2, first GPU call: bf16 drift over the benched chain, full-size oracle parity of the existing modes, the bench line with
# the real reference arms, ncu captures of the non-GEMM kernel families.
set -u
mkdir -p gpurun_out
python tools/bf16_drift.py --ref fp32 --test bf16 --out gpurun_out/r2_bf16_drift_vs_fp32.json > gpurun_out/drift.log 2>&1; tail -3 gpurun_out/drift.log
python -m pytest tests/test_parity_full_gpu.py -m gpu -q -s -k "test_shape_step_full_size_vs_oracle and not x3" > gpurun_out/parity_full.log 2>&1
echo "parity rc=$?"; grep -E "parity\]|passed|failed" gpurun_out/parity_full.log | tail -12
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 800 gpurun_out/bench_ref.json
ECHO_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'linear_rows|edge_combine|node_pool|embedding_rows' -c 40 -o gpurun_out/r2_layout_kernels -f \
   python tools/profile_step.py --branch layout > gpurun_out/ncu_layout.log 2>&1; tail -2 gpurun_out/ncu_layout.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'gn_apply_cs|layer_norm|edge_combine|node_pool|ddim_update|splitk_reduce' -c 24 -o gpurun_out/r2_shape_elem -f \
   python tools/profile_step.py --branch shape > gpurun_out/ncu_shape.log 2>&1; tail -2 gpurun_out/ncu_shape.log
ls -la gpurun_out
