#!/bin/bash
# Round 2, second GPU call: the persistent layout executor (tests, timing, ncu), and the ncu captures of the non-GEMM kernel
# families as CSV (the .ncu-rep files are too large to travel back).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layout_mk_gpu.py -m gpu -x -q > gpurun_out/mk_tests.log 2>&1
echo "mk tests rc=$?"; tail -25 gpurun_out/mk_tests.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "layout" > gpurun_out/layout_tests.log 2>&1
echo "layout tests rc=$?"; tail -5 gpurun_out/layout_tests.log
timeout 200 python tools/time_layout.py fp32 2>&1 | tail -2
ECHO_NO_MK=1 timeout 200 python tools/time_layout.py fp32 2>&1 | tail -2
raw() {  # raw <rep> <csv>: every metric of every captured launch as CSV, then drop the report
  ncu -i "$1" --page raw --csv > "$2" 2>/dev/null; rm -f "$1"
}
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'layout_mk' -c 1 \
   -o gpurun_out/r2_layout_mk -f python tools/profile_step.py --branch layout > gpurun_out/ncu_mk.log 2>&1; tail -2 gpurun_out/ncu_mk.log
raw gpurun_out/r2_layout_mk.ncu-rep gpurun_out/r2_ncu_layout_mk.csv
ECHO_NO_MK=1 ECHO_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --profile-from-start off \
   -k regex:'linear_rows|edge_combine|node_pool|embedding_rows' -c 40 -o gpurun_out/r2_layout_kernels -f \
   python tools/profile_step.py --branch layout > gpurun_out/ncu_layout.log 2>&1; tail -2 gpurun_out/ncu_layout.log
raw gpurun_out/r2_layout_kernels.ncu-rep gpurun_out/r2_ncu_layout_kernels.csv
timeout 600 ncu --set full --clock-control none --profile-from-start off \
   -k regex:'gn_apply_cs|layer_norm|edge_combine|node_pool|ddim_update|splitk_reduce|linear_rows' -c 40 -o gpurun_out/r2_shape_elem -f \
   python tools/profile_step.py --branch shape > gpurun_out/ncu_shape.log 2>&1; tail -2 gpurun_out/ncu_shape.log
raw gpurun_out/r2_shape_elem.ncu-rep gpurun_out/r2_ncu_shape_elem.csv
du -sh gpurun_out; ls -la gpurun_out
