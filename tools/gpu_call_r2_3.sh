#!/bin/bash
# Round 2, third GPU call: the split-precision (ECHO_PREC_X3) shape path -- parity at the benched sizes, its speed --
# and the ncu capture of the shape step's elementwise kernel families.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_full_gpu.py -m gpu -x -q -s -k "x3" > gpurun_out/x3_tests.log 2>&1
echo "x3 tests rc=$?"; grep -E "parity\]|passed|failed|Error|error" gpurun_out/x3_tests.log | tail -15
timeout 300 python tools/time_step.py --precision x3 --steps 10 2>&1 | tail -2
timeout 300 python tools/time_step.py --precision bf16 --steps 50 2>&1 | tail -1
raw() { ncu -i "$1" --page raw --csv > "$2" 2>/dev/null; rm -f "$1"; }
timeout 600 ncu --set full --clock-control none --profile-from-start off \
   -k regex:'gn_apply_cs|layer_norm|ddim_update|splitk_reduce|geglu|s2d_kernel|ncdhw' -c 40 -o gpurun_out/r2_shape_elem -f \
   python tools/profile_step.py --branch shape > gpurun_out/ncu_shape.log 2>&1; tail -2 gpurun_out/ncu_shape.log
raw gpurun_out/r2_shape_elem.ncu-rep gpurun_out/r2_ncu_shape_elem2.csv
ls -la gpurun_out
