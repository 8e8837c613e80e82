#!/bin/bash
# training forward: two-arm loss parity
mkdir -p gpurun_out
timeout 900 python tools/trainfwd_check.py --out gpurun_out/trainfwd.json > gpurun_out/trainfwd.log 2>&1
echo "trainfwd rc=$?"
tail -40 gpurun_out/trainfwd.log
