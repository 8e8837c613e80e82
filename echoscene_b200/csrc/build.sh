#!/bin/bash
# Builds libechoscene_b200.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $@"
pids=()
for f in runtime elem linear gemm_simt sgemm_x3 gemm_tc flash_tc conv_small gcn gcn_train scene unet_plan layout layout_mk train metrics mesh shape vqvae capi attention; do
  [ -f "$HERE/$f.cu" ] || continue
  if [ ! -f "$HERE/obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/obj/$f.o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.cuh' -o -name '*.inc' \) -newer "$HERE/obj/$f.o")" ] || [ "$HERE/../../include/echoscene_b200.h" -nt "$HERE/obj/$f.o" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT/libechoscene_b200.so" "$HERE"/obj/*.o -cudart static
echo "built $OUT/libechoscene_b200.so"
