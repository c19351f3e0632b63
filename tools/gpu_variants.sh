#!/bin/bash
# A/B of differently-compiled builds: bash tools/gpu_variants.sh <tag> [reads]
set -u
TAG=${1:-var}; R=${2:-2000000}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in c2 c3; do
  echo "== default" | tee -a $OUT/quick.txt
  timeout 300 python tools/quick_walk.py $wl $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  for lib in .variants/*.so; do
    [ -f "$lib" ] || continue
    echo "== $lib" | tee -a $OUT/quick.txt
    SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py $wl $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
