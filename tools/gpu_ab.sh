#!/bin/bash
# A/B of env knobs: bash tools/gpu_ab.sh <tag> "<env1>" "<env2>" ...   (each env string like "SBWT_B200_BLOCKS_PER_SM=5")
set -u
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for e in "$@"; do
  for wl in c2 c3; do
    env $e timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/ab.txt
  done
done
