#!/bin/bash
# r03h: classic-sector kernels at 3 blocks/SM (80 registers, no spills) against the default 4 blocks/SM (64 registers, 62 bytes of spills)
set -u
TAG=${1:-r03h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c4s c5s c3; do
  echo "== $wl default" | tee -a $OUT/quick.txt; q $wl 10000000
  echo "== $wl mb3" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$PWD/.variants/mb3.so q $wl 10000000
done
