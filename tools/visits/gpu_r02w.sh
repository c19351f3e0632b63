#!/bin/bash
# r02w: A/B of compile-time variants on the r02v kernel: two chains per lane at 2 blocks/SM, round length, block size, singleton hold
set -u
TAG=${1:-r02w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c4s; do
  echo "== $wl default" | tee -a $OUT/quick.txt; q $wl 10000000
  for v in nch2mb2 r8 r32 t224mb4 t128mb6 sh0 sh1; do
    echo "== $wl $v" | tee -a $OUT/quick.txt
    SBWT_B200_LIB=$PWD/.variants/$v.so q $wl 10000000
  done
done
