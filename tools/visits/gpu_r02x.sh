#!/bin/bash
# r02x: PROBE with two batches of probes in flight, against the r02v library
set -u
TAG=${1:-r02x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c4s c5s; do
  for lib in r02v new ${EXTRA:-}; do
    echo "== $wl $lib" | tee -a $OUT/quick.txt
    if [ $lib = new ]; then q $wl 10000000; else SBWT_B200_LIB=$PWD/.variants/$lib.so q $wl 10000000; fi
  done
done
