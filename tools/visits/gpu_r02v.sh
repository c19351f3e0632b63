#!/bin/bash
# r02v: code-word refill issued with the step's sector load + register-held stage row address, against the r02s library (base)
set -u
TAG=${1:-r02v}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c3 c4s c5s; do
  for lib in base new; do
    echo "== $wl $lib" | tee -a $OUT/quick.txt
    if [ $lib = base ]; then SBWT_B200_LIB=$PWD/.variants/base.so q $wl 10000000; else q $wl 10000000; fi
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
