#!/bin/bash
# walk2 bring-up: parity tests, then old vs new kernel timings. usage: bash tools/gpu_w2.sh <tag>
set -u
TAG=${1:-w2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
for wl in c2 c3; do
  for g in 1 2; do
    SBWT_B200_WALK=$g timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
for lib in .variants/*.so; do
  [ -f "$lib" ] || continue
  for wl in c2 c3; do
    echo "== $lib" | tee -a $OUT/quick.txt
    SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
