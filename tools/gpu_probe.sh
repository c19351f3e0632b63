#!/bin/bash
# probe-stride bring-up: parity tests, then walk timings per stride. usage: bash tools/gpu_probe.sh <tag>
set -u
TAG=${1:-pr}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
for wl in c2 c4s c5s; do
  for d in 0 default; do
    if [ $d = default ]; then unset SBWT_B200_PROBE; else export SBWT_B200_PROBE=$d; fi
    timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
unset SBWT_B200_PROBE
for lib in .variants/*.so; do
  [ -f "$lib" ] || continue
  for wl in c2 c4s; do
    echo "== $lib" | tee -a $OUT/quick.txt
    SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
for d in 12 14 16 18; do
  SBWT_B200_PROBE=$d timeout 300 python tools/quick_walk.py c2 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
done
