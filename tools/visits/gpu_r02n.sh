#!/bin/bash
# r02n: the streaming walk as passes over global lists (walk_passes.cuh): GPU tests, then timings against the persistent kernel alone
set -u
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c4s c5s c2q; do
  echo "== $wl passes" | tee -a $OUT/quick.txt; q $wl 10000000
  echo "== $wl persistent kernel only" | tee -a $OUT/quick.txt; SBWT_B200_PASSES=0 q $wl 10000000
done
echo "== c2 passes out32" | tee -a $OUT/quick.txt; QUICK_OUT32=1 q c2 10000000
echo "t=$(( $(date +%s) - T0 ))s"
