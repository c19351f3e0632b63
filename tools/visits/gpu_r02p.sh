#!/bin/bash
# r02p: multi-pass walk with the probes of the first k-mers' ranges running beside the chains (two streams)
set -u
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c2q c4s; do
  echo "== $wl passes, overlapped" | tee -a $OUT/quick.txt; q $wl 10000000
done
echo "== c2 persistent kernel only" | tee -a $OUT/quick.txt; SBWT_B200_PASSES=0 q c2 10000000
echo "t=$(( $(date +%s) - T0 ))s"
