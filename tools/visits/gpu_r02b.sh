#!/bin/bash
# r02b: first run of the rewritten streaming CHAIN (phase-aligned register staging, several chains per lane, csector64):
# GPU parity tests, then A/B of the compiled variants x csector formats on c2 at full size, c3 / c4s with the default build.
set -u
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
R=10000000
for lay in c64 c96; do
  echo "== default build, layout $lay" | tee -a $OUT/quick.txt
  SBWT_B200_LAYOUT=$lay timeout 300 python tools/quick_walk.py c2 $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  for lib in .variants/*.so; do
    echo "== $lib, layout $lay" | tee -a $OUT/quick.txt
    SBWT_B200_LAYOUT=$lay SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py c2 $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
echo "t=$(( $(date +%s) - T0 ))s"
for wl in c3 c4s; do
  echo "== default build $wl" | tee -a $OUT/quick.txt
  timeout 300 python tools/quick_walk.py $wl $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
