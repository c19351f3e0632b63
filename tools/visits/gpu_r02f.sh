#!/bin/bash
# r02f: streaming CHAIN with line-sized, warp-coalesced result writes (shared-memory stage): parity tests, then timings.
set -u
TAG=${1:-r02f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
R=10000000
V=$PWD/.variants
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for lay in c64 c96; do
  export SBWT_B200_LAYOUT=$lay
  echo "== c2 default(nch1mb4) $lay" | tee -a $OUT/quick.txt; q c2 $R
  echo "== c2 nch1mb3 $lay" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb3.so q c2 $R
  echo "== c2 nch2mb2 $lay" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch2mb2.so q c2 $R
done
export SBWT_B200_LAYOUT=c64
echo "== c2 default out32" | tee -a $OUT/quick.txt; QUICK_OUT32=1 q c2 $R
echo "== c2 nostore" | tee -a $OUT/quick.txt; QUICK_NOSTORE=1 SBWT_B200_LIB=$V/nch1mb4ns.so q c2 $R
echo "== c2q default" | tee -a $OUT/quick.txt; q c2q $R
echo "== c4s default" | tee -a $OUT/quick.txt; q c4s $R
echo "== c4s nch1mb3" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb3.so q c4s $R
echo "== c3 default" | tee -a $OUT/quick.txt; q c3 $R
echo "t=$(( $(date +%s) - T0 ))s"
