#!/bin/bash
# r02h: staged-output kernel: L2 set-aside sweep, and two chains per lane at 3 blocks/SM
set -u
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
V=$PWD/.variants
for cfg in "c64 nch1mb3" "c96 nch1mb3" "c64 nch2r8mb3" "c96 nch2r8mb3"; do
  set -- $cfg
  SBWT_B200_LAYOUT=$1 SBWT_B200_LIB=$V/$2.so timeout 300 python tools/exp_persist.py c2 10000000 0,24,36,44,52,60 2>&1 | grep -v "^\[bench\]\|Warning\|max persist" | tee -a $OUT/persist.txt
done
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
export SBWT_B200_LAYOUT=c64
echo "== c2q nch2r8mb3" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch2r8mb3.so q c2q 10000000
echo "== c4s nch2r8mb3" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch2r8mb3.so q c4s 10000000
echo "== c2 nch2r16mb2" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch2r16mb2.so q c2 10000000
