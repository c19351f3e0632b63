#!/bin/bash
# r02e: where does the c2 walk spend its time? (a) an index that certainly fits in L2 (c2q), (b) int32 results, (c) no result
# stores at all, (d) fewer resident warps (is it latency-bound?).
set -u
TAG=${1:-r02e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
R=10000000
export SBWT_B200_LAYOUT=c64
V=$PWD/.variants
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
echo "== c2q nch1mb4 c64" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb4.so q c2q $R
echo "== c2q nch1mb4 c96" | tee -a $OUT/quick.txt; SBWT_B200_LAYOUT=c96 SBWT_B200_LIB=$V/nch1mb4.so q c2q $R
echo "== c2 nch1mb4 out32" | tee -a $OUT/quick.txt; QUICK_OUT32=1 SBWT_B200_LIB=$V/nch1mb4.so q c2 $R
echo "== c2 nch1mb4 nostore" | tee -a $OUT/quick.txt; QUICK_NOSTORE=1 SBWT_B200_LIB=$V/nch1mb4ns.so q c2 $R
echo "== c2 nch1mb3" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb3.so q c2 $R
echo "== c2 nch1mb2" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb2.so q c2 $R
echo "== c2q nch2mb3 c64" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch2mb3.so q c2q $R
echo "== c2 nch1mb4 tp=10" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb4.so q c2 $R 10
echo "== c2 nch1mb4 tp=12" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$V/nch1mb4.so q c2 $R 12
