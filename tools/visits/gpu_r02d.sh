#!/bin/bash
# r02d: does an L2 set-aside (cudaLimitPersistingL2CacheSize) keep the csectors resident? c2, both csector formats.
set -u
TAG=${1:-r02d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for lay in c96 c64; do
  SBWT_B200_LAYOUT=$lay SBWT_B200_LIB=$PWD/.variants/nch1mb4.so timeout 300 python tools/exp_persist.py c2 10000000 0,36,44,56,72,88 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/persist.txt
done
SBWT_B200_LAYOUT=c96 SBWT_B200_LIB=$PWD/.variants/nch2mb3.so timeout 300 python tools/exp_persist.py c2 10000000 0,44,56 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/persist.txt
