#!/bin/bash
# r02c: ncu --set full of the rewritten walk kernel (variant nch1mb4, csector64) on c2 at full size; c4s with the variants.
set -u
TAG=${1:-r02c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
R=10000000
export SBWT_B200_LAYOUT=c64
SBWT_B200_LIB=$PWD/.variants/nch1mb4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_nch1mb4 \
    python tools/quick_walk.py c2 $R > $OUT/ncu_c2.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_c2.log
SBWT_B200_LIB=$PWD/.variants/nch2mb3.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_nch2mb3 \
    python tools/quick_walk.py c2 $R > $OUT/ncu_c2b.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_c2b.log
for lib in .variants/nch1mb4.so .variants/nch2mb3.so; do
  echo "== $lib c4s" | tee -a $OUT/quick.txt
  SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py c4s $R 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
done
