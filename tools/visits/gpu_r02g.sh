#!/bin/bash
# r02g: ncu --set full of the staged-output walk kernel (3 blocks/SM build) on c2 at full size
set -u
TAG=${1:-r02g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export SBWT_B200_LAYOUT=c64
SBWT_B200_LIB=$PWD/.variants/nch1mb3.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_nch1mb3 \
    python tools/quick_walk.py c2 10000000 > $OUT/ncu_c2.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_c2.log
