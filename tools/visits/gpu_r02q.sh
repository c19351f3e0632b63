#!/bin/bash
# r02q: search-table length on the final kernel (L2 set-aside on): does a table that fits in L2 beat the 2.1 GB one now?
set -u
TAG=${1:-r02q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for tp in 14 13 12 11 10 9; do echo "== c2 tp=$tp" | tee -a $OUT/quick.txt; q c2 10000000 $tp; done
for tp in 12 10; do echo "== c4s tp=$tp" | tee -a $OUT/quick.txt; q c4s 10000000 $tp; done
