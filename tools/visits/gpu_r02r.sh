#!/bin/bash
# r02r: search tables of 15 and 16 characters (8.6 / 34 GB of rows in HBM)
set -u
TAG=${1:-r02r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for tp in 15 16; do echo "== c2 tp=$tp" | tee -a $OUT/quick.txt; q c2 10000000 $tp; done
for tp in 14 15 16; do echo "== c4s tp=$tp" | tee -a $OUT/quick.txt; q c4s 10000000 $tp; done
for tp in 15; do echo "== c5s tp=$tp" | tee -a $OUT/quick.txt; q c5s 10000000 $tp; echo "== c3 tp=$tp" | tee -a $OUT/quick.txt; q c3 10000000 $tp; done
