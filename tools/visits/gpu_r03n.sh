#!/bin/bash
# r03n: search-mode / LITERAL chain loop with the code word loaded behind the sector load (as in the streaming CHAIN), against the r03i library
set -u
TAG=${1:-r03n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c3 c2; do
  echo "== $wl r03i" | tee -a $OUT/quick.txt; SBWT_B200_LIB=$PWD/.variants/r03i.so q $wl 10000000
  echo "== $wl new" | tee -a $OUT/quick.txt; q $wl 10000000
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
