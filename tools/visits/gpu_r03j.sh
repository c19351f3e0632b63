#!/bin/bash
# r03j: search table built level by level (table_extend_kernel) against every row walked from scratch (r03i library): build times, then the GPU suite
set -u
TAG=${1:-r03j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in c2 c4s; do
  echo "== $wl from scratch (r03i)" | tee -a $OUT/load.txt; SBWT_B200_LIB=$PWD/.variants/r03i.so timeout 600 python tools/index_load_time.py $wl 16 14 12 10 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/load.txt
  echo "== $wl level by level" | tee -a $OUT/load.txt; timeout 600 python tools/index_load_time.py $wl 16 14 12 10 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/load.txt
done
timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/load.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
