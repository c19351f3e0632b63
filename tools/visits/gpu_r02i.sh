#!/bin/bash
# r02i: full GPU test suite (new: batched f-4 queries, hits-only results, > 2^31-column narrow layout, long reads on the
# invariant-violating index) + c2 / c4s timings of the default build
set -u
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
echo "== c2 default" | tee -a $OUT/quick.txt; q c2 10000000
echo "== c2 default out32" | tee -a $OUT/quick.txt; QUICK_OUT32=1 q c2 10000000
echo "== c4s default" | tee -a $OUT/quick.txt; q c4s 10000000
echo "t=$(( $(date +%s) - T0 ))s"
