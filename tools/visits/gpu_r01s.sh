#!/bin/bash
# r01s: 14-character search table as default: parity tests, walk timings with oracle spot parity on every workload, c2 bench line.
set -u
TAG=${1:-r01s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
for wl in c3 c4s c5s; do
  timeout 300 python tools/quick_walk.py $wl 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
