#!/bin/bash
# full-size bench lines for the four single-GPU workloads. usage: bash tools/gpu_bench4.sh <tag>
set -u
TAG=${1:-b4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for wl in c2 c3 c4s c5s; do
  extra=""; [ $wl = c3 ] && extra="--no-cpu"
  timeout 1200 python bench.py --workload $wl $extra > $OUT/bench_$wl.json 2> $OUT/bench_$wl.log; echo "bench $wl rc=$?"; cat $OUT/bench_$wl.json
done
