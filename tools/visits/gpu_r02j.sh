#!/bin/bash
# r02j: the default bench run (all single-GPU workloads, parity against sbwt_ref, e2e legs, CPU baseline variants, CLI figure) + reference arm
set -u
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.log; echo "bench rc=$?"; tail -12 $OUT/bench.log; cat $OUT/bench.json | head -c 6000
echo; echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.log; echo "ref rc=$?"; cat $OUT/bench_ref.json
echo "t=$(( $(date +%s) - T0 ))s"
