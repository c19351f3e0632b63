#!/bin/bash
# r01n: search-table length sweep on every workload (the table is runtime-only; results do not depend on it)
set -u
TAG=${1:-r01n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in c2 c3 c4s c5s; do
  timeout 600 python tools/exp_knobs.py $wl 10000000 tp 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/knobs_tp.txt
done
