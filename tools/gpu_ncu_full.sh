#!/bin/bash
# usage: bash tools/gpu_ncu_full.sh <tag> <workload> [reads] : one full ncu capture of the walk kernel
set -u
TAG=${1:-p}; WL=${2:-c2}; R=${3:-2000000}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 3 -c 1 -f -o $OUT/walk_$WL \
    python tools/quick_walk.py $WL $R > $OUT/ncu_full_$WL.log 2>&1; echo "ncu full $WL rc=$?"
ls -la $OUT
