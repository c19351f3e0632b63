#!/bin/bash
# r01t: ncu launch list + full capture of the walk kernel for the final binary of round 1 (14-character table)
set -u
TAG=${1:-r01t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk -s 4 -c 1 -f -o $OUT/walk2_c2_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
