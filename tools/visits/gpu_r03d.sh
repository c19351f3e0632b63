#!/bin/bash
# r03d: full-size configs[3] / [4] (2 Gbp +RC: 2.56 G / 3.78 G columns) with the final round-2 library (16-character table where memory allows)
set -u
TAG=${1:-r03d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo); echo "MemAvailable ${AVAIL} GB"
timeout 1500 python bench.py --workload c4 --steps 10 --warmup 3 --no-cli --quick-cpu --legs none > $OUT/bench_c4.json 2> $OUT/bench_c4.log; echo "bench c4 rc=$?"; tail -3 $OUT/bench_c4.log | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
rm -f .cache/bench/pangenome_400_*k31*
[ "$AVAIL" -ge 150 ] && { timeout 1800 python bench.py --workload c5 --steps 10 --warmup 3 --no-cli --quick-cpu --legs none > $OUT/bench_c5.json 2> $OUT/bench_c5.log; echo "bench c5 rc=$?"; tail -3 $OUT/bench_c5.log | cut -c1-300; }
echo "t=$(( $(date +%s) - T0 ))s"
