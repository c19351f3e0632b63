#!/bin/bash
# r02a: what the GPU box offers (cores, RAM, NUMA, disk) and whether the FULL-SIZE configs[3]/[4] indexes (2 Gbp +RC, ~3.2 G columns)
# can be built on it with the in-repo constructor; bench lines of both with the round-1 kernel as the "before" numbers.
set -u
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
{
  nproc; free -g; lscpu | egrep 'Model name|Socket|NUMA|Thread|Core|L3'; df -h /tmp . ; nvidia-smi topo -m; nvidia-smi --query-gpu=name,memory.total --format=csv
  cat /proc/meminfo | head -5
} > $OUT/box.txt 2>&1
cat $OUT/box.txt
( while true; do free -g | sed -n 2p; sleep 5; done ) > $OUT/free.log 2>&1 &
FREEPID=$!
T0=$(date +%s)
AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo); echo "MemAvailable ${AVAIL} GB"
[ "$AVAIL" -ge 110 ] && timeout 1500 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e > $OUT/bench_c4.json 2> $OUT/bench_c4.log; echo "bench c4 rc=$?"; cat $OUT/bench_c4.log | tail -5; cat $OUT/bench_c4.json
echo "t=$(( $(date +%s) - T0 ))s"
[ "$AVAIL" -ge 220 ] && timeout 1500 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e > $OUT/bench_c5.json 2> $OUT/bench_c5.log; echo "bench c5 rc=$?"; cat $OUT/bench_c5.log | tail -5; cat $OUT/bench_c5.json
echo "t=$(( $(date +%s) - T0 ))s"
kill $FREEPID
sort -k3 -n -r $OUT/free.log | head -2
ls -la .cache/bench
