#!/bin/bash
# r01o: sparse result wire format + new table-length default: parity, host pipeline sweep (dense vs sparse), tp = 14, bench.
set -u
TAG=${1:-r01o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 python tools/exp_e2e.py c2 10000000 wire 2>&1 | grep -v "^\[bench\]" | tee $OUT/e2e_wire.txt
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 python tools/exp_knobs.py c2 10000000 tp14 2>&1 | grep -v "^\[bench\]" | tee $OUT/knobs_tp14.txt
timeout 600 python tools/exp_knobs.py c4s 10000000 tp14 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/knobs_tp14.txt
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
