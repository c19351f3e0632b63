#!/bin/bash
# r02y: coalesced packer + fused plan kernels (prep 0.85 ms before): tests, quick lines, launch list of one step
set -u
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p $OUT
q() { timeout 300 python tools/quick_walk.py "$@" 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt; }
for wl in c2 c3 c5s; do echo "== $wl" | tee -a $OUT/quick.txt; q $wl 10000000; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe --no-parity --no-cli --legs none > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
