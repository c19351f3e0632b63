#!/bin/bash
# r01p: sparse wire format with the vectorised mixed-group expansion: parity + host pipeline sweep
set -u
TAG=${1:-r01p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 python tools/exp_e2e.py c2 10000000 wire 2>&1 | grep -v "^\[bench\]" | tee $OUT/e2e_wire.txt
echo "t=$(( $(date +%s) - T0 ))s"
