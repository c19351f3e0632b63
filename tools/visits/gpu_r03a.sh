#!/bin/bash
# r03a: host-side passes removed from the host-buffer calls (chunk cut by bisection, one pass per chunk), robust clock sampling: tests + default bench + reference arm
set -u
TAG=${1:-r03a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
bash tools/visits/gpu_r02j.sh $TAG
