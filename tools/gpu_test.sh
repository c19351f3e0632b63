#!/bin/bash
# usage: bash tools/gpu_test.sh <tag> [pytest args]: the GPU parity tests only
set -u
TAG=${1:-t}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/pytest_gpu.log
