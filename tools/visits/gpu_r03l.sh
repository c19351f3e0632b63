#!/bin/bash
# r03l: confirmation of the final tree: GPU suite, smoke, default bench run + reference arm
set -u
TAG=${1:-r03l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
bash tools/visits/gpu_r02j.sh $TAG
