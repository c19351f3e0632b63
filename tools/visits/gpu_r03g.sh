#!/bin/bash
# r03g: compute-sanitizer on the FINAL round-2 library (the CHAIN loop, packer, plan kernels, CASE_API and k > 64 kernels changed after r02k):
# memcheck over probe strides 0 / 9 and both csector formats, racecheck + synccheck on the default format
set -u
TAG=${1:-r03g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; probe=$2; lay=$3; extra=${4:-}; log=$OUT/${tool}_probe${probe}_${lay}.log
  SBWT_B200_PROBE=$probe SBWT_B200_LAYOUT=$lay SBWT_B200_COMPACT=2 timeout 900 $CS --tool $tool --error-exitcode 9 python tools/sanitize_run.py $extra > $log 2>&1
  echo "$tool probe=$probe layout=$lay rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"; }
run memcheck 9 c64 --with-c1
run memcheck 0 c64
run memcheck 9 c96
run racecheck 9 c64
run racecheck 0 c64
run synccheck 9 c64
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2
