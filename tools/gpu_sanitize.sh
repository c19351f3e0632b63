#!/bin/bash
# compute-sanitizer over the walk kernel's warp-level queues and staged stores: memcheck, racecheck and
# synccheck on the golden fixtures (+ part of config 1 under memcheck) in both modes, probe strides 0 / 1 / 9, both csector formats.
set -u
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for probe in 0 1 9; do
    for lay in c64 c96; do
      [ "$tool" != memcheck ] && [ "$lay" = c96 ] && [ "$probe" != 9 ] && continue
      extra=""; [ "$tool" = memcheck ] && [ "$probe" = 9 ] && [ "$lay" = c64 ] && extra="--with-c1"
      log=$OUT/${tool}_probe${probe}_${lay}.log
      SBWT_B200_PROBE=$probe SBWT_B200_LAYOUT=$lay SBWT_B200_COMPACT=2 timeout 900 $CS --tool $tool --error-exitcode 9 python tools/sanitize_run.py $extra > $log 2>&1
      echo "$tool probe=$probe layout=$lay rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    done
  done
done
