#!/bin/bash
set -u
TAG=${1:-w2b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for wl in c2 c3; do
  for tp in 10 11 12; do
    timeout 300 python tools/quick_walk.py $wl 2000000 $tp 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
for lib in .variants/*.so; do
  [ -f "$lib" ] || continue
  for wl in c2 c3; do
    echo "== $lib" | tee -a $OUT/quick.txt
    SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
  done
done
for b in 3 4 5; do
  SBWT_B200_BLOCKS_PER_SM=$b timeout 300 python tools/quick_walk.py c2 2000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
done
for wl in c2 c3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o $OUT/walk2_$wl \
    python tools/quick_walk.py $wl 2000000 > $OUT/ncu_full_$wl.log 2>&1; echo "ncu full $wl rc=$?"
done
