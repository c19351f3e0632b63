#!/bin/bash
# r01g experiments: (1) random-sector ceiling vs buffer size and L2 fetch granularity (+ DRAM bytes per load under ncu),
# (2) walk kernel with the fetch-granularity knob on c2 / c4s, (3) host pipeline sweep.
set -u
TAG=${1:-exp6}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
for f in default 32 64; do
  if [ $f = default ]; then unset SBWT_B200_L2_FETCH; else export SBWT_B200_L2_FETCH=$f; fi
  timeout 300 python tools/exp_probe2.py 2>&1 | tee -a $OUT/probe.txt
done
for f in default 32; do
  if [ $f = default ]; then unset SBWT_B200_L2_FETCH; else export SBWT_B200_L2_FETCH=$f; fi
  timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_ltcfabric.sum \
     --clock-control none -k regex:probe_kernel --csv --log-file $OUT/probe_ncu_$f.csv python tools/exp_probe2.py 0 > $OUT/probe_underncu_$f.txt 2>&1
  python tools/ncu_table.py $OUT/probe_ncu_$f.csv 1 | tee -a $OUT/probe_ncu.txt
done
unset SBWT_B200_L2_FETCH
echo "t=$(( $(date +%s) - T0 ))s"
for wl in c2 c4s; do
  for f in default 32 64; do
    if [ $f = default ]; then unset SBWT_B200_L2_FETCH; else export SBWT_B200_L2_FETCH=$f; fi
    timeout 300 python tools/quick_walk.py $wl 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_fetch.txt
  done
done
unset SBWT_B200_L2_FETCH
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 python tools/exp_e2e.py c2 10000000 main 2>&1 | grep -v "^\[bench\]" | tee $OUT/e2e.txt
echo "t=$(( $(date +%s) - T0 ))s"
for lib in .variants/*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" | tee -a $OUT/variants.txt
  SBWT_B200_LIB=$PWD/$lib timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/variants.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
