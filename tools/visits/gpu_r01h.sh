#!/bin/bash
# r01h: compact one-hot layout bring-up: parity tests, then A/B (SBWT_B200_COMPACT=0/1) walk timings on c2 / c3 at bench
# size with the DRAM traffic of the same launches.
set -u
TAG=${1:-r01h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
for wl in c2 c3; do
  for m in 0 1; do
    SBWT_B200_COMPACT=$m timeout 300 python tools/quick_walk.py $wl 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
  done
done
echo "t=$(( $(date +%s) - T0 ))s"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum
for m in 0 1; do
  SBWT_B200_COMPACT=$m timeout 600 ncu --metrics $M --clock-control none -k regex:walk2_kernel --csv --log-file $OUT/ncu_c2_compact$m.csv \
     python tools/quick_walk.py c2 10000000 > $OUT/ncu_c2_compact$m.log 2>&1
  echo "compact=$m" | tee -a $OUT/ncu_compact.txt
  python tools/ncu_table.py $OUT/ncu_c2_compact$m.csv 1 | tail -2 | tee -a $OUT/ncu_compact.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
