#!/bin/bash
# r01r: DRAM traffic of one walk launch per workload at the final defaults (for profiles/traffic.json)
set -u
TAG=${1:-r01r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for wl in c3 c4s c5s; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:walk2_kernel --csv --log-file $OUT/ncu_$wl.csv python tools/quick_walk.py $wl 10000000 > $OUT/ncu_$wl.log 2>&1
  echo $wl | tee -a $OUT/ncu.txt; python tools/ncu_table.py $OUT/ncu_$wl.csv 1 | tail -1 | tee -a $OUT/ncu.txt
done
