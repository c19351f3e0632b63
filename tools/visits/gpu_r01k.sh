#!/bin/bash
# r01k: steady-state fast loop of CHAIN: parity + timings on all streaming workloads (+ instruction counts for c2).
set -u
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
SBWT_B200_COMPACT=1 timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk.txt
SBWT_B200_COMPACT=0 timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk.txt
timeout 300 python tools/quick_walk.py c4s 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk.txt
timeout 300 python tools/quick_walk.py c5s 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk.txt
echo "t=$(( $(date +%s) - T0 ))s"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for wl in c2 c4s; do
timeout 600 ncu --metrics $M --clock-control none -k regex:walk2_kernel --csv --log-file $OUT/ncu_$wl.csv python tools/quick_walk.py $wl 10000000 > $OUT/ncu_$wl.log 2>&1
echo $wl | tee -a $OUT/ncu.txt; python tools/ncu_table.py $OUT/ncu_$wl.csv 1 | tail -1 | tee -a $OUT/ncu.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
