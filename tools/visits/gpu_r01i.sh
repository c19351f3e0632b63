#!/bin/bash
# r01i: compact layout, second pass (base load in flight with the csector, branchless counts): parity + A/B timings.
set -u
TAG=${1:-r01i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
SBWT_B200_COMPACT=0 timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
SBWT_B200_COMPACT=1 timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
SBWT_B200_COMPACT=1 SBWT_B200_DEBUG_NOSTORE=1 timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
SBWT_B200_COMPACT=1 SBWT_B200_LIB=$PWD/.variants/mb5.so timeout 300 python tools/quick_walk.py c2 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
SBWT_B200_COMPACT=1 SBWT_B200_COMPACT_SEARCH=1 timeout 300 python tools/quick_walk.py c3 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/walk_compact.txt
echo "t=$(( $(date +%s) - T0 ))s"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
SBWT_B200_COMPACT=1 timeout 600 ncu --metrics $M --clock-control none -k regex:walk2_kernel --csv --log-file $OUT/ncu_c2_compact1.csv \
     python tools/quick_walk.py c2 10000000 > $OUT/ncu_c2_compact1.log 2>&1
python tools/ncu_table.py $OUT/ncu_c2_compact1.csv 1 | tail -1 | tee -a $OUT/ncu_compact.txt
echo "t=$(( $(date +%s) - T0 ))s"
