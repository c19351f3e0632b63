#!/bin/bash
set -u
TAG=${1:-exp5}; SET=${2:-frac}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python tools/exp_knobs.py c2 10000000 $SET 2>&1 | grep -v "^\[bench\]" | tee $OUT/knobs_c2.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:walk2_kernel --csv --log-file $OUT/knobs_c2_ncu.csv \
   python tools/exp_knobs.py c2 10000000 $SET > $OUT/knobs_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_table.py $OUT/knobs_c2_ncu.csv 3
timeout 600 python tools/exp_knobs.py c3 10000000 base 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/knobs_other.txt
