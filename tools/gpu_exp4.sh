#!/bin/bash
set -u
TAG=${1:-exp4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/exp_knobs.py c2 10000000 ${2:-quick} 2>&1 | grep -v "^\[bench\]" | tee $OUT/knobs_c2.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum -k regex:walk2_kernel --csv --log-file $OUT/knobs_c2_ncu.csv \
   python tools/exp_knobs.py c2 10000000 ${2:-quick} > $OUT/knobs_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_table.py $OUT/knobs_c2_ncu.csv 3
for wl in c3 c4s c5s; do timeout 600 python tools/exp_knobs.py $wl 10000000 base 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/knobs_other.txt; done
