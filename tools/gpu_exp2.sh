#!/bin/bash
set -u
OUT=gpurun_out/exp2; mkdir -p $OUT
python tools/exp_probe.py > $OUT/probe.txt 2>&1
timeout 900 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:probe_kernel --csv --log-file $OUT/probe_ncu.csv python tools/exp_probe.py > $OUT/probe_underncu.txt 2>&1
cat $OUT/probe.txt
