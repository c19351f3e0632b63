#!/bin/bash
set -u
OUT=gpurun_out/exp1; mkdir -p $OUT
g++ -O3 -march=native -pthread -o /tmp/host_bw_probe tools/host_bw_probe.cpp && /tmp/host_bw_probe 400000000 > $OUT/host_bw.txt 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum
for wl in c2 c3; do
  timeout 600 python tools/exp_l2.py $wl 2000000 > $OUT/exp_l2_$wl.txt 2>&1
  timeout 900 ncu --metrics $M --clock-control none -k regex:walk_kernel --csv --log-file $OUT/exp_l2_${wl}_ncu.csv python tools/exp_l2.py $wl 2000000 > $OUT/exp_l2_${wl}_underncu.txt 2>&1
done
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
cat $OUT/host_bw.txt $OUT/exp_l2_c2.txt $OUT/exp_l2_c3.txt
