#!/bin/bash
set -u
TAG=${1:-exp3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python tools/exp_knobs.py c2 10000000 l2 2>&1 | grep -v "^\[bench\]" | tee $OUT/knobs_c2.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:walk2_kernel --csv --log-file $OUT/knobs_c2_ncu.csv \
   python tools/exp_knobs.py c2 10000000 l2 > $OUT/knobs_ncu.log 2>&1; echo "ncu rc=$?"
g++ -O3 -march=native -pthread -o /tmp/host_bw_probe tools/host_bw_probe.cpp && /tmp/host_bw_probe 400000000 > $OUT/host_bw.txt 2>&1
cat $OUT/host_bw.txt; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ; free -g | head -2
