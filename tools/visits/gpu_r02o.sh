#!/bin/bash
# r02o: per-kernel durations of the multi-pass streaming walk on c2 / c2q (ncu launch list)
set -u
TAG=${1:-r02o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in c2 c2q; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"pass_kernel|walk_kernel" -s 24 -c 11 --csv --log-file $OUT/passes_$wl.csv \
    python tools/quick_walk.py $wl 10000000 > $OUT/passes_$wl.log 2>&1; echo "ncu rc=$?"
done
