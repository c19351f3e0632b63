#!/bin/bash
# r02t: ncu --set full of the default library on c2 (16-character table), launch list of one bench step
set -u
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_tp16 python tools/quick_walk.py c2 10000000 > $OUT/ncu_full_c2.log 2>&1; echo "ncu full rc=$?"; tail -1 $OUT/ncu_full_c2.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe --no-parity --no-cli --legs none > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for wl in c2 c3 c4s c5s; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:walk -s 3 -c 1 --csv --log-file $OUT/traffic_$wl.csv \
      python tools/quick_walk.py $wl 10000000 > $OUT/traffic_$wl.log 2>&1; echo "traffic $wl rc=$?"
done
