#!/bin/bash
# r02z: the final library: ncu --set full + DRAM traffic per workload (for profiles/traffic.json), then the default bench run + reference arm
set -u
TAG=${1:-r02z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_final python tools/quick_walk.py c2 10000000 > $OUT/ncu_full_c2.log 2>&1; echo "ncu full rc=$?"
for wl in c2 c3 c4s c5s; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:walk -s 3 -c 1 --csv --log-file $OUT/traffic_$wl.csv \
      python tools/quick_walk.py $wl 10000000 > $OUT/traffic_$wl.log 2>&1; echo "traffic $wl rc=$?"
done
python tools/update_traffic.py $OUT "profiles/${TAG}_dram_traffic.txt"
bash tools/visits/gpu_r02j.sh $TAG
