#!/bin/bash
# r03o: traffic / L1TEX requests per workload on the final sources (for profiles/traffic.json), then the default bench run + reference arm
set -u
TAG=${1:-r03o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in c2 c3 c4s c5s; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum -k regex:walk -s 3 -c 1 --csv --log-file $OUT/traffic_$wl.csv \
      python tools/quick_walk.py $wl 10000000 > $OUT/traffic_$wl.log 2>&1; echo "traffic $wl rc=$?"
done
python tools/update_traffic.py $OUT "profiles/${TAG}_dram_traffic.txt"
bash tools/visits/gpu_r02j.sh $TAG
