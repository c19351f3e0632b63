#!/bin/bash
# r02m: full-size configs[3] / [4] (2 Gbp +RC) with the round-2 kernel; DRAM traffic of one walk launch per workload (ncu) for
# profiles/traffic.json; ncu --set full + launch list of the default binary on c2; csector format on an index between 60 and 120 MB (c2m)
set -u
TAG=${1:-r02m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
sha256sum sbwt_b200/libsbwt_b200.so | cut -c1-12 > $OUT/lib_sha.txt
for wl in c2 c3 c4s c5s; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:walk -s 3 -c 1 --csv --log-file $OUT/traffic_$wl.csv \
      python tools/quick_walk.py $wl 10000000 > $OUT/traffic_$wl.log 2>&1; echo "traffic $wl rc=$?"
done
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 3 -c 1 -f -o $OUT/walk_c2_final python tools/quick_walk.py c2 10000000 > $OUT/ncu_full_c2.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe --no-parity --no-cli --legs none > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
echo "t=$(( $(date +%s) - T0 ))s"
for lay in c64 c96; do
  echo "== c2m $lay" | tee -a $OUT/quick.txt
  SBWT_B200_LAYOUT=$lay timeout 300 python tools/quick_walk.py c2m 10000000 2>&1 | grep -v "^\[bench\]" | tee -a $OUT/quick.txt
done
echo "t=$(( $(date +%s) - T0 ))s"
AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo); echo "MemAvailable ${AVAIL} GB"
timeout 1500 python bench.py --workload c4 --steps 10 --warmup 3 --no-cli --quick-cpu > $OUT/bench_c4.json 2> $OUT/bench_c4.log; echo "bench c4 rc=$?"; tail -3 $OUT/bench_c4.log | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
rm -f .cache/bench/pangenome_400_*k31*
[ "$AVAIL" -ge 150 ] && { timeout 1800 python bench.py --workload c5 --steps 10 --warmup 3 --no-cli --quick-cpu > $OUT/bench_c5.json 2> $OUT/bench_c5.log; echo "bench c5 rc=$?"; tail -3 $OUT/bench_c5.log | cut -c1-300; }
echo "t=$(( $(date +%s) - T0 ))s"
