#!/bin/bash
# usage: bash tools/gpu_quick.sh <tag> [skip-tests]
set -u
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
fi
for wl in c2 c3; do
  timeout 300 python tools/quick_walk.py $wl 2000000 2>&1 | tee -a $OUT/quick.txt
done
SBWT_B200_STORE_STREAMING=0 timeout 300 python tools/quick_walk.py c2 2000000 2>&1 | tee -a $OUT/quick.txt
timeout 300 python tools/quick_walk.py c2 2000000 12 2>&1 | tee -a $OUT/quick.txt
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
for wl in c2 c3; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:walk_kernel -s 3 -c 1 --csv --log-file $OUT/ncu_$wl.csv python tools/quick_walk.py $wl 2000000 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("$OUT/ncu_$wl.csv") if l.startswith('"'))]
print("$wl", {r[-3]: r[-1] for r in rows[1:]})
PY
done
