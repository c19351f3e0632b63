#!/bin/bash
# GPU visit r01f: parity tests, full-size c2 bench line (walk2 with staged sector stores; 32-bit result wire format on the
# host path), reference arm, ncu launch list + full capture of the walk kernel.
set -u
TAG=${1:-r01f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1; nproc > $OUT/nproc.txt; free -g > $OUT/free.txt
T0=$(date +%s)
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.log; echo "bench ref rc=$?"; cat $OUT/bench_ref_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 4 -c 1 -f -o $OUT/walk2_c2_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
echo "t=$(( $(date +%s) - T0 ))s"
ls -la $OUT
