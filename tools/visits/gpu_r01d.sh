#!/bin/bash
# GPU visit r01d: parity tests, full-size c2/c3 bench lines (walk2 + PROBE pass), reference arm, ncu launch list + full capture.
set -u
TAG=${1:-r01d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
timeout 600 python bench.py --workload c3 --no-cpu > $OUT/bench_c3.json 2> $OUT/bench_c3.log; echo "bench c3 rc=$?"; cat $OUT/bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 4 -c 1 -f -o $OUT/walk2_c2_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
ls -la $OUT
