#!/bin/bash
# GPU visit r01c: parity tests, full-size c2 bench line, ncu launch list + full capture of the walk kernel at bench size.
set -u
TAG=${1:-r01c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
(nproc; lscpu | head -25; free -g) > $OUT/host.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.log; echo "bench ref rc=$?"; cat $OUT/bench_ref_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 4 -c 1 -f -o $OUT/walk2_c2_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
ls -la $OUT
