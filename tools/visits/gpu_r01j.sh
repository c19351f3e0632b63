#!/bin/bash
# GPU visit r01j: parity tests, bench lines of all four single-GPU workloads at full size (c2 = the headline config, compact
# one-hot layout), reference arm, ncu launch list + full capture of the walk kernel, CLI end-to-end timing.
set -u
TAG=${1:-r01j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.log; echo "bench c2 rc=$?"; cat $OUT/bench_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.log; echo "bench ref rc=$?"; cat $OUT/bench_ref_c2.json
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 4 -c 1 -f -o $OUT/walk2_c2_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe > $OUT/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
echo "t=$(( $(date +%s) - T0 ))s"
# the command line, end to end: 2M reads of the c2 workload as FASTA -> text output
python - > $OUT/cli.txt 2>&1 <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, ".")
import bench
from sbwt_b200.testing import synth
w = bench.WORKLOADS["c2"]
path, ref = bench.ensure_index("c2", w)
reads = synth.sample_reads(ref, 2_000_000, 150, 0.5, seed=43)
q = os.path.join(bench.CACHE, "cli_q.fna")
bench.write_sample_fasta(q, reads)
for th in (1, 8, 16):
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run(["sbwt_b200/csrc/sbwt_search", "search", "-i", path, "-q", q, "-o", os.path.join(bench.CACHE, "cli_out.txt"), "--threads", str(th)],
                           capture_output=True, text=True)
        dt = time.time() - t0
        us = [l for l in r.stderr.splitlines() + r.stdout.splitlines() if "us/query" in l]
        print(f"threads={th} rc={r.returncode} wall={dt:.2f}s  {240e6 / dt / 1e6:.1f} M lookups/s wall (index load included)  {us[-1] if us else ''}", flush=True)
print("output bytes", os.path.getsize(os.path.join(bench.CACHE, "cli_out.txt")))
PY
cat $OUT/cli.txt
echo "t=$(( $(date +%s) - T0 ))s"
for wl in c3 c4s c5s; do
  timeout 1200 python bench.py --workload $wl > $OUT/bench_$wl.json 2> $OUT/bench_$wl.log; echo "bench $wl rc=$?"; cat $OUT/bench_$wl.json
  echo "t=$(( $(date +%s) - T0 ))s"
done
ls -la $OUT
