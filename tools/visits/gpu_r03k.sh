#!/bin/bash
# r03k: FASTA parse rate of the batch reader alone (old / new: scratch kept across batches, header flags noted by the scan, learnt window) and the CLI figure
set -u
TAG=${1:-r03k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python - <<'PY'
import numpy as np
rng=np.random.default_rng(1)
n=10_000_000; L=150
m=np.empty((n,L+6),dtype=np.uint8)
m[:,0]=ord('>'); m[:,1:4]=ord('r'); m[:,4]=ord('\n'); m[:,5:5+L]=np.frombuffer(b"ACGT",np.uint8)[rng.integers(0,4,size=(n,L))]; m[:,5+L]=ord('\n')
m.tofile('/tmp/reads10m.fna')
PY
for t in 1 4 8 16; do for v in old new; do echo -n "$v "; tools/probes/parse_time_$v /tmp/reads10m.fna $t; done; done 2>&1 | tee $OUT/parse.txt
timeout 900 python bench.py --no-e2e --no-cpu --no-probe --no-parity --legs none --steps 3 > $OUT/bench_cli.json 2> $OUT/bench_cli.log
python - <<PY | tee -a $OUT/parse.txt
import json
d=json.loads([l for l in open("$OUT/bench_cli.json") if l.startswith("{")][0]); print("cli_e2e", json.dumps(d.get("cli_e2e")))
PY
