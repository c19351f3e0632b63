#!/bin/bash
# r03b: host threads of the dense-result rebuild (8 by default), e2e legs only
set -u
TAG=${1:-r03b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for t in 8 10 12 14 16; do
  SBWT_B200_WIDEN_THREADS=$t timeout 600 python bench.py --no-cpu --no-parity --no-cli --legs none > $OUT/bench_t$t.json 2> $OUT/bench_t$t.log
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_t$t.json") if l.startswith("{")][0]); e=d["e2e"]
print("widen_threads=$t dense %.2f G/s (%.1f ms) host_ceiling %s frac %.3f | int32 %.2f | hits %.2f | bitmap %.2f" % (e["value"]/1e9, e["ms_per_step"], e.get("host_ceiling",{}).get("values_per_s"), e.get("host_bw_frac",0), e["int32_results"]["value"]/1e9, e["hits_only"]["value"]/1e9, e["bitmap_only"]["value"]/1e9))
PY
done 2>&1 | tee $OUT/summary.txt
