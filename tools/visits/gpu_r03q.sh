#!/bin/bash
# r03q: command line with the batches staged in page-locked memory by the parse-ahead thread: CLI tests, then the CLI figure
set -u
TAG=${1:-r03q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_text.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_cli.log
timeout 900 python bench.py --no-e2e --no-cpu --no-probe --no-parity --legs none --steps 3 > $OUT/bench_cli.json 2> $OUT/bench_cli.log
python - <<PY | tee $OUT/cli.txt
import json
d=json.loads([l for l in open("$OUT/bench_cli.json") if l.startswith("{")][0]); print("cli_e2e", json.dumps(d.get("cli_e2e")))
PY
