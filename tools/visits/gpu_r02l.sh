#!/bin/bash
# r02l: the contract's multi-GPU launch (torchrun, 2 ranks), both arms
set -u
TAG=${1:-r02l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.log; echo "bench rc=$?"; tail -5 $OUT/bench_2gpu.log | cut -c1-300; head -c 1500 $OUT/bench_2gpu.json
echo; echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_2gpu.json 2> $OUT/bench_ref_2gpu.log; echo "ref rc=$?"; cat $OUT/bench_ref_2gpu.json | head -c 600
echo; echo "t=$(( $(date +%s) - T0 ))s"
