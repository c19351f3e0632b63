#!/bin/bash
# r01l (2 GPUs): the contract's multi-GPU launch of bench.py (torchrun, one rank per GPU), both arms.
set -u
TAG=${1:-r01l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
nvidia-smi -L > $OUT/gpus.txt 2>&1; nproc >> $OUT/gpus.txt; free -g >> $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    > $OUT/bench_c2_2gpu.json 2> $OUT/bench_c2_2gpu.log; echo "bench 2gpu rc=$?"; cat $OUT/bench_c2_2gpu.json; tail -3 $OUT/bench_c2_2gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 \
    > $OUT/bench_ref_2gpu.json 2> $OUT/bench_ref_2gpu.log; echo "ref 2gpu rc=$?"; cat $OUT/bench_ref_2gpu.json
echo "t=$(( $(date +%s) - T0 ))s"
