#!/bin/bash
# Multi-GPU bench pass: bash tools/gpu_multi.sh N "extra bench args" tag
N=$1; EXTRA=${2:-}; TAG=${3:-run}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.log
echo "rc=$?"; tail -c 3000 gpurun_out/bench_n${N}_${TAG}.json; tail -3 gpurun_out/bench_n${N}_${TAG}.log
