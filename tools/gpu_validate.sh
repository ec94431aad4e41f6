#!/bin/bash
# One GPU-box pass: GPU parity tests, smoke, the N=1 bench line and (optionally) the synthetic sweep.
# Usage (from the repo root, via gpurun): bash tools/gpu_validate.sh [sweep]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
if [ "${1:-}" = "sweep" ]; then
  timeout 600 python tools/msm_bench.py --group 1 --precompute 1 --logn 16 18 20 22 24 > gpurun_out/sweep_g1_tables.jsonl 2>gpurun_out/sweep.err
  timeout 600 python tools/msm_bench.py --group 1 --precompute 0 --logn 16 18 20 22 24 > gpurun_out/sweep_g1_notables.jsonl 2>>gpurun_out/sweep.err
  timeout 600 python tools/msm_bench.py --group 2 --precompute 1 --logn 16 18 20 22 > gpurun_out/sweep_g2_tables.jsonl 2>>gpurun_out/sweep.err
  timeout 600 python tools/msm_bench.py --group 2 --precompute 0 --logn 16 18 20 22 > gpurun_out/sweep_g2_notables.jsonl 2>>gpurun_out/sweep.err
  timeout 600 python tools/msm_bench.py --ntt 1 --logn 16 18 20 22 24 26 > gpurun_out/sweep_ntt.jsonl 2>>gpurun_out/sweep.err
  tail -2 gpurun_out/sweep_*.jsonl
fi
