#!/bin/bash
# Experiment pass: parity tests, then the option sweep (stream plan / radix-4 NTT / sliced-ELL SpMV) and the NTT A/B.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --durations=8 > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_parity.log
timeout 600 python tools/sched_sweep.py --combos "split_chains=0,ntt_radix4=0,spmv_sell=0;split_chains=1,ntt_radix4=0,spmv_sell=0;split_chains=1,ntt_radix4=1,spmv_sell=0;split_chains=1,ntt_radix4=1,spmv_sell=1;split_chains=1,ntt_radix4=1,spmv_sell=1,wm_priority=1;split_chains=0,ntt_radix4=1,spmv_sell=1,wm_priority=1;split_chains=1,ntt_radix4=1,spmv_sell=1,serialize=1" > gpurun_out/sched_sweep.jsonl 2> gpurun_out/sched_sweep.err; echo "sweep rc=$?"
cat gpurun_out/sched_sweep.jsonl; tail -3 gpurun_out/sched_sweep.err
timeout 300 python tools/msm_bench.py --ntt 1 --logn 21 22 --opt ntt_radix4=0 > gpurun_out/ntt_r2.jsonl 2>gpurun_out/ntt.err
timeout 300 python tools/msm_bench.py --ntt 1 --logn 21 22 --opt ntt_radix4=1 > gpurun_out/ntt_r4.jsonl 2>>gpurun_out/ntt.err
cat gpurun_out/ntt_r2.jsonl gpurun_out/ntt_r4.jsonl
