#!/bin/bash
# One GPU-box pass over HEAD: GPU parity tests, smoke, the option sweep, the N=1 bench line, and the ncu launch list of one
# bench step (share-of-step evidence for the roofline kernel).  Usage (via gpurun, from the repo root): bash tools/gpu_pass1.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
t0=$SECONDS
timeout 400 python tools/sched_sweep.py --combos "split_chains=0,ntt_radix4=0,spmv_sell=0;split_chains=1,ntt_radix4=0,spmv_sell=0;split_chains=1,ntt_radix4=1,spmv_sell=0;split_chains=1,ntt_radix4=1,spmv_sell=1;split_chains=1,ntt_radix4=1,spmv_sell=1,wm_priority=1;split_chains=1,ntt_radix4=1,spmv_sell=1,serialize=1" > gpurun_out/sched_sweep.jsonl 2> gpurun_out/sched_sweep.err; echo "sweep rc=$? ($((SECONDS-t0)) s)"
cut -c1-400 gpurun_out/sched_sweep.jsonl; tail -3 gpurun_out/sched_sweep.err
t0=$SECONDS
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$? ($((SECONDS-t0)) s)"; cut -c1-1500 gpurun_out/bench_n1.json
t0=$SECONDS
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_prove.csv \
  python tools/prof_prove.py --precompute 1 --serialize 0 --reps 1 > gpurun_out/prof_prove.log 2>&1; echo "ncu rc=$? ($((SECONDS-t0)) s)"
wc -l gpurun_out/launches_prove.csv
