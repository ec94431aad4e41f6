#!/bin/bash
# End-of-round GPU pass: all GPU tests, smoke, the N=1 bench line, the verification sweep with its CPU baseline, the small-circuit
# latency line and a fresh ncu launch list of one serialised proof.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 ))s"
tail -9 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-400 gpurun_out/bench_n1.json
timeout 300 python bench.py --what verify > gpurun_out/verify_bench.jsonl 2> gpurun_out/verify_bench.err; echo "verify rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-700 gpurun_out/verify_bench.jsonl; tail -3 gpurun_out/verify_bench.err
for w in S-2^12 S-2^16; do
  timeout 200 python bench.py --workload $w --no-cpu-baseline --inflight 0 --steps 50 --extras "" > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.log; echo "$w rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
done
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_prove.csv \
  python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/prof_prove.log 2>&1; echo "ncu list rc=$? $(( $(date +%s) - t0 ))s"
python tools/agg_launches.py gpurun_out/launches_prove.csv 2>/dev/null | head -30
