#!/bin/bash
# Round 2, second GPU pass: validate the safegcd inversion + CUDA-graph replay build, A/B graph on/off (full size and small
# circuits), shard-geometry launch lists (N = 8 ranks 1 and 0), three single-launch ncu --set full captures.
# Keeps gpurun_out small (< 64 MiB comes back).
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q -x -k "not mdl1" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"
tail -5 gpurun_out/pytest_gpu.log
for g in 1 0; do
  timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 --opt graph=$g > gpurun_out/bench_graph$g.json 2> gpurun_out/bench_graph$g.log
  echo "bench graph=$g rc=$? $((SECONDS-t0))s"; python -c "
import json; d=json.loads(open('gpurun_out/bench_graph$g.json').read().strip().splitlines()[-1]); print('graph=$g', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'roof', d['roofline'] and round(d['roofline']['frac'],3), d['roofline'] and round(d['roofline']['launch_ms'],3))"
  for w in 'S-2^12' 'S-2^16'; do
    timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --extras '' --no-cpu-baseline --inflight 0 --opt graph=$g > gpurun_out/bench_${w}_graph$g.json 2> gpurun_out/bench_${w}_graph$g.log
    python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_graph$g.json').read().strip().splitlines()[-1]); print('$w graph=$g', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['stage_ms'])"
  done
done
for rk in 1 0; do
  timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard8_rank$rk.csv \
    python tools/prof_shard.py --world 8 --rank $rk > gpurun_out/prof_shard8_rank$rk.log 2>&1; echo "ncu shard list rank $rk rc=$? $((SECONDS-t0))s"
  grep ms_per_shard_run gpurun_out/prof_shard8_rank$rk.log | cut -c1-420
  python tools/agg_launches.py gpurun_out/launches_shard8_rank$rk.csv 2>/dev/null | head -25
done
# single-launch ncu --set full captures (serialised proof: k_ba_add launch order a, l, b_g1 (5 each), b_g2 (5), h (5))
cap() {  # tag regex skip count
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/ncu_$1 \
    python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$? $((SECONDS-t0))s"
  python tools/ncu_summary.py gpurun_out/ncu_$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
}
cap ba_add_h '^k_ba_add$' 20 2
cap ba_add_g2 '^k_ba_add$' 15 2
cap ntt4 'k_ntt_pass4' 0 2
cap reduce 'k_chunk_reduce|k_ba_invert' 0 2
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
