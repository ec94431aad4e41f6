#!/bin/bash
# Round 2, third GPU pass: restructured k_ba_add (loads issued together, next references prefetched) against its variants,
# batched NTT, inlined reduction / windowed scaling kernels; window c = 20; shard timings.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_verify.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"
tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', r.get('frac') and round(r['frac'],3), 'L0 ms', r.get('launch_ms') and round(r['launch_ms'],3), 'graph', d.get('graph'), {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})"; }
for tag in default v0 free mul2 mul2free; do
  lib=$PWD/crescent_credentials_b200/libg16b200.so; [ $tag != default ] && lib=$PWD/crescent_credentials_b200/libg16b200_$tag.so
  [ -f $lib ] || { echo "missing $lib"; continue; }
  G16_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log
  echo "bench $tag rc=$? $((SECONDS-t0))s"; show gpurun_out/bench_$tag.json $tag
done
timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 --window-bits 20 > gpurun_out/bench_c20.json 2> gpurun_out/bench_c20.log; show gpurun_out/bench_c20.json c20
timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 --opt ntt_batch=0 > gpurun_out/bench_nobatch.json 2> gpurun_out/bench_nobatch.log; show gpurun_out/bench_nobatch.json ntt_batch0
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_conn32.json 2> gpurun_out/bench_conn32.log; show gpurun_out/bench_conn32.json conn32
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/prof_shard.py --world 8 --rank 1 --serialize 0 > gpurun_out/prof_shard8_rank1_conn32.log 2>&1; grep ms_per_shard_run gpurun_out/prof_shard8_rank1_conn32.log | cut -c1-520
for w in 'S-2^12' 'S-2^16'; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.log; show gpurun_out/bench_$w.json $w
done
for rk in 1 0; do
  timeout 300 python tools/prof_shard.py --world 8 --rank $rk > gpurun_out/prof_shard8_rank$rk.log 2>&1; echo "shard rank $rk rc=$? $((SECONDS-t0))s"
  grep ms_per_shard_run gpurun_out/prof_shard8_rank$rk.log | cut -c1-520
done
timeout 300 python tools/prof_shard.py --world 8 --rank 0 --opt ntt_batch=0 > gpurun_out/prof_shard8_rank0_nobatch.log 2>&1; grep ms_per_shard_run gpurun_out/prof_shard8_rank0_nobatch.log | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard8_rank1.csv \
  python tools/prof_shard.py --world 8 --rank 1 > gpurun_out/prof_shard8_rank1_ncu.log 2>&1; echo "ncu shard list rc=$? $((SECONDS-t0))s"
python tools/agg_launches.py gpurun_out/launches_shard8_rank1.csv 2>/dev/null | head -16
cap() {  # tag regex skip count
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/ncu_$1 \
    python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$? $((SECONDS-t0))s"
  python tools/ncu_summary.py gpurun_out/ncu_$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
}
cap ba_add_h '^k_ba_add$' 20 1
cap ba_products_h '^k_ba_products$' 20 1
du -sh gpurun_out
