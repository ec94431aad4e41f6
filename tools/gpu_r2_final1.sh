#!/bin/bash
# Round 2, single-GPU closing pass: every GPU test, the default bench line (extras + full-length CPU proof), the reference arm on
# the box's host cores, the full synthetic sweep, verification, small circuits, circom witness, a launch list of one proof.
set -u
mkdir -p gpurun_out
t0=$SECONDS
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -12 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', r.get('frac') and round(r['frac'],3), 'L0 ms', r.get('launch_ms') and round(r['launch_ms'],3), 'traffic', r.get('traffic'), 'cpu', (d.get('cpu_baseline') or {}).get('seconds_per_proof'), (d.get('cpu_baseline') or {}).get('proof_bytes_equal_gpu'), 'pipelined', (d.get('pipelined') or {}).get('ms_per_proof'))
print('   stage', {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})
ex=d.get('extra') or {}
for k,v in ex.items():
    if k=='sweep' and v and 'records' in v:
        for rec in v['records']: print('   sweep', {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in rec.items() if kk in ('what','log_n','n_gpus','ms','Mpts_per_s','frac','checked','error')})
    elif v: print('   extra', k, {kk: v.get(kk) for kk in ('ms_per_step','error','proof_verified_in_exponent')}, 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
"; }
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$? $((SECONDS-t0))s"; tail -3 gpurun_out/bench_n1.log | cut -c1-250; show gpurun_out/bench_n1.json n1
G16_LIB=$PWD/gpurun_variants/libg16_lb6.so timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_lb6.json 2> gpurun_out/bench_lb6.log; show gpurun_out/bench_lb6.json k_ba_add_80regs_6blocks
timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_lb5.json 2> gpurun_out/bench_lb5.log; show gpurun_out/bench_lb5.json k_ba_add_92regs_5blocks
timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 --witness circom > gpurun_out/bench_circom.json 2> gpurun_out/bench_circom.log; show gpurun_out/bench_circom.json circom
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.log; echo "reference rc=$? $((SECONDS-t0))s"; tail -2 gpurun_out/bench_reference.log; cut -c1-400 gpurun_out/bench_reference.json
timeout 900 python tools/sweep.py > gpurun_out/sweep_n1.jsonl 2> gpurun_out/sweep_n1.log; echo "sweep rc=$? $((SECONDS-t0))s"; python -c "
import json
for l in open('gpurun_out/sweep_n1.jsonl'):
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('what','log_n','ms','Mpts_per_s','frac','checked','precompute','error')})"
tail -2 gpurun_out/sweep_n1.log
for w in 'S-2^12' 'S-2^16'; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.log; show gpurun_out/bench_$w.json $w
done
timeout 300 python bench.py --what verify --verify-sizes 1 65536 > gpurun_out/verify_bench.jsonl 2> gpurun_out/verify_bench.err; echo "verify rc=$? $((SECONDS-t0))s"; cut -c1-500 gpurun_out/verify_bench.jsonl
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_prove.csv \
  python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/prof_prove.log 2>&1; echo "ncu list rc=$? $((SECONDS-t0))s"
python tools/agg_launches.py gpurun_out/launches_prove.csv 2>/dev/null | head -28
du -sh gpurun_out
