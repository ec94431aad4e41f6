#!/bin/bash
# Round 2, 2-GPU pass: the ShardedProver class over NCCL (pytest), N = 1 sanity line, N = 2 bench with extras (S-mdl1, sweep slice).
set -u
mkdir -p gpurun_out
t0=$SECONDS
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$? $((SECONDS-t0))s"; tail -4 gpurun_out/pytest_sharded.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', r.get('frac') and round(r['frac'],3), 'graph', d.get('graph'), 'rank_stage', d.get('rank_stage_ms'))
ex=d.get('extra') or {}
for k,v in ex.items():
    if k=='sweep' and v and 'records' in v:
        for rec in v['records']: print('   sweep', {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in rec.items() if kk in ('what','log_n','n_gpus','ms','Mpts_per_s','frac','checked','error')})
    else: print('   extra', k, json.dumps(v)[:400])
"; }
timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/bench_n1_quick.json 2> gpurun_out/bench_n1_quick.log; echo "n1 rc=$? $((SECONDS-t0))s"; show gpurun_out/bench_n1_quick.json n1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; echo "n2 rc=$? $((SECONDS-t0))s"
tail -3 gpurun_out/bench_n2.log | cut -c1-300; show gpurun_out/bench_n2.json n2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 tools/sweep.py --g1 18 22 --g2 18 22 --ntt 20 > gpurun_out/sweep_n2.jsonl 2> gpurun_out/sweep_n2.log; echo "sweep n2 rc=$? $((SECONDS-t0))s"; cut -c1-330 gpurun_out/sweep_n2.jsonl; tail -3 gpurun_out/sweep_n2.log
du -sh gpurun_out
