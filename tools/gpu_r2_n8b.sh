#!/bin/bash
# Round 2, second 8-GPU pass: the witness map split over ranks 0 and 1 (plan "wm_split") -- NCCL class test on 4 ranks, A/B at N = 8,
# then the full N = 8 line with extras.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$? $((SECONDS-t0))s"; tail -4 gpurun_out/pytest_sharded.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'), 'plan', d['config'].get('plan'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
ex=d.get('extra') or {}
for k,v in ex.items():
    if k=='sweep' and v and 'records' in v:
        for rec in v['records']: print('   sweep', {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in rec.items() if kk in ('what','log_n','n_gpus','ms','Mpts_per_s','frac','checked','error')})
    elif v: print('   extra', k, {kk: v.get(kk) for kk in ('ms_per_step','error','proof_verified_in_exponent')}, 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
"; }
for sp in 1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2975$sp bench.py --gpus 8 --steps 20 --warmup 5 --extras '' --wm-split $sp > gpurun_out/bench_n8_split$sp.json 2> gpurun_out/bench_n8_split$sp.log; echo "n8 split=$sp rc=$? $((SECONDS-t0))s"
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n8_split$sp.log | tail -2 | cut -c1-300; show gpurun_out/bench_n8_split$sp.json n8_split$sp
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29759 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.log; echo "n8 full rc=$? $((SECONDS-t0))s"; show gpurun_out/bench_n8.json n8
du -sh gpurun_out
