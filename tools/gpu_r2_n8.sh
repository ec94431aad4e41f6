#!/bin/bash
# Round 2, 8-GPU pass: N = 8 and N = 4 bench lines (with the S-mdl1 and sweep extras) on one box.
set -u
mkdir -p gpurun_out
t0=$SECONDS
nvidia-smi -L | wc -l
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
ex=d.get('extra') or {}
for k,v in ex.items():
    if k=='sweep' and v and 'records' in v:
        for rec in v['records']: print('   sweep', {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in rec.items() if kk in ('what','log_n','n_gpus','ms','Mpts_per_s','frac','checked','error')})
    elif v: print('   extra', k, {kk: v.get(kk) for kk in ('ms_per_step','error','proof_verified_in_exponent')}, 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
"; }
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.log; echo "n$n rc=$? $((SECONDS-t0))s"
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n$n.log | tail -3 | cut -c1-300; show gpurun_out/bench_n$n.json n$n
done
du -sh gpurun_out
