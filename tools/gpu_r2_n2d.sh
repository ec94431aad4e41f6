#!/bin/bash
# Round 2: one N = 2 bench line with the final code, extras off.
set -u
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29782 bench.py --gpus 2 --steps 20 --warmup 5 --extras '' > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.log; echo "n2 rc=$? ${SECONDS}s"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1])
print('n2 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'graph', d.get('graph'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
"
