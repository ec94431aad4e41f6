#!/bin/bash
# Round 2: one N = 8 bench line with the final code (automatic window per shard size), extras off.
set -u
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 8 --steps 20 --warmup 5 --extras '' > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.log; echo "n8 rc=$? ${SECONDS}s"
grep -i "error\|Traceback" gpurun_out/bench_n8_final.log | head -3
python -c "
import json
d=json.loads(open('gpurun_out/bench_n8_final.json').read().strip().splitlines()[-1])
print('n8 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'graph', d.get('graph'), d['config'].get('window_bits'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
"
