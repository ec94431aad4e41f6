#!/bin/bash
# Round 2: N = 4 and N = 2 after the scheduling change (witness map first) and the GLV scaling kernel, extras off, on a 4-GPU box.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'), 'plan', d['config'].get('plan'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
"; }
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $2 --steps 20 --warmup 5 --extras '' $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; grep -i "error\|Traceback" gpurun_out/bench_$1.log | head -3; show gpurun_out/bench_$1.json $1; }
run n4 4 "" 29771
run n4_wmf0 4 "--opt wm_first=0" 29772
run n2 2 "" 29773
run n2_wmf0 2 "--opt wm_first=0" 29774
run n4_share15 4 "--rank0-share 0.15" 29775
