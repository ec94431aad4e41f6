#!/bin/bash
# Round 2: Pippenger window at N = 1 (never tried below 19) and c = 16 on one rank of two.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'frac', (d.get('roofline') or {}).get('frac'), {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})"; }
run() { timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/win_$1.json 2> gpurun_out/win_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/win_$1.json $1; }
run c18 "--window-bits 18"
run c17 "--window-bits 17"
run c16 "--window-bits 16"
run mdl1_c18 "--workload S-mdl1 --window-bits 18"
timeout 200 python tools/prof_shard.py --world 2 --rank 1 --opt window_bits=16 > gpurun_out/tune_w2_c16.log 2>&1; grep ms_per_shard gpurun_out/tune_w2_c16.log | cut -c1-80
