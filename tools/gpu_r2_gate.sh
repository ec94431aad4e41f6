#!/bin/bash
# Round 2 (historical: the gated variant was removed afterwards): wm_first = 2 (digit stages of the z-only MSMs beside the witness map, point stages after it) against wm_first = 1.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'graph', d.get('graph'), {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})"; }
run() { timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/gate_$1.json 2> gpurun_out/gate_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/gate_$1.json $1; }
run base ""
run gate "--opt wm_first=2"
run gate_nobatch "--opt wm_first=2 ntt_batch=0"
run gate_radix2 "--opt wm_first=2 ntt_batch=0 ntt_radix4=0"
run base2 ""
run gate2 "--opt wm_first=2"
run mdl1_base "--workload S-mdl1"
run mdl1_gate "--workload S-mdl1 --opt wm_first=2"
