#!/bin/bash
# Round 2: scheduling A/B at N = 1 (witness map first, more batched-affine levels), same box, back to back.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'), {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})"; }
run() { timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/sched_$1.json 2> gpurun_out/sched_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/sched_$1.json $1; }
run base ""
run wm_first "--opt wm_first=1"
run levels6 "--ba-levels 6"
run levels7 "--ba-levels 7"
run wm_first_levels6 "--opt wm_first=1 --ba-levels 6"
run base2 ""
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "golden_prove or stream_plans" > gpurun_out/pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sched.log
