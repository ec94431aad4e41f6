#!/bin/bash
# Round 2: GLV scaling kernel + scheduling options at N = 1, small circuits, one shard rank.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_sched.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"; tail -2 gpurun_out/pytest_sched.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'), {k: round(v,2) for k,v in d['stage_ms'].items() if isinstance(v,float) and k.endswith('_ms')})"; }
run() { timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/sched_$1.json 2> gpurun_out/sched_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/sched_$1.json $1; }
run base ""
run wm_first "--opt wm_first=1"
run wm_first_prio "--opt wm_first=1 chain_priority=1"
run prio "--opt chain_priority=1"
run wm_first2 "--opt wm_first=1"
for w in 'S-2^12' 'S-2^16'; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --extras '' --no-cpu-baseline --inflight 0 > gpurun_out/sched_$w.json 2> gpurun_out/sched_$w.log; show gpurun_out/sched_$w.json $w
  timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --extras '' --no-cpu-baseline --inflight 0 --opt wm_first=1 > gpurun_out/sched_${w}_wmf.json 2> gpurun_out/sched_${w}_wmf.log; show gpurun_out/sched_${w}_wmf.json ${w}_wm_first
done
timeout 300 python tools/prof_shard.py --world 8 --rank 1 > gpurun_out/prof_shard8_rank1.log 2>&1; grep ms_per_shard_run gpurun_out/prof_shard8_rank1.log | cut -c1-520
