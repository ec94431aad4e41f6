#!/bin/bash
# Round 2 (historical: neither knob changed anything, the option and the variant were removed): DRAM fetch granularity of the level-0 point gathers (364 B of DRAM traffic per slot against 232 algorithmic):
# the device-wide L2 fetch-granularity limit and the PTX L2::64B prefetch-size hint on the gathers.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'frac', round(r.get('frac'),4), 'L0 ms', round(r.get('launch_ms'),4), 'wm', round(d['stage_ms']['witness_map_ms'],2))"; grep "L2 fetch" ${1%.json}.log | head -1; }
run() { G16_LIB=${3:-$PWD/crescent_credentials_b200/libg16b200.so} timeout 100 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/l2_$1.json 2> gpurun_out/l2_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/l2_$1.json $1; }
run base ""
run fetch64 "--opt l2_fetch=64"
run fetch32 "--opt l2_fetch=32"
run hint64 "" $PWD/gpurun_variants/libg16_g64.so
run hint64_fetch32 "--opt l2_fetch=32" $PWD/gpurun_variants/libg16_g64.so
