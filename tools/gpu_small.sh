#!/bin/bash
# Small-circuit latency experiments: assembly tables on/off, MSM options, at S-2^12 / S-2^16 / S-rs256.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "assembly_fixed_base or golden_prove" > gpurun_out/pytest_asm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_asm.log
run() { # workload, label, extra args
  timeout 200 python bench.py --workload $1 --no-cpu-baseline --inflight 0 --steps 50 --extras "" ${@:3} > gpurun_out/small_$2.json 2> gpurun_out/small_$2.log
  python -c "
import json
d=json.loads(open('gpurun_out/small_$2.json').read().strip().splitlines()[-1]); st=d['stage_ms']
print('$1 $2', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in st.items() if isinstance(v,float)})"
}
run S-2^12 t1 --opt asm_tables=1
run S-2^12 t0 --opt asm_tables=0
run S-2^12 t1_ba0 --opt asm_tables=1 --ba-levels 0
run S-2^12 t1_nopre --opt asm_tables=1 --precompute 0
run S-2^12 t1_ba0_nopre --opt asm_tables=1 --ba-levels 0 --precompute 0
run S-2^12 t1_noshare --opt asm_tables=1 --share-digits 0
run S-2^16 t1_2^16 --opt asm_tables=1
run S-2^16 t1_ba0_2^16 --opt asm_tables=1 --ba-levels 0
timeout 200 python bench.py --no-cpu-baseline --inflight 0 --extras "" > gpurun_out/small_rs256.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/small_rs256.json').read().strip().splitlines()[-1]); print('S-rs256', d['ms_per_step'], d['e2e']['ms_per_step'])"
