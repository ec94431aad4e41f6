#!/bin/bash
# Round 2: window and level count of the shard-sized MSMs (one rank of 8, run alone on one GPU).
set -u
mkdir -p gpurun_out
t0=$SECONDS
run() { timeout 200 python tools/prof_shard.py --rank 1 $2 > gpurun_out/tune_$1.log 2>&1; echo "$1 rc=$? $((SECONDS-t0))s"; python - <<PY
import json
for l in open('gpurun_out/tune_$1.log'):
    if l.startswith('{'):
        d=json.loads(l); t=d['timings']
        print('  $1 ser', d['serialize'], 'ms', round(d['ms_per_shard_run'],3), {k: round(v,2) for k,v in t.items() if isinstance(v,float) and k.startswith('msm_')}, 'c', {k: (v.get('window_bits'), v.get('levels')) for k,v in d['msm_stats'].items()})
PY
}
run w8_c14 "--world 8 --opt window_bits=14"
run w8_c13 "--world 8 --opt window_bits=13"
run w8_c14_lv6 "--world 8 --opt window_bits=14 ba_levels=6"
run w8_c13_lv6 "--world 8 --opt window_bits=13 ba_levels=6"
run w4_base "--world 4"
run w4_c17 "--world 4 --opt window_bits=17"
run w4_c16 "--world 4 --opt window_bits=16"
run w4_c15 "--world 4 --opt window_bits=15"
run w2_base "--world 2"
run w2_c18 "--world 2 --opt window_bits=18"
run w2_c17 "--world 2 --opt window_bits=17"
