#!/bin/bash
# Round 2: the new automatic window (round(log2 n) - 3 from 2^17 points on) at N = 1, level count beside it, smaller windows
# for the small circuits, and a fresh ncu --set full capture of the roofline launch (h-query level-0 k_ba_add) for bench.py's traffic.
set -u
mkdir -p gpurun_out
t0=$SECONDS
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sha', d.get('proof_sha256','')[:8], 'frac', r.get('frac'), 'slots', r.get('slots_per_launch'), 'L0 ms', r.get('launch_ms'), d['config'].get('window_bits'))"; }
run() { timeout 300 python bench.py --steps 20 --warmup 3 --extras '' --no-cpu-baseline --inflight 0 $2 > gpurun_out/win2_$1.json 2> gpurun_out/win2_$1.log; echo "$1 rc=$? $((SECONDS-t0))s"; show gpurun_out/win2_$1.json $1; }
run auto ""
run auto_lv6 "--ba-levels 6"
run c17 "--window-bits 17"
run c17_lv6 "--window-bits 17 --ba-levels 6"
run s12 "--workload S-2^12 --steps 50"
run s12_c10 "--workload S-2^12 --steps 50 --window-bits 10"
run s12_c9 "--workload S-2^12 --steps 50 --window-bits 9"
run s16 "--workload S-2^16 --steps 50"
run s16_c14 "--workload S-2^16 --steps 50 --window-bits 14"
run s16_c13 "--workload S-2^16 --steps 50 --window-bits 13"
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:^k_ba_add$" -s 20 -c 1 -f -o gpurun_out/ncu_ba_add_h_c18 \
    python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/ncu_ba_add_h_c18.log 2>&1; echo "ncu rc=$? $((SECONDS-t0))s"
python tools/ncu_summary.py gpurun_out/ncu_ba_add_h_c18.ncu-rep > gpurun_out/ncu_ba_add_h_c18.txt 2>&1
ncu -i gpurun_out/ncu_ba_add_h_c18.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=dict(zip(h,r)); print({k:d[k] for k in d if k in ('Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread')})"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
