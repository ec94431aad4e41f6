#!/bin/bash
# Round 2, third 8-GPU pass: witness map split over three ranks that keep their wire MSMs (b on rank 1, c on rank 2), with and
# without more NCCL point-to-point channels.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x -k "nccl or parts" > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$? $((SECONDS-t0))s"; tail -3 gpurun_out/pytest_sharded.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'graph', d.get('graph'), 'plan', d['config'].get('plan'))
for k,v in (d.get('rank_stage_ms') or {}).items(): print('   ', k, v)
"; }
run() {  # tag, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 20 --warmup 5 --extras '' $2 > gpurun_out/bench_n8_$1.json 2> gpurun_out/bench_n8_$1.log; echo "n8 $1 rc=$? $((SECONDS-t0))s"
  grep -i "error\|Traceback" gpurun_out/bench_n8_$1.log | head -3; show gpurun_out/bench_n8_$1.json n8_$1
}
run split3 "--wm-split 1" 29761
run split3_p2p "--wm-split 1 --nccl-env NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" 29762
run nosplit "" 29763
du -sh gpurun_out
