#!/bin/bash
# Experiment pass: affine-level parity tests, then the option sweep (walked twice for a noise estimate).
# Usage: bash tools/gpu_pass2.sh "<combos>" [pytest -k expression]
set -u
mkdir -p gpurun_out
COMBOS=${1:-"ba_prefetch=0;ba_prefetch=1"}
KEXPR=${2:-"batched_affine or golden_prove or msm"}
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$KEXPR" > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"
tail -3 gpurun_out/pytest_k.log
t0=$SECONDS
timeout 600 python tools/sched_sweep.py --reps 20 --repeat 2 --combos "$COMBOS" > gpurun_out/sched_sweep2.jsonl 2> gpurun_out/sched_sweep2.err; echo "sweep rc=$? ($((SECONDS-t0)) s)"
python - <<'PY'
import json
for l in open("gpurun_out/sched_sweep2.jsonl"):
    d = json.loads(l)
    st = d["stage_ms"]
    print({k: v for k, v in d["opts"].items()}, round(d["ms_per_proof"], 2), d["same_proof"],
          {k: st[k] for k in ("witness_map_ms", "msm_h_ms", "msm_l_ms", "msm_a_ms", "msm_b_g1_ms", "msm_b_g2_ms") if k in st})
PY
tail -3 gpurun_out/sched_sweep2.err
