#!/bin/bash
# A/B for the next round: Fq2 products with the three Montgomery products in lock-step (fp.cuh mul_cios3, -DG16_FQ2_MUL3) against
# the default build -- parity first (the alternative library must pass the G2 / verifier tests), then the verification sweep and the
# N = 1 proof.  Build the alternative library HERE before calling gpurun (it travels with the snapshot):
#   make -C crescent_credentials_b200/csrc EXTRA=-DG16_FQ2_MUL3 BUILD=build_mul3 OUT=../libg16b200_mul3.so -j8
# Usage (via gpurun, repo root):  bash tools/gpu_ab_mul3.sh
set -u
mkdir -p gpurun_out
ALT=$PWD/crescent_credentials_b200/libg16b200_mul3.so
[ -f "$ALT" ] || { echo "build $ALT first (see the header of this script)"; exit 2; }
G16_LIB=$ALT timeout 400 python -m pytest tests/test_gpu_verify.py tests/test_gpu_parity.py -q -x -k "verify or pairing or g2 or golden_prove" > gpurun_out/ab_mul3_pytest.log 2>&1
echo "pytest(alt) rc=$?"; tail -2 gpurun_out/ab_mul3_pytest.log
for lib in "" "$ALT"; do
  tag=$([ -z "$lib" ] && echo default || echo mul3)
  G16_LIB=${lib:-$PWD/crescent_credentials_b200/libg16b200.so} timeout 200 python tools/verify_bench.py --sizes 1 65536 --reps 3 > gpurun_out/ab_mul3_verify_$tag.jsonl 2>&1
  G16_LIB=${lib:-$PWD/crescent_credentials_b200/libg16b200.so} timeout 200 python bench.py --no-cpu-baseline --inflight 0 --extras '' > gpurun_out/ab_mul3_bench_$tag.json 2>/dev/null
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for l in open(f"gpurun_out/ab_mul3_verify_{tag}.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(tag, "verify", d["proofs"], d["device_ms"], d["device_proofs_per_s"])
d = json.loads(open(f"gpurun_out/ab_mul3_bench_{tag}.json").read().strip().splitlines()[-1])
print(tag, "prove ms", d["ms_per_step"], "b_g2 chain", d["stage_ms"].get("msm_b_g2_ms"))
PY
done
