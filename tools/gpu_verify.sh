#!/bin/bash
# GPU pass for row f-4: verifier parity tests, the verification throughput sweep, ncu launch list + one full capture of k_verify.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_verify.py -q -x > gpurun_out/pytest_verify.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 ))s"
tail -3 gpurun_out/pytest_verify.log
timeout 400 python tools/verify_bench.py ${VB_ARGS:-} > gpurun_out/verify_bench.jsonl 2> gpurun_out/verify_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"
cat gpurun_out/verify_bench.jsonl; tail -5 gpurun_out/verify_bench.err
if [ "${VB_NCU:-0}" = "1" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_verify.csv \
    python tools/verify_bench.py --sizes 8192 --reps 1 > /dev/null 2> gpurun_out/ncu_verify_list.err; echo "ncu list rc=$?"
  timeout 400 ncu --set full --import-source on --clock-control none -k "regex:k_verify" -c 1 -f -o gpurun_out/ncu_verify \
    python tools/verify_bench.py --sizes 8192 --reps 1 > gpurun_out/ncu_verify.log 2>&1; echo "ncu full rc=$? $(( $(date +%s) - t0 ))s"
  python tools/ncu_summary.py gpurun_out/ncu_verify.ncu-rep > gpurun_out/ncu_verify.txt 2>&1
  grep -E "k_verify|k_prepare_inputs|k_abc_table|k_g2_prepare|k_pairing" gpurun_out/launches_verify.csv | cut -d, -f5,12- | head -20
  head -40 gpurun_out/ncu_verify.txt
fi
