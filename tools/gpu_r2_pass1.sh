#!/bin/bash
# Round 2, first GPU pass: all GPU tests (incl. the full-size parity tests), the N = 1 bench line, launch lists of a whole proof
# and of one rank's share at N = 8, one ncu --set full sweep over the kernels that dominate, and the Fq2 lock-step product A/B.
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $((SECONDS-t0))s"
tail -22 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$? $((SECONDS-t0))s"
tail -4 gpurun_out/bench_n1.log; cut -c1-300 gpurun_out/bench_n1.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_prove.csv \
  python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/prof_prove.log 2>&1; echo "ncu list rc=$? $((SECONDS-t0))s"
python tools/agg_launches.py gpurun_out/launches_prove.csv 2>/dev/null | head -40
for rk in 1 0; do
  timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard8_rank$rk.csv \
    python tools/prof_shard.py --world 8 --rank $rk > gpurun_out/prof_shard8_rank$rk.log 2>&1; echo "ncu shard list rank $rk rc=$? $((SECONDS-t0))s"
  grep ms_per_shard_run gpurun_out/prof_shard8_rank$rk.log | cut -c1-600
done
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  -k regex:'k_ba_add|k_ba_products|k_ba_invert|k_ntt_pass4|k_spmv_sell|k_chunk_reduce|k_bucket_tail|k_rs_scatter' -f -o gpurun_out/ncu_r02_all \
  python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 > gpurun_out/ncu_r02_all.log 2>&1; echo "ncu full rc=$? $((SECONDS-t0))s"
python tools/ncu_summary.py gpurun_out/ncu_r02_all.ncu-rep > gpurun_out/ncu_r02_all.txt 2>&1; ls -la gpurun_out/ncu_r02_all.ncu-rep
# (round 2, pass 1 also ran the A/B of the -DG16_FQ2_MUL3 build here: slower on both counts, script and variant removed)
