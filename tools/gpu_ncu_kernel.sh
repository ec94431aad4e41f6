#!/bin/bash
# ncu --set full on the first COUNT launches matching REGEX inside one proof.
# Usage: bash tools/gpu_ncu_kernel.sh TAG REGEX COUNT "opt1=v opt2=v"
set -u
TAG=$1; RE=$2; CNT=${3:-2}; OPTS=${4:-}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:$RE" -c $CNT -f -o gpurun_out/ncu_$TAG \
  python tools/prof_prove.py --precompute 1 --serialize 1 --reps 1 --opt $OPTS > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu $TAG rc=$?"
python tools/ncu_summary.py gpurun_out/ncu_$TAG.ncu-rep > gpurun_out/ncu_$TAG.txt 2>&1
ls -la gpurun_out/ncu_$TAG.ncu-rep
