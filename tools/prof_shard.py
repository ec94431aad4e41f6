"""Profiling driver for ONE rank's share of an MSM-sharded proof, in a single process on one GPU: loads the key with the
ranges rank `--rank` of `--world` would hold under the staggered plan, then brackets ONE shard run (begin -> h hand-over ->
finish) with cudaProfilerStart/Stop so that `ncu --profile-from-start off` lists exactly that rank's kernels.  Prints the
rank's stage timings (serialised and overlapped)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import R_INT, S_INT, build_problem  # noqa: E402
from crescent_credentials_b200 import ffi, sharded  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="S-rs256")
ap.add_argument("--witness", default="uniform")
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=1)
ap.add_argument("--serialize", type=int, default=1)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--opt", nargs="*", default=[])
args = ap.parse_args()
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
ctx = ffi.Context(0, tstream.cuda_stream)
inst, pk, qap, td = build_problem(ctx, args.workload, args.witness)
m = inst.matrices
ctx.load_r1cs(inst.nc, inst.ni, inst.m, m.row_ptr, m.col, m.val, m.encoding)
h_len = len(pk.arrays["h_query"])
m1 = len(pk.arrays["a_query"]) - 1
plan = sharded.staggered_plan(h_len, m1, args.world)
rank = args.rank
for kv in args.opt:   # window_bits / ba_levels act on bases loaded afterwards
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
ctx.load_pk(pk.arrays, pk.encoding, rank, args.world, True, h_range=plan.h_ranges[rank], z_range=plan.z_ranges[rank])
for kv in args.opt:
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
r_m, s_m = g.fr_to_mont([R_INT % g.R_MOD])[0], g.fr_to_mont([S_INT % g.R_MOD])[0]
owner = rank == plan.wm_rank
ctx.upload_witness(inst.z_mont)
h_dev = 0
if not owner:   # the chunk this rank would receive from the scatter
    hctx = ffi.Context(0)
    hctx.load_r1cs(inst.nc, inst.ni, inst.m, m.row_ptr, m.col, m.val, m.encoding)
    h = hctx.witness_map(inst.z_mont)
    hctx.close()
    lo = plan.h_ranges[rank][0]
    chunk = np.zeros((plan.h_chunk, 4), dtype=np.uint64)
    part = h[lo:lo + plan.h_chunk]
    chunk[:len(part)] = part
    h_dev = ctx.dev_alloc(plan.h_chunk * 32)
    ctx.dev_upload(h_dev, chunk)


def once():
    ctx.prove_shard_begin_dev(r_m, s_m, run_witness_map=owner)
    if owner:
        ctx.prove_shard_finish_dev()
    else:
        ctx.prove_shard_finish_dev(h_dev, plan.h_ranges[rank][0], plan.h_chunk)
    ctx.sync()


ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for ser in ([0, 1] if args.serialize else [0]):
    ctx.set_option("serialize", ser)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(5):
        once()
    ev1.record()
    torch.cuda.synchronize()
    print(json.dumps({"rank": rank, "world": args.world, "serialize": ser, "ms_per_shard_run": ev0.elapsed_time(ev1) / 5,
                      "z_range": plan.z_ranges[rank], "h_range": plan.h_ranges[rank], "timings": ctx.timings(),
                      "msm_stats": {k: ctx.msm_stats(i) for i, k in enumerate(("h", "l", "a", "b_g1", "b_g2"))}}), flush=True)
ctx.set_option("serialize", args.serialize)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.reps):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
