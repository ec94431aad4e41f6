"""Option sweep on ONE resident problem: builds the workload once, then times full proofs (witness resident) under each
combination of runtime options and checks that every combination returns the same proof bytes.

  python tools/sched_sweep.py --combos "split_chains=0,wm_priority=0,ntt_radix4=0;split_chains=1;split_chains=1,wm_priority=1"

Options not named in a combination keep the library default.  One JSON line per combination on stdout."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import build_problem  # noqa: E402
from crescent_credentials_b200 import ffi  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402

DEFAULTS = {"split_chains": 1, "wm_priority": 0, "ntt_radix4": -1, "spmv_sell": 1, "serialize": 0}

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="S-rs256")
ap.add_argument("--witness", default="uniform")
ap.add_argument("--precompute", type=int, default=1)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--repeat", type=int, default=1, help="walk the list of combinations this many times (noise estimate)")
ap.add_argument("--combos", default="split_chains=0,ntt_radix4=0;split_chains=1,ntt_radix4=0;split_chains=1,ntt_radix4=1;"
                                    "split_chains=1,ntt_radix4=1,wm_priority=1;split_chains=0,ntt_radix4=1,wm_priority=1")
args = ap.parse_args()
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
ctx = ffi.Context(0, tstream.cuda_stream)
inst, pk, qap, td = build_problem(ctx, args.workload, args.witness)
ctx.load_r1cs(inst.nc, inst.ni, inst.m, inst.matrices.row_ptr, inst.matrices.col, inst.matrices.val, inst.matrices.encoding)
ctx.load_pk(pk.arrays, pk.encoding, 0, 1, bool(args.precompute))
r_m = g.fr_to_mont([0x1111222233334444555566667777888899990000AAAABBBBCCCCDDDD % g.R_MOD])[0]
s_m = g.fr_to_mont([0x0F0E0D0C0B0A09080706050403020100FFEEDDCCBBAA9988 % g.R_MOD])[0]  # same (r, s) as bench.py
ctx.upload_witness(inst.z_mont)
first = None
for combo in args.combos.split(";") * args.repeat:
    opts = dict(DEFAULTS)
    for kv in combo.split(","):
        if kv.strip():
            k, v = kv.split("=")
            opts[k.strip()] = int(v)
    for k, v in opts.items():
        ctx.set_option(k, v)
    for _ in range(3):
        proof = ctx.prove_resident(r_m, s_m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for _ in range(args.reps):
        proof = ctx.prove_resident(r_m, s_m)
    e1.record(tstream)
    torch.cuda.synchronize()
    raw = bytes(proof)
    if first is None:
        first = raw
    tm = ctx.timings()
    print(json.dumps({"opts": opts, "ms_per_proof": e0.elapsed_time(e1) / args.reps, "same_proof": raw == first,
                      "stage_ms": {k: round(v, 3) for k, v in tm.items() if isinstance(v, float)}}), flush=True)
