"""Standalone MSM / NTT micro-benchmark (the synthetic sweep of BASELINE.json config 5).
Bases are random multiples of the generator minted on the GPU; scalars uniform Fr; result is checked against
(sum s_i k_i) * G computed through an independent path (field ops + one fixed-base multiplication)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from crescent_credentials_b200 import ffi, synth  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logn", type=int, nargs="+", default=[21])
ap.add_argument("--group", type=int, default=1)
ap.add_argument("--precompute", type=int, default=0)
ap.add_argument("--window-bits", type=int, default=0)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--ntt", type=int, default=0)
ap.add_argument("--opt", nargs="*", default=[], help="library options, key=value (e.g. ntt_radix4=0)")
args = ap.parse_args()
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
ctx = ffi.Context(0, tstream.cuda_stream)
ctx.set_option("kernel_events", 1)
ctx.set_option("acc_variant", args.variant)
for kv in args.opt:
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for logn in args.logn:
    n = 1 << logn
    if args.ntt:
        x = ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, synth.uniform_fr_canonical(7, 1, n))
        d = ctx.dev_alloc(n * 32)
        ctx.dev_upload(d, x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = {}
        for name, inv, cos in (("fwd", 0, 0), ("inv", 1, 0), ("coset_fwd", 0, 1), ("coset_inv", 1, 1)):
            ctx.check(ctx.lib.g16_ntt_dev(ctx.h, d, logn, inv, cos))
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                ctx.check(ctx.lib.g16_ntt_dev(ctx.h, d, logn, inv, cos))
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / args.reps
        ctx.dev_free(d)
        print(json.dumps({"ntt_log_n": logn, "opt": args.opt, "ms": res, "GBps_alg_1pass": 64.0 * n / (res["fwd"] * 1e-3) / 1e9,
                          "gmul_per_s": (n / 2) * logn / (res["fwd"] * 1e-3) / 1e9}))
        continue
    ks = ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, synth.uniform_fr_canonical(11, 1, n))
    sc = ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, synth.uniform_fr_canonical(12, 1, n))
    pts = ctx.fixed_base(args.group, ks)
    ctx.check(ctx.lib.g16_msm_set_bases(ctx.h, 0, args.group, pts.ctypes.data, n, args.window_bits, args.precompute))
    dsc = ctx.dev_alloc(n * 32)
    ctx.dev_upload(dsc, sc)
    out = np.zeros(8 * args.group, dtype=np.uint64)
    import ctypes as C
    inf = C.c_int(0)
    ctx.check(ctx.lib.g16_msm_run_dev(ctx.h, 0, dsc, n, out.ctypes.data, C.byref(inf)))
    if args.check:
        dot = ctx.field_op(ffi.FIELD_FR, ffi.OP_MUL, ks, sc)
        tot = sum(g.fr_from_mont(dot)) % g.R_MOD
        want = ctx.fixed_base(args.group, g.fr_to_mont([tot]))[0]
        assert np.array_equal(want, out), "MSM result differs from (sum s_i k_i) * G"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms, acc_ms = [], []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        e0.record()
        ctx.check(ctx.lib.g16_msm_run_dev(ctx.h, 0, dsc, n, None, None))
        e1.record()
        torch.cuda.synchronize()
        tot_ms.append(e0.elapsed_time(e1))
        acc_ms.append(ctx.timings()["acc_ms"]["h"])
    ctx.dev_free(dsc)
    t = min(tot_ms)
    print(json.dumps({"msm_group": args.group, "log_n": logn, "precompute": args.precompute, "window_bits": args.window_bits,
                      "variant": args.variant, "ms": t, "acc_ms": min(acc_ms), "Mpts_per_s": n / (t * 1e-3) / 1e6,
                      "checked": bool(args.check)}))
