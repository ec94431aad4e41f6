"""BASELINE config 5: the synthetic sweep -- stand-alone BN254 MSM (G1 / G2) and Fr NTT, 2^16 .. 2^26, on 1 / 2 / 4 / 8 GPUs.

  python tools/sweep.py [--g1 16 18 20 22 24 26] [--g2 16 18 20 22 24] [--ntt 16 18 20 22 24 26] [--precompute 1]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tools/sweep.py ...       # N = 2, 4, 8

MSM: uniform Fr scalars, distinct points P_i = k_i * G (k_i from the seeded SplitMix64 stream of SURVEY 8d, minted on the
GPU), sharded by contiguous point range: rank g keeps pairs [g N / G, (g + 1) N / G) and their 2^(c w) window tables, runs the
full Pippenger pipeline on them, and the G partial sums (XYZZ, 128 / 256 bytes) are gathered with ONE all_gather and added on
rank 0 (g16_msm_copy_result_dev / g16_msm_combine_dev) -- no other collective.  Every result is checked against
(sum_i s_i k_i) * G, computed through an independent path (element-wise products + one fixed-base multiplication).
NTT: one GPU (it does not shard), forward / inverse / coset, natural order in and out.

Every record carries `frac`: SURVEY 8d's algorithmic Fq/Fr products (MSM: 160 N for G1, 480 N for G2 at the canonical c = 16, or
the minimising c for N <= 2^18; NTT: (n / 2) log2 n) per second, over the MEASURED product peak of this GPU x the GPUs used.
Time = device time with CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_RINV = pow(1 << 256, -1, R_MOD)


def alg_fq_mul(n: int, group: int):
    """bucket accumulation 10 N W + reduction 28 2^(c-1) W Fq products (x3 in G2); canonical c = 16 (160 N) above
    2^18 points, the minimising c below.  Returns (products, c)."""
    mult = 1 if group == 1 else 3
    if n > (1 << 18):
        return 160.0 * n * mult, 16
    best = None
    for c in range(4, 21):
        w = -(-254 // c)
        cost = 10.0 * n * w + 28.0 * (1 << (c - 1)) * w
        if best is None or cost < best[0]:
            best = (cost, c)
    return best[0] * mult, best[1]


def mont_sum(a: np.ndarray) -> int:
    """Exact sum of (n, 4) uint64 Montgomery elements as a canonical integer mod r (linear: sum of representatives)."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    tot = 0
    for k in range(4):
        lo = int((a[:, k] & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))   # n < 2^32 elements: no overflow
        hi = int((a[:, k] >> np.uint64(32)).sum(dtype=np.uint64))
        tot += (lo + (hi << 32)) << (64 * k)
    return tot % R_MOD * _RINV % R_MOD


def _timed(torch, dist, world, fn, reps):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(reps):
        barrier()
        ev0.record()
        fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        best = ms if best is None else min(best, ms)
    if world > 1:
        t = torch.tensor([best], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t[0])
    return best


def msm_point(ctx, group: int, logn: int, rank: int, world: int, tstream, peak_gmul: float, reps: int = 5,
              precompute: bool = True, window_bits: int = 0, slot: int = 7):
    """One (group, size) point of the sweep on `world` ranks.  ctx's main stream must be torch's current stream `tstream`."""
    import torch
    import torch.distributed as dist
    from crescent_credentials_b200 import ffi, synth
    from crescent_credentials_b200 import groth16 as g
    n = 1 << logn
    lo, hi = n * rank // world, n * (rank + 1) // world
    cnt = hi - lo
    F = ffi.FIELD_FR
    ks = ctx.field_op(F, ffi.OP_TO_MONT, synth.uniform_fr_canonical(11 + logn, group, cnt, first=lo))
    sc = ctx.field_op(F, ffi.OP_TO_MONT, synth.uniform_fr_canonical(12 + logn, group, cnt, first=lo))
    pb = 64 * group
    d_k, d_s, d_p = ctx.dev_alloc(cnt * 32), ctx.dev_alloc(cnt * 32), ctx.dev_alloc(cnt * pb)
    words = 16 * group
    with torch.cuda.stream(tstream):
        mine = torch.zeros((words,), dtype=torch.int64, device="cuda")
        allp = torch.zeros((world, words), dtype=torch.int64, device="cuda")
    try:
        ctx.dev_upload(d_k, ks)
        ctx.dev_upload(d_s, sc)
        ctx.fixed_base_dev(group, d_k, cnt, d_p)
        ctx.msm_set_bases_dev(slot, group, d_p, cnt, window_bits, precompute)
        ctx.dev_free(d_p)
        d_p = 0
        ctx.dev_free(d_k)
        d_k = 0

        result = [None]

        def once():
            ctx.msm_run_dev(slot, d_s, cnt, want_result=False)
            ctx.msm_copy_result_dev(slot, mine.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(allp.view(-1), mine)
                if rank == 0:
                    result[0] = ctx.msm_combine_dev(group, allp.data_ptr(), world)
            else:
                result[0] = ctx.msm_combine_dev(group, mine.data_ptr(), 1)

        with torch.cuda.stream(tstream):
            once()   # warm-up + the checked result
            torch.cuda.synchronize()
            part = mont_sum(ctx.field_op(F, ffi.OP_MUL, ks, sc))
            if world > 1:
                tot_t = torch.tensor([(part >> (62 * k)) & ((1 << 62) - 1) for k in range(5)], device="cuda", dtype=torch.int64)
                all_t = torch.zeros((world, 5), device="cuda", dtype=torch.int64)
                dist.all_gather_into_tensor(all_t.view(-1), tot_t)
                rows = all_t.cpu().tolist()
                total = sum(sum(int(v) << (62 * k) for k, v in enumerate(row)) for row in rows) % R_MOD
            else:
                total = part
            checked = None
            if rank == 0:
                want = ctx.fixed_base(group, g.fr_to_mont([total]))[0]
                got, inf = result[0]
                checked = bool(np.array_equal(want, got)) and not inf
                if not checked:
                    raise RuntimeError(f"sweep: MSM g{group} 2^{logn} on {world} ranks differs from (sum s_i k_i) * G")
            ms = _timed(torch, dist, world, once, reps)
        stats = None
        alg, c_alg = alg_fq_mul(n, group)
        rec = {"what": f"msm_g{group}", "log_n": logn, "n_gpus": world, "points_per_rank": cnt, "precompute": int(precompute),
               "ms": ms, "Mpts_per_s": n / (ms * 1e-3) / 1e6, "checked": checked,
               "algorithmic_fq_mul": alg, "algorithmic_c": c_alg, "achieved_gmul_s": alg / (ms * 1e-3) / 1e9,
               "peak_gmul_s": peak_gmul * world, "frac": alg / (ms * 1e-3) / 1e9 / (peak_gmul * world),
               "algorithmic_bytes": (96 if group == 1 else 160) * n}
        return rec if rank == 0 else None
    finally:
        for p in (d_k, d_s, d_p):
            if p:
                ctx.dev_free(p)
        # drop the slot's tables and scratch (GBs at the large sizes) before the next point
        ctx.lib.g16_msm_set_bases_dev(ctx.h, slot, group, None, 0, 0, 0)


def ntt_point(ctx, logn: int, peak_gmul: float, reps: int = 5):
    """Fr NTT of size 2^logn on one GPU: forward / inverse / coset variants through the natural-order API."""
    import torch
    from crescent_credentials_b200 import ffi, synth
    n = 1 << logn
    x = ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, synth.uniform_fr_canonical(7, 1, n))
    d = ctx.dev_alloc(n * 32)
    try:
        ctx.dev_upload(d, x)
        res = {}
        for name, inv, cos in (("fwd", 0, 0), ("inv", 1, 0), ("coset_fwd", 0, 1), ("coset_inv", 1, 1)):
            ctx.ntt_dev(d, logn, inv, cos)
            res[name] = _timed(torch, None, 1, lambda: ctx.ntt_dev(d, logn, inv, cos), reps)
        # round trip leaves the data where it started: fwd/inv and coset_fwd/coset_inv ran the same number of times
        y = ctx.dev_download(d, np.empty_like(x))
        ok = bool(np.array_equal(x, y))
    finally:
        ctx.dev_free(d)
    muls = (n / 2) * logn
    t = res["fwd"] * 1e-3
    return {"what": "ntt_fr", "log_n": logn, "n_gpus": 1, "ms": res, "round_trip_identity": ok, "algorithmic_fr_mul": muls,
            "achieved_gmul_s": muls / t / 1e9, "peak_gmul_s": peak_gmul, "frac": muls / t / 1e9 / peak_gmul,
            "algorithmic_GBps_1pass": 64.0 * n / t / 1e9,
            "note": "natural-order API: includes the bit-reversal pass the witness map never runs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g1", type=int, nargs="*", default=[16, 18, 20, 22, 24, 26])
    ap.add_argument("--g2", type=int, nargs="*", default=[16, 18, 20, 22, 24])
    ap.add_argument("--ntt", type=int, nargs="*", default=[16, 18, 20, 22, 24, 26])
    ap.add_argument("--precompute", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--opt", nargs="*", default=[])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from crescent_credentials_b200 import ffi
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = ffi.Context(local, tstream.cuda_stream)
    for kv in args.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    peak = ctx.bench_int_pipe(3)
    for group, sizes in ((1, args.g1), (2, args.g2)):
        for logn in sizes:
            # tables of 2^(c w) multiples need windows x the points: fall back to plain bases when they would not fit
            pre = bool(args.precompute) and (64 * group << logn) * 14 // world < 60e9
            try:
                rec = msm_point(ctx, group, logn, rank, world, tstream, peak, args.reps, precompute=pre)
            except Exception as e:
                rec = {"what": f"msm_g{group}", "log_n": logn, "n_gpus": world, "error": repr(e)}
                if rank != 0:
                    rec = None
            if rank == 0:
                print(json.dumps(rec), flush=True)
    if rank == 0:
        for logn in args.ntt:
            print(json.dumps(ntt_point(ctx, logn, peak, args.reps)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
