"""Throughput of the device verifier (row f-4): n independent verify_proof calls per launch under one verifying key.

Workload: a synthetic key with rs256's instance size (24 gamma_abc points = 23 public inputs) whose discrete
logs are known, and proofs that are VALID by construction: for random a, b and public inputs x the proof
(aG, bH, cG) with c = (ab - alpha*beta - gamma*sum x_i*abc_i) / delta satisfies the pairing equation of verifier.rs:44-65.
Every 7th proof is corrupted; the verdict vector is checked before anything is timed.
Prints one JSON line per batch size: device-resident time (CUDA events on the library's stream) and the end-to-end call
with host buffers (H2D of proofs + inputs, D2H of verdicts inside)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crescent_credentials_b200 import ffi  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402

R = g.R_MOD


def splitmix(seed):
    x = seed & 0xFFFFFFFFFFFFFFFF
    while True:
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        yield z ^ (z >> 31)


def fr_stream(seed):
    it = splitmix(seed)
    while True:
        yield (next(it) | (next(it) << 64) | (next(it) << 128) | (next(it) << 192)) % R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs", type=int, default=23)
    ap.add_argument("--sizes", type=int, nargs="+", default=[1, 64, 1024, 8192, 32768])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--occupancy", type=int, nargs="+", default=[8])
    run(ap.parse_args())


def run(args, cpu_baseline=None):
    """cpu_baseline (bench.py --what verify only): callable(vk_arrays, proofs (n, 32), inputs (n, k, 4), want) -> dict."""
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()  # a real (non-null) stream: the library orders its work on it and the events see it
    torch.cuda.set_stream(stream)
    ctx = ffi.Context(0, stream.cuda_stream)
    fr = fr_stream(0xF4)
    k = args.inputs
    alpha, beta, gamma, delta = (next(fr) for _ in range(4))
    abc = [next(fr) for _ in range(k + 1)]
    g1 = lambda ks: ctx.fixed_base(1, g.fr_to_mont(ks))
    g2 = lambda ks: ctx.fixed_base(2, g.fr_to_mont(ks))
    t0 = time.time()
    vk_arrays = (g1([alpha])[0], g2([beta])[0], g2([gamma])[0], g2([delta])[0], g1(abc))
    ctx.load_vk(*vk_arrays)
    ctx.sync()
    load_s = time.time() - t0
    nmax = max(args.sizes)
    a = [next(fr) for _ in range(nmax)]
    b = [next(fr) for _ in range(nmax)]
    xs = [[next(fr) for _ in range(k)] for _ in range(nmax)]
    di = pow(delta, -1, R)
    c = [((a[i] * b[i] - alpha * beta - gamma * (abc[0] + sum(x * w for x, w in zip(xs[i], abc[1:])))) * di) % R for i in range(nmax)]
    bad = np.arange(nmax) % 7 == 3
    for i in np.nonzero(bad)[0]:
        c[i] = (c[i] + 1) % R
    A, B, Cc = g1(a), g2(b), g1(c)
    proofs = np.zeros((nmax, 34), dtype=np.uint64)  # g16_proof = 8 + 16 + 8 words + 4 x int32
    proofs[:, 0:8], proofs[:, 8:24], proofs[:, 24:32] = A, B, Cc
    x_mont = g.fr_to_mont([v for row in xs for v in row]).reshape(nmax, k, 4)
    want = (~bad).astype(np.uint8)
    cpu = cpu_baseline(vk_arrays, proofs[:, :32], x_mont, want) if cpu_baseline else None
    d_proofs = ctx.dev_alloc(proofs.nbytes)
    d_x = ctx.dev_alloc(max(x_mont.nbytes, 8))
    d_v = ctx.dev_alloc(nmax)
    ctx.dev_upload(d_proofs, proofs)
    ctx.dev_upload(d_x, x_mont)
    for occ, n in [(o_, n_) for o_ in args.occupancy for n_ in args.sizes]:
        ctx.set_option("verify_occupancy", occ)
        got = ctx.verify_batch(proofs[:n].view(np.uint8).reshape(-1), x_mont[:n], n)
        assert (got == want[:n]).all(), f"verdicts differ at n={n}"
        dev_ms, e2e_ms = [], []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            ctx.verify_batch_dev(d_proofs, d_x, n, d_v)
            e1.record(stream)
            torch.cuda.synchronize()
            dev_ms.append(e0.elapsed_time(e1))
            t0 = time.perf_counter()
            ctx.verify_batch(proofs[:n].view(np.uint8).reshape(-1), x_mont[:n], n)
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        out = np.zeros(nmax, dtype=np.uint8)
        ctx.dev_download(d_v, out)
        assert (out[:n] == want[:n]).all()
        d, e = min(dev_ms), min(e2e_ms)
        print(json.dumps({"what": "groth16_verify_batch", "proofs": n, "public_inputs": k, "verify_occupancy": occ, "device_ms": round(d, 3),
                          "device_proofs_per_s": round(n / d * 1e3, 1), "e2e_ms": round(e, 3), "e2e_proofs_per_s": round(n / e * 1e3, 1),
                          "h2d_bytes": int(n * (272 + 32 * k)), "d2h_bytes": n, "verdicts_checked": True,
                          "vk_load_s": round(load_s, 3), "cpu_baseline": cpu}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
