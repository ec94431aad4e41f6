"""Synthetic rs256-class R1CS instances: the real circuit artifacts (main_c.r1cs, prover_params.bin)
cannot be generated offline (no circom / cargo), so the bench and the full-size tests prove over a seeded synthetic
R1CS of the same constraint count, wire count, non-zero count and domain size.

  S-rs256 : n = 2^21, nc = 1,500,000, l = 24, m = 1,450,000, nnz(A,B,C) ~ 7M / 5M / 4M
  S-mdl1  : n = 2^22, nc = 3,000,000, l = 40, m = 2,900,000, nnz ~ 14M / 10M / 8M

Row lengths are geometric (cap 256), 30 % of the column mass sits on the first 2^16 wires, coefficients are 70 % +-1,
20 % +-2^k (k < 121), 10 % uniform.  The witness is either uniform Fr (worst case, the headline) or circom-like
(85 % bits, 10 % bytes, 4 % < 2^121, 1 % uniform).  Every row is made satisfiable by giving C a term on the constant
wire 0 whose coefficient is solved on the GPU: k0 = <A_i,z><B_i,z> - <C_i,z>.  (nc > m, so "one fresh wire per row"
is impossible; the constant-wire term is the closest equivalent.)

All streams are SplitMix64 of (seed, tag, index): every rank and the CPU baseline regenerate identical data."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import ffi
from .groth16 import ConstraintMatrices, R_MOD

CONFIGS = {
    "S-rs256": dict(nc=1_500_000, ni=24, m=1_450_000, mean=(4.7, 3.3, 2.7)),
    "S-mdl1": dict(nc=3_000_000, ni=40, m=2_900_000, mean=(4.7, 3.3, 2.7)),
    "S-2^16": dict(nc=60_000, ni=8, m=58_000, mean=(4.7, 3.3, 2.7)),
    "S-2^12": dict(nc=3_500, ni=6, m=3_300, mean=(4.7, 3.3, 2.7)),
}

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream(seed: int, tag: int, n: int, lane: int = 0, first: int = 0) -> np.ndarray:
    """Elements [first, first + n) of the stream (seed, tag, lane): a rank can draw its own slice of a long vector."""
    base = np.uint64((seed ^ (tag << 40) ^ (lane << 56)) & 0xFFFFFFFFFFFFFFFF)
    return splitmix64(np.arange(first, first + n, dtype=np.uint64) ^ base)


def uniform_fr_canonical(seed: int, tag: int, n: int, first: int = 0) -> np.ndarray:
    """n elements uniform in [0, 2^253) as canonical little-endian limbs (2^253 < r); elements [first, first + n) of the stream."""
    out = np.empty((n, 4), dtype=np.uint64)
    for k in range(4):
        out[:, k] = stream(seed, tag, n, lane=k + 1, first=first)
    out[:, 3] &= np.uint64((1 << 61) - 1)
    return out


def _int_limbs(v: int):
    return [(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]


def _coeff_table() -> np.ndarray:
    rows = [_int_limbs(1), _int_limbs(R_MOD - 1)]
    rows += [_int_limbs(1 << k) for k in range(121)]
    rows += [_int_limbs(R_MOD - (1 << k)) for k in range(121)]
    return np.array(rows, dtype=np.uint64)


def _matrix(seed: int, tag: int, nc: int, m: int, mean: float, min_len: int, col_lo: int, absent_pct: int = 0):
    """Returns row_ptr, col, val (canonical) of one random sparse matrix.  absent_pct > 0 keeps that share of the wires
    out of the matrix altogether (w % 20 < absent_pct / 5), which is what makes b_g1 / b_g2 query points infinity."""
    u = (stream(seed, tag, nc) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    p = 1.0 / mean
    lens = np.floor(np.log(np.maximum(u, 1e-300)) / np.log(1.0 - p)).astype(np.int64) + min_len
    lens = np.minimum(lens, 256)
    row_ptr = np.zeros(nc + 1, dtype=np.uint64)
    row_ptr[1:] = np.cumsum(lens).astype(np.uint64)
    nnz = int(row_ptr[-1])
    sel = stream(seed, tag + 1, nnz)
    pick = stream(seed, tag + 2, nnz)
    span = np.uint64(m - col_lo)
    hot = np.uint64(min(1 << 16, m - col_lo))
    col = np.where((sel % np.uint64(10)) < np.uint64(3), pick % hot, pick % span).astype(np.uint32) + np.uint32(col_lo)
    if absent_pct:
        skip = absent_pct // 5                      # wires with (w % 20) < skip never occur
        keep = 20 - skip
        k = (col.astype(np.int64) * keep) // 20     # uniform over the kept wires, same hot/cold shape
        col = np.minimum(20 * (k // keep) + skip + (k % keep), m - 1).astype(np.uint32)
    kind = (sel >> np.uint64(8)) % np.uint64(10)
    sign = ((sel >> np.uint64(16)) & np.uint64(1)).astype(np.int64)
    k = ((sel >> np.uint64(20)) % np.uint64(121)).astype(np.int64)
    idx = np.where(kind < np.uint64(7), sign, 2 + k + 121 * sign)
    val = _coeff_table()[idx]
    uni = np.nonzero(kind >= np.uint64(9))[0]
    if uni.size:
        val[uni] = uniform_fr_canonical(seed, tag + 3, nnz)[uni]
    return row_ptr, col, val


def witness_canonical(seed: int, m: int, mode: str) -> np.ndarray:
    z = uniform_fr_canonical(seed, 0x77, m)
    if mode == "circom":
        sel = stream(seed, 0x78, m) % np.uint64(100)
        bits = sel < np.uint64(85)
        byte = (sel >= np.uint64(85)) & (sel < np.uint64(95))
        mid = (sel >= np.uint64(95)) & (sel < np.uint64(99))
        z[bits, 0] &= np.uint64(1)
        z[bits, 1:] = 0
        z[byte, 0] &= np.uint64(0xFF)
        z[byte, 1:] = 0
        z[mid, 1] &= np.uint64((1 << 57) - 1)
        z[mid, 2:] = 0
    elif mode != "uniform":
        raise ValueError("witness mode must be 'uniform' or 'circom'")
    z[0] = (1, 0, 0, 0)
    return z


@dataclass
class Instance:
    name: str
    matrices: ConstraintMatrices  # canonical coefficients
    z_mont: np.ndarray            # (m, 4) Montgomery
    nc: int
    ni: int
    m: int
    n: int


def make_instance(ctx: ffi.Context, name: str, seed: int = 0xC0FFEE, witness: str = "uniform", **override) -> Instance:
    cfg = dict(CONFIGS[name]) if name in CONFIGS else {}
    cfg.update(override)
    nc, ni, m, mean = cfg["nc"], cfg["ni"], cfg["m"], cfg["mean"]
    A = _matrix(seed, 0x10, nc, m, mean[0], 1, 0)
    B = _matrix(seed, 0x20, nc, m, mean[1], 1, 0, absent_pct=35)  # 35 % of the wires absent from B
    Cr = _matrix(seed, 0x30, nc, m, mean[2], 0, 1)  # columns >= 1: wire 0 is reserved for the solved term
    z = ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, witness_canonical(seed, m, witness))
    # solve k0 on the GPU with the library's own sparse evaluation
    sctx = ffi.Context(ctx.device)
    try:
        sctx.load_r1cs(nc, ni, m, [A[0], B[0], Cr[0]], [A[1], B[1], Cr[1]], [A[2], B[2], Cr[2]], ffi.ENC_CANONICAL)
        az, bz, cz = sctx.r1cs_eval(z, nc)
    finally:
        sctx.close()
    k0 = ctx.field_op(ffi.FIELD_FR, ffi.OP_SUB, ctx.field_op(ffi.FIELD_FR, ffi.OP_MUL, az, bz), cz)
    k0 = ctx.field_op(ffi.FIELD_FR, ffi.OP_FROM_MONT, k0)
    starts = Cr[0][:-1].astype(np.int64)
    c_col = np.insert(Cr[1], starts, np.uint32(0))
    c_val = np.insert(Cr[2], starts, k0, axis=0)
    c_ptr = Cr[0] + np.arange(nc + 1, dtype=np.uint64)
    mats = ConstraintMatrices(ni, m - ni, nc, [A[0], B[0], c_ptr], [A[1], B[1], c_col],
                              [A[2], B[2], np.ascontiguousarray(c_val)], ffi.ENC_CANONICAL)
    n = 1
    while n < nc + ni:
        n <<= 1
    return Instance(name, mats, z, nc, ni, m, n)
