"""Groth16 key generation on the GPU ("next" row f-2 of SURVEY section 8): the arithmetic of
generate_parameters_with_qap (forks/groth16/src/generator.rs:50-228) and
LibsnarkReduction::instance_map_with_evaluation (forks/groth16/src/r1cs_to_qap.rs:106-148) expressed with the
library's kernels -- Lagrange coefficients as one inverse NTT of the powers of t, the column sums of A, B, C as sparse
mat-vecs over the transposed matrices, and every query as fixed-base multiples of the canonical generators
(generator.rs:34-35; gamma is whatever the caller passes, the fork uses 1 at generator.rs:28).

Entry points, named as in the reference:
    generate_random_parameters_with_reduction(ctx, matrices, rng, reduction)   generator.rs:19-47  (alpha, beta, delta <- rng,
        gamma = 1 as in the fork, t <- sample_element_outside_domain(rng), canonical generators)
    generate_parameters_with_qap(ctx, matrices, trapdoor, reduction)            generator.rs:50-228 with the toxic waste given

Besides serving `zksetup`-style callers it lets tests and the bench mint *real* proving keys with a known trapdoor at full
rs256 scale (proofs are then checkable in the exponent); it is not on the prove hot path."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import ffi
from .groth16 import ConstraintMatrices, ProvingKey, R_MOD, fr_to_mont, sample_fr


@dataclass
class Trapdoor:
    alpha: int
    beta: int
    gamma: int
    delta: int
    t: int


def transpose_csr(nrows: int, ncols: int, row_ptr: np.ndarray, col: np.ndarray, val: np.ndarray):
    """CSR (nrows x ncols) -> CSR of the transpose (ncols x nrows)."""
    lens = np.diff(row_ptr.astype(np.int64))
    rows = np.repeat(np.arange(nrows, dtype=np.uint32), lens)
    order = np.argsort(col, kind="stable")
    t_ptr = np.zeros(ncols + 1, dtype=np.uint64)
    t_ptr[1:] = np.cumsum(np.bincount(col, minlength=ncols)).astype(np.uint64)
    return t_ptr, rows[order], np.ascontiguousarray(val[order])


def domain_size(nc: int, ni: int) -> int:
    n = 1
    while n < nc + ni:
        n <<= 1
    return n


def generate_parameters_with_qap(ctx: ffi.Context, m: ConstraintMatrices, td: Trapdoor, reduction: str = "libsnark"):
    """Returns (ProvingKey, qap) where qap holds the Montgomery scalar vectors behind every query
    (a, b, c per wire, l, h scalars, zt, n) for checks in the exponent."""
    nc, ni = m.num_constraints, m.num_instance_variables
    wires = ni + m.num_witness_variables
    n = domain_size(nc, ni)
    mont = lambda v: fr_to_mont([v % R_MOD])[0]
    one = mont(1)
    # Lagrange coefficients at t: u = iNTT([t^k])  (== evaluate_all_lagrange_coefficients, r1cs_to_qap.rs:118)
    u = ctx.ntt(ctx.pow_table(mont(td.t), one, n), inverse=True)
    zt = (pow(td.t, n, R_MOD) - 1) % R_MOD
    # column sums over the transposed matrices: a_w = sum_i u_i * A[i][w]   (r1cs_to_qap.rs:135-145)
    tp, tc, tv = [], [], []
    for k in range(3):
        p, c, v = transpose_csr(nc, wires, m.row_ptr[k], m.col[k], m.val[k])
        tp.append(p)
        tc.append(c)
        tv.append(v)
    tctx = ffi.Context(ctx.device)
    try:
        tctx.load_r1cs(wires, 1, n, tp, tc, tv, m.encoding)
        a_w, b_w, c_w = tctx.r1cs_eval(u, wires)
    finally:
        tctx.close()
    a_w[:ni] = ctx.field_op(ffi.FIELD_FR, ffi.OP_ADD, a_w[:ni], u[nc:nc + ni])  # r1cs_to_qap.rs:128-133
    F = ffi.FIELD_FR
    comb = ctx.field_op(F, ffi.OP_ADD, ctx.field_op(F, ffi.OP_ADD, ctx.field_op(F, ffi.OP_MUL_BCAST, a_w, mont(td.beta)),
                                                    ctx.field_op(F, ffi.OP_MUL_BCAST, b_w, mont(td.alpha))), c_w)
    gi, di = pow(td.gamma, -1, R_MOD), pow(td.delta, -1, R_MOD)
    gamma_abc = ctx.field_op(F, ffi.OP_MUL_BCAST, comb[:ni], mont(gi))      # generator.rs:113-117
    l = ctx.field_op(F, ffi.OP_MUL_BCAST, comb[ni:], mont(di)) if wires > ni else np.zeros((0, 4), dtype=np.uint64)
    if reduction == "libsnark":                                             # r1cs_to_qap.rs:215-225, m_raw - 1 powers
        hs = ctx.pow_table(mont(td.t), mont(zt * di), n - 1) if n > 1 else np.zeros((0, 4), dtype=np.uint64)
    else:                                                                   # qap.rs:92-107
        sc = np.zeros((2 * n, 4), dtype=np.uint64)
        sc[:2 * (n - 1) + 1] = ctx.pow_table(mont(td.t), mont(di), 2 * (n - 1) + 1)
        hs = np.ascontiguousarray(ctx.ntt(sc, inverse=True)[1::2])
    g1 = lambda s: ctx.fixed_base(1, s)
    g2 = lambda s: ctx.fixed_base(2, s)
    singles = np.stack([mont(td.alpha), mont(td.beta), mont(td.delta), mont(td.gamma)])
    s1, s2 = g1(singles), g2(singles)
    arrays = dict(alpha_g1=s1[0].copy(), beta_g1=s1[1].copy(), delta_g1=s1[2].copy(), beta_g2=s2[1].copy(), delta_g2=s2[2].copy(),
                  a_query=g1(a_w), b_g1_query=g1(b_w), b_g2_query=g2(b_w), h_query=g1(hs), l_query=g1(l))
    pk = ProvingKey(arrays, ffi.ENC_MONTGOMERY)
    pk.gamma_g2 = s2[3].copy()
    pk.gamma_abc_g1 = g1(gamma_abc)
    qap = dict(a=a_w, b=b_w, c=c_w, l=l, hs=hs, zt=zt, n=n, gamma_abc=gamma_abc)
    return pk, qap


def sample_element_outside_domain(rng, n: int) -> int:
    """EvaluationDomain::sample_element_outside_domain (ark-poly 0.4): draw Fr::rand until Z(t) = t^n - 1 != 0."""
    while True:
        t = sample_fr(rng)
        if pow(t, n, R_MOD) != 1:
            return t


def generate_random_parameters_with_reduction(ctx: ffi.Context, m: ConstraintMatrices, rng, reduction: str = "libsnark"):
    """Groth16::generate_random_parameters_with_reduction (forks/groth16/src/generator.rs:19-47): alpha, beta, delta are drawn
    from `rng` in that order, gamma is the constant 1 (fork, :28), the group generators are the canonical ones (:34-35), and t is
    drawn afterwards by sample_element_outside_domain (:93).  `rng`: rng.StdRng (the rand 0.8 mirror) for a run that follows a
    seeded reference run, or any CSPRNG exposing getrandbits(64).  Returns the ProvingKey; the toxic waste is dropped."""
    alpha = sample_fr(rng)
    beta = sample_fr(rng)
    delta = sample_fr(rng)
    n = domain_size(m.num_constraints, m.num_instance_variables)
    if n > (1 << 28):
        raise ffi.PolynomialDegreeTooLarge(ffi.ERR_DEGREE_TOO_LARGE, "no evaluation domain of that size (generator.rs:92)")
    t = sample_element_outside_domain(rng, n)
    pk, _ = generate_parameters_with_qap(ctx, m, Trapdoor(alpha, beta, 1, delta, t), reduction)
    return pk
