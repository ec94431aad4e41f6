"""CPU twin of crescent_credentials_b200/synth.py + generator.py built on the C++ oracle (coracle).

TEST INFRASTRUCTURE (lives under oracle/): imported only by tests/ and by bench.py's `--impl reference` arm, which must
prove the SAME instance under the SAME key as the GPU arm without touching the CUDA library.  The index / coefficient
streams are the product's own numpy generators (synth._matrix, synth.witness_canonical: no device code); every field or
group operation -- the solved constant-wire coefficients k0 = <A_i,z><B_i,z> - <C_i,z>, the Montgomery conversions and
the whole key (generator.rs:50-228 with a known trapdoor) -- is computed by oracle/libg16oracle.so.

The tests check that both builders return identical arrays (tests/test_gpu_fullsize.py), so "same_config" between the
two bench arms is a tested property, not a convention."""
from __future__ import annotations

import numpy as np

import coracle as c

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _mont(v: int) -> np.ndarray:
    x = (v % R_MOD) << 256
    x %= R_MOD
    return np.array([(x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)


def make_instance_cpu(name: str, seed: int = 0xC0FFEE, witness: str = "uniform", **override):
    """Same recipe, same streams, same arrays as synth.make_instance -- computed on the host cores."""
    from crescent_credentials_b200 import synth  # numpy stream generators only (no device call on this path)
    from crescent_credentials_b200.groth16 import ConstraintMatrices
    cfg = dict(synth.CONFIGS[name]) if name in synth.CONFIGS else {}
    cfg.update(override)
    nc, ni, m, mean = cfg["nc"], cfg["ni"], cfg["m"], cfg["mean"]
    A = synth._matrix(seed, 0x10, nc, m, mean[0], 1, 0)
    B = synth._matrix(seed, 0x20, nc, m, mean[1], 1, 0, absent_pct=35)
    Cr = synth._matrix(seed, 0x30, nc, m, mean[2], 0, 1)
    z = c.field_op(0, 5, synth.witness_canonical(seed, m, witness))
    val_m = [c.field_op(0, 5, M[2]) if len(M[2]) else M[2] for M in (A, B, Cr)]
    r1 = c.r1cs_struct(nc, ni, m, [A[0], B[0], Cr[0]], [A[1], B[1], Cr[1]], val_m)
    az, bz, cz = c.r1cs_eval(r1, z)
    k0 = c.field_op(0, 6, c.field_op(0, 2, c.field_op(0, 0, az, bz), cz))
    starts = Cr[0][:-1].astype(np.int64)
    c_col = np.insert(Cr[1], starts, np.uint32(0))
    c_val = np.insert(Cr[2], starts, k0, axis=0)
    c_ptr = Cr[0] + np.arange(nc + 1, dtype=np.uint64)
    mats = ConstraintMatrices(ni, m - ni, nc, [A[0], B[0], c_ptr], [A[1], B[1], c_col],
                              [A[2], B[2], np.ascontiguousarray(c_val)], 1)  # 1 = ENC_CANONICAL
    n = 1
    while n < nc + ni:
        n <<= 1
    return synth.Instance(name, mats, z, nc, ni, m, n)


def r1cs_of(inst):
    """coracle r1cs struct (Montgomery coefficients) of an Instance whose matrices are canonical."""
    mats = inst.matrices
    val = [c.field_op(0, 5, v) if len(v) else v for v in mats.val] if mats.encoding == 1 else mats.val
    return c.r1cs_struct(inst.nc, inst.ni, inst.m, mats.row_ptr, mats.col, val)


def generate_parameters_cpu(inst, td, r1=None):
    """generate_parameters_with_qap (forks/groth16/src/generator.rs:50-228) + LibsnarkReduction::instance_map_with_evaluation
    and h_query_scalars (r1cs_to_qap.rs:106-148,215-225) for a known trapdoor, canonical generators, Montgomery arrays
    in the layout of ffi.Context.load_pk / coracle.pk_struct.  Returns (arrays, qap)."""
    r1 = r1 or r1cs_of(inst)
    ni, m = inst.ni, inst.m
    a_w, b_w, c_w, zt_m, n = c.instance_map(r1, _mont(td.t))
    assert n == inst.n
    comb = c.field_op(0, 1, c.field_op(0, 1, c.field_op(0, 8, a_w, _mont(td.beta)), c.field_op(0, 8, b_w, _mont(td.alpha))), c_w)
    gi, di = pow(td.gamma, -1, R_MOD), pow(td.delta, -1, R_MOD)
    gamma_abc = c.field_op(0, 8, comb[:ni], _mont(gi))
    l = c.field_op(0, 8, comb[ni:], _mont(di)) if m > ni else np.zeros((0, 4), dtype=np.uint64)
    zt = (pow(td.t, n, R_MOD) - 1) % R_MOD
    hs = c.pow_table(_mont(td.t), _mont(zt * di), n - 1) if n > 1 else np.zeros((0, 4), dtype=np.uint64)
    singles = np.stack([_mont(td.alpha), _mont(td.beta), _mont(td.delta), _mont(td.gamma)])
    s1, s2 = c.fixed_base(1, singles), c.fixed_base(2, singles)
    arrays = dict(alpha_g1=s1[0].copy(), beta_g1=s1[1].copy(), delta_g1=s1[2].copy(), beta_g2=s2[1].copy(),
                  delta_g2=s2[2].copy(), a_query=c.fixed_base(1, a_w), b_g1_query=c.fixed_base(1, b_w),
                  b_g2_query=c.fixed_base(2, b_w), h_query=c.fixed_base(1, hs), l_query=c.fixed_base(1, l))
    qap = dict(a=a_w, b=b_w, c=c_w, l=l, hs=hs, zt=zt, n=n, gamma_abc=gamma_abc, gamma_g2=s2[3].copy(),
               gamma_abc_g1=c.fixed_base(1, gamma_abc))
    return arrays, qap
