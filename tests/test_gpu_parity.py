"""GPU parity tests: every call goes through the C ABI (libg16b200.so) and is compared bit-for-bit with the
big-integer oracle (oracle/pyref.py) and with the committed golden fixtures."""
import random

import numpy as np
import pytest

import pyref as o
from conftest import GOLDEN_NAMES, load_golden
from crescent_credentials_b200 import ffi
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200.r1cs import load_matrices

pytestmark = pytest.mark.gpu


def _rand_vals(p, n, seed):
    rnd = random.Random(seed)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (p - 1) // 2, (p + 1) // 2, (1 << 253) % p]
    return edge + [rnd.randrange(p) for _ in range(n - len(edge))]


@pytest.mark.parametrize("field,p", [(ffi.FIELD_FR, o.R_MOD), (ffi.FIELD_FQ, o.Q_MOD)])
def test_field_ops_bit_exact(gpu_ctx, field, p):
    n = 4096
    a = _rand_vals(p, n, 1)
    b = list(reversed(_rand_vals(p, n, 2)))
    A, B = g.ints_to_limbs(a), g.ints_to_limbs(b)
    rinv = pow(1 << 256, -1, p)
    got = g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_MUL, A, B))
    assert got == [x * y * rinv % p for x, y in zip(a, b)]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_ADD, A, B)) == [(x + y) % p for x, y in zip(a, b)]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_SUB, A, B)) == [(x - y) % p for x, y in zip(a, b)]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_NEG, A)) == [(-x) % p for x in a]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_SQR, A)) == [x * x * rinv % p for x in a]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_TO_MONT, A)) == [(x << 256) % p for x in a]
    assert g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_FROM_MONT, A)) == [x * rinv % p for x in a]
    nz = [x for x in a if x][:300]
    inv = g.limbs_to_ints(gpu_ctx.field_op(field, ffi.OP_INV, g.ints_to_limbs([(x << 256) % p for x in nz])))
    assert inv == [(pow(x, -1, p) << 256) % p for x in nz]


def test_fq2_ops_bit_exact(gpu_ctx):
    rnd = random.Random(5)
    n = 512
    a = [(rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD)) for _ in range(n)]
    b = [(rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD)) for _ in range(n)]
    enc = lambda v: g.fq_to_mont([c for x in v for c in x]).reshape(-1, 8)
    dec = lambda arr: [tuple(t) for t in np.array(g.fq_from_mont(arr.reshape(-1, 4)), dtype=object).reshape(-1, 2)]
    assert dec(gpu_ctx.field_op(ffi.FIELD_FQ2, ffi.OP_MUL, enc(a), enc(b))) == [o.Fq2.mul(x, y) for x, y in zip(a, b)]
    assert dec(gpu_ctx.field_op(ffi.FIELD_FQ2, ffi.OP_SQR, enc(a))) == [o.Fq2.sqr(x) for x in a]
    assert dec(gpu_ctx.field_op(ffi.FIELD_FQ2, ffi.OP_INV, enc(a))) == [o.Fq2.inv(x) for x in a]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 10, 11, 12, 13])
def test_ntt_matches_oracle(gpu_ctx, log_n):
    n = 1 << log_n
    vals = [o.stream_fr(0x17 + log_n, i) for i in range(n)]
    dom = o.Domain(n)
    x = g.fr_to_mont(vals)
    assert g.fr_from_mont(gpu_ctx.ntt(x)) == dom.fft(vals)
    assert g.fr_from_mont(gpu_ctx.ntt(x, inverse=True)) == dom.ifft(vals)
    assert g.fr_from_mont(gpu_ctx.ntt(x, coset=True)) == dom.coset_fft(vals)
    assert g.fr_from_mont(gpu_ctx.ntt(x, inverse=True, coset=True)) == dom.coset_ifft(vals)


@pytest.mark.parametrize("log_n", [4, 11, 14, 21, 22, 23])
def test_ntt_radix4_pass_equals_radix2_pass(gpu_ctx, log_n):
    """k_ntt_pass4 (two levels per shared-memory round trip) against k_ntt_pass (one level): identical bits for every
    transform kind, including the multi-pass sizes whose tiles have column bits (t_log > 0) and odd level counts."""
    n = 1 << log_n
    rng = np.random.default_rng(100 + log_n)
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    x = gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_FROM_MONT, x))
    try:
        for inverse in (False, True):
            for coset in (False, True):
                gpu_ctx.set_option("ntt_radix4", 0)
                want = gpu_ctx.ntt(x, inverse=inverse, coset=coset)
                gpu_ctx.set_option("ntt_radix4", 1)
                got = gpu_ctx.ntt(x, inverse=inverse, coset=coset)
                assert np.array_equal(got, want), (log_n, inverse, coset)
    finally:
        gpu_ctx.set_option("ntt_radix4", -1)  # library default (auto)


@pytest.mark.parametrize("log_n", [16, 20])
def test_ntt_round_trip_large(gpu_ctx, log_n):
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)  # < 2^254 < 4r: reduce by one Montgomery pass
    x = gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_TO_MONT, gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_FROM_MONT, x))
    for coset in (False, True):
        y = gpu_ctx.ntt(x, coset=coset)
        assert not np.array_equal(x, y)
        assert np.array_equal(gpu_ctx.ntt(y, inverse=True, coset=coset), x)
    # linearity: NTT(a + b) = NTT(a) + NTT(b)
    x2 = np.roll(x, 1, axis=0)
    lhs = gpu_ctx.ntt(gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_ADD, x, x2))
    rhs = gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_ADD, gpu_ctx.ntt(x), gpu_ctx.ntt(x2))
    assert np.array_equal(lhs, rhs)


def _points(curve, n, seed, with_inf=True):
    tbl = curve.fixed_base_table()
    pts = []
    for i in range(n):
        if with_inf and i % 7 == 3:
            pts.append(None)
        else:
            pts.append(curve.to_affine(curve.fixed_mul_j(tbl, o.stream_fr(seed, i) >> (200 if i % 2 else 0))))
    return pts


def _scalars(n, seed):
    out = []
    for i in range(n):
        k = i % 6
        v = o.stream_fr(seed, i)
        out.append([v, 0, 1, v & 0xFF, o.R_MOD - 1, v >> 130][k])
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 33, 300])
def test_msm_g1_matches_oracle(gpu_ctx, n):
    pts, sc = _points(o.G1, n, 21), _scalars(n, 22)
    out, inf = gpu_ctx.msm(1, g.g1_points_to_mont(pts), g.fr_to_mont(sc))
    assert g.g1_from_mont(out, inf) == o.G1.to_affine(o.G1.msm(pts, sc))


@pytest.mark.parametrize("n", [1, 40, 150])
def test_msm_g2_matches_oracle(gpu_ctx, n):
    pts, sc = _points(o.G2, n, 31), _scalars(n, 32)
    out, inf = gpu_ctx.msm(2, g.g2_points_to_mont(pts), g.fr_to_mont(sc))
    assert g.g2_from_mont(out, inf) == o.G2.to_affine(o.G2.msm(pts, sc))


def test_msm_cancels_to_infinity(gpu_ctx):
    p = o.G1.mul(o.G1_GEN, 12345)
    out, inf = gpu_ctx.msm(1, g.g1_points_to_mont([p, o.G1.neg(p), p]), g.fr_to_mont([5, 5, 0]))
    assert inf and g.g1_from_mont(out, inf) is None
    # same point many times: exercises the doubling branch of the mixed addition
    out, inf = gpu_ctx.msm(1, g.g1_points_to_mont([p] * 50), g.fr_to_mont([3] * 50))
    assert g.g1_from_mont(out, inf) == o.G1.mul(p, 150)


def test_fixed_base_matches_oracle(gpu_ctx):
    ks = [0, 1, 2, o.R_MOD - 1] + [o.stream_fr(77, i) for i in range(20)]
    got = gpu_ctx.fixed_base(1, g.fr_to_mont(ks))
    assert [g.g1_from_mont(r) for r in got] == [o.G1.mul(o.G1_GEN, k) if k else None for k in ks]
    got = gpu_ctx.fixed_base(2, g.fr_to_mont(ks[:10]))
    assert [g.g2_from_mont(r) for r in got] == [o.G2.mul(o.G2_GEN, k) if k else None for k in ks[:10]]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_prove(name):
    """R1CS file -> CSR, arkworks-serialised pk -> device, prove with injected (r, s): H coefficients, the five MSM
    outputs (through the closed form of A/B/C) and the serialised proof bytes must equal the golden fixture."""
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    assert (mats.num_instance_variables, mats.num_witness_variables, mats.num_constraints) == (
        meta["num_instance"], meta["num_witness"], meta["num_constraints"])
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    qap = g.CircomReduction if meta["reduction"] == "circom" else g.LibsnarkReduction
    prover = g.Groth16(0, qap)
    try:
        z = [int(v, 16) for v in meta["z"]]
        h = prover.witness_map_from_matrices(mats, mats.num_instance_variables, mats.num_constraints, z)
        assert h == [int(v, 16) for v in meta["h"]]
        proof = prover.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), mats,
                                                                mats.num_instance_variables, mats.num_constraints, z)
        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        assert proof.serialize_compressed().hex() == meta["proof_compressed"]
        # closed form in the exponent, independent of every MSM code path
        A, B, C = (int(v, 16) for v in meta["proof_dlog"])
        assert proof.a == o.G1.mul(o.G1_GEN, A) and proof.b == o.G2.mul(o.G2_GEN, B) and proof.c == o.G1.mul(o.G1_GEN, C)
        # second proof on the resident context (different r, s) still verifies in the exponent via linearity in r, s:
        t = prover.timings()
        assert t["total_ms"] > 0
    finally:
        prover.close()


@pytest.mark.parametrize("name", ["rand100", "rand300"])
def test_golden_msm_outputs(gpu_ctx, name):
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    z = [int(v, 16) for v in meta["z"]]
    h = [int(v, 16) for v in meta["h"]]
    ni = meta["num_instance"]
    to_m = lambda arr, fld: gpu_ctx.field_op(fld, ffi.OP_TO_MONT, arr)
    cases = {"h": (1, pk.arrays["h_query"], h), "l": (1, pk.arrays["l_query"], z[ni:]),
             "a": (1, pk.arrays["a_query"][1:], z[1:]), "b_g1": (1, pk.arrays["b_g1_query"][1:], z[1:]),
             "b_g2": (2, pk.arrays["b_g2_query"][1:], z[1:])}
    for key, (grp, pts, sc) in cases.items():
        pts_m = to_m(np.ascontiguousarray(pts).reshape(-1, 4), ffi.FIELD_FQ).reshape(pts.shape)
        out, inf = gpu_ctx.msm(grp, pts_m, g.fr_to_mont(sc))
        got = g.g1_from_mont(out, inf) if grp == 1 else g.g2_from_mont(out, inf)
        ser = g._ser_g1(got, False) if grp == 1 else g._ser_g2(got, False)
        assert ser.hex() == meta["msm"][key], key


def test_r1cs_eval_matches_oracle(gpu_ctx):
    meta, r1cs_bytes, _ = load_golden("rand300")
    mats = load_matrices(r1cs_bytes)
    m = mats.num_instance_variables + mats.num_witness_variables
    gpu_ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, m, mats.row_ptr, mats.col, mats.val, mats.encoding)
    z = [int(v, 16) for v in meta["z"]]
    az, bz, cz = gpu_ctx.r1cs_eval(g.fr_to_mont(z), mats.num_constraints)
    om = o.r1cs_to_matrices(o.read_r1cs(r1cs_bytes))
    assert g.fr_from_mont(az) == [o.evaluate_constraint(r, z) for r in om.a]
    assert g.fr_from_mont(bz) == [o.evaluate_constraint(r, z) for r in om.b]
    assert g.fr_from_mont(cz) == [o.evaluate_constraint(r, z) for r in om.c]
    # satisfied system: a*b == c row by row
    prod = gpu_ctx.field_op(ffi.FIELD_FR, ffi.OP_MUL, az, bz)
    assert np.array_equal(prod, cz)


def test_errors_are_loud(gpu_ctx):
    with pytest.raises(ffi.G16Error):
        gpu_ctx.ntt(np.zeros((3, 4), dtype=np.uint64))
    fresh = ffi.Context(0)
    try:
        with pytest.raises(ffi.G16Error):
            fresh.prove(np.zeros((4, 4), dtype=np.uint64), np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
        with pytest.raises(ffi.G16Error):  # column out of range
            fresh.load_r1cs(1, 1, 2, [np.array([0, 1], dtype=np.uint64)] * 3, [np.array([5], dtype=np.uint32)] * 3,
                            [np.zeros((1, 4), dtype=np.uint64)] * 3)
    finally:
        fresh.close()


def test_int_pipe_probe(gpu_ctx):
    for which in range(4):
        assert gpu_ctx.bench_int_pipe(which) > 1.0


@pytest.mark.parametrize("name", ["rand100", "rand300", "dummy924_nozk"])
def test_golden_prove_precomputed_windows(name):
    """Same golden proofs with the 2^(c*w) base tables (single bucket set, no Horner tail)."""
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    prover = g.Groth16(0, precompute=True)
    try:
        z = [int(v, 16) for v in meta["z"]]
        proof = prover.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), mats,
                                                                mats.num_instance_variables, mats.num_constraints, z)
        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
    finally:
        prover.close()


@pytest.mark.parametrize("c", [4, 7, 13, 18])
def test_msm_window_sizes(gpu_ctx, c):
    n = 200
    pts, sc = _points(o.G1, n, 41), _scalars(n, 42)
    gpu_ctx.set_option("window_bits", c)
    try:
        out, inf = gpu_ctx.msm(1, g.g1_points_to_mont(pts), g.fr_to_mont(sc))
    finally:
        gpu_ctx.set_option("window_bits", 0)
    assert g.g1_from_mont(out, inf) == o.G1.to_affine(o.G1.msm(pts, sc))


@pytest.mark.parametrize("levels", [0, 1, 2, 3, 8])
def test_msm_batched_affine_levels(gpu_ctx, levels):
    """Every split between batched-affine levels and the XYZZ tail gives the same group element (G1 and G2), including
    infinity points, repeated points (tangent case), P + (-P) pairs and one heavily loaded bucket (equal scalars)."""
    n = 260
    pts, sc = _points(o.G1, n, 51), _scalars(n, 52)
    pts[10] = pts[11] = pts[12] = pts[13]          # same point, and ...
    sc[10] = sc[11] = sc[12] = sc[13] = 0x1234567  # ... same digits: P + P inside a bucket at level 0 and 2P + 2P at level 1
    pts[21] = o.G1.neg(pts[20])
    sc[20] = sc[21] = 0xABCDEF                     # P + (-P) inside a bucket
    for i in range(100, 230):
        sc[i] = 3                                  # one bucket with 130 points: survivors go through the task path or the
                                                   # one-thread tail depending on the number of levels
    gpu_ctx.set_option("ba_levels", levels)
    try:
        out, inf = gpu_ctx.msm(1, g.g1_points_to_mont(pts), g.fr_to_mont(sc))
        assert g.g1_from_mont(out, inf) == o.G1.to_affine(o.G1.msm(pts, sc))
        n2 = 70
        p2, s2 = _points(o.G2, n2, 53), _scalars(n2, 54)
        p2[4] = p2[5] = p2[6]
        s2[4] = s2[5] = s2[6] = 0x7654321
        p2[9] = o.G2.neg(p2[8])
        s2[8] = s2[9] = 0x13579B
        for i in range(30, 50):
            s2[i] = 5
        out, inf = gpu_ctx.msm(2, g.g2_points_to_mont(p2), g.fr_to_mont(s2))
        assert g.g2_from_mont(out, inf) == o.G2.to_affine(o.G2.msm(p2, s2))
    finally:
        gpu_ctx.set_option("ba_levels", -1)


@pytest.mark.parametrize("share,levels", [(0, 0), (0, 5), (1, 0), (1, 5)])
def test_golden_prove_digit_sharing(share, levels):
    """l reusing a's digit stage and b_g2 reusing b_g1's (and the batched-affine levels) do not change a proof byte."""
    for name in ("rand300", "dummy924_nozk"):
        meta, r1cs_bytes, pk_bytes = load_golden(name)
        mats = load_matrices(r1cs_bytes)
        pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
        prover = g.Groth16(0, precompute=True)
        prover.ctx.set_option("share_digits", share)
        prover.ctx.set_option("ba_levels", levels)
        try:
            z = [int(v, 16) for v in meta["z"]]
            proof = prover.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), mats,
                                                                    mats.num_instance_variables, mats.num_constraints, z)
            assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        finally:
            prover.close()


@pytest.mark.parametrize("split,prio", [(0, 0), (1, 0), (0, 1), (1, 1), (1, 2), (1, 3)])
def test_golden_prove_stream_plans(split, prio):
    """Where the MSMs are queued (the digit-sharing MSM beside / behind the one that built the stage; witness map + h MSM on
    the high-priority stream) changes the order of execution only: same proof bytes, also when proofs run back to back."""
    for name in ("rand300", "dummy924_nozk", "silly"):
        meta, r1cs_bytes, pk_bytes = load_golden(name)
        mats = load_matrices(r1cs_bytes)
        pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
        prover = g.Groth16(0, precompute=True)
        prover.ctx.set_option("share_digits", 1)
        prover.ctx.set_option("split_chains", split)
        prover.ctx.set_option("wm_priority", min(prio, 2))   # 2: only the witness map on the high-priority stream
        prover.ctx.set_option("wm_first", int(prio == 3))     # 3: the wire chains start after the witness map
        try:
            z = [int(v, 16) for v in meta["z"]]
            for _ in range(3):
                proof = prover.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), mats,
                                                                        mats.num_instance_variables, mats.num_constraints, z)
                assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        finally:
            prover.close()


@pytest.mark.parametrize("tables", [0, 1])
def test_golden_prove_assembly_fixed_base_tables(tables):
    """The (r, s)-only points of the assembly from the per-key fixed-base tables (one scalar byte per lane + tree sum) or
    from single-lane double-and-add: same proof bytes for every fixture, including r = s = 0 (B-in-G1 skipped, prover.rs:102),
    and for (r, s) with zero and all-ones bytes."""
    for name in GOLDEN_NAMES:
        meta, r1cs_bytes, pk_bytes = load_golden(name)
        mats = load_matrices(r1cs_bytes)
        pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
        prover = g.Groth16(0, qap=g.CircomReduction if meta["reduction"] == "circom" else g.LibsnarkReduction)
        prover.ctx.set_option("asm_tables", tables)
        try:
            z = [int(v, 16) for v in meta["z"]]
            proof = prover.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), mats,
                                                                    mats.num_instance_variables, mats.num_constraints, z)
            assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"], name
            if name == "rand100":  # scalars with empty / full bytes, r = 0 with s != 0 and the reverse: against the other variant
                from crescent_credentials_b200 import verifier as v
                other = g.Groth16(0)
                other.ctx.set_option("asm_tables", 1 - tables)
                ver = v.Verifier(0)
                try:
                    pvk = ver.prepare_verifying_key(pk)
                    public = z[1:mats.num_instance_variables]
                    for r, s in [(0, 5), (7, 0), (1 << 248, (1 << 253) + 255), (o.R_MOD - 1, 0xFF00FF00FF << 100), (0xFF, 1 << 8)]:
                        args = (pk, r, s, mats, mats.num_instance_variables, mats.num_constraints, z)
                        p1 = prover.create_proof_with_reduction_and_matrices(*args)
                        assert p1.serialize_uncompressed() == other.create_proof_with_reduction_and_matrices(*args).serialize_uncompressed()
                        assert ver.verify_proof(pvk, p1, public) is True  # any (r, s) gives a valid proof of the same statement
                finally:
                    other.close()
                    ver.close()
        finally:
            prover.close()


def test_spmv_sliced_ell_ragged_rows(gpu_ctx):
    """Sliced-ELL SpMV (rows sorted by length, 32 per slice, +1 / -1 classes in the column word) against the oracle's
    evaluate_constraint and against the row-per-thread CSR kernel: empty rows, rows longer than a slice, repeated wires in
    a row, +1 / -1 / 0 / general coefficients, a row count that is not a multiple of 32; natural and bit-reversed order."""
    rnd = random.Random(77)
    R = o.R_MOD
    m, nc, ni = 257, 101, 3
    lens = [0, 1, 2, 33, 300, 64, 0, 5] + [rnd.choice([0, 1, 2, 3, 4, 7, 40]) for _ in range(nc - 8)]
    pool = [1, R - 1, 0, 2, R - 2, 1 << 120, rnd.randrange(R), rnd.randrange(R)]
    rows = [[(rnd.choice(pool) if rnd.random() < 0.8 else rnd.randrange(R), rnd.randrange(m)) for _ in range(L)] for L in lens]
    ptr = np.zeros(nc + 1, dtype=np.uint64)
    ptr[1:] = np.cumsum(lens)
    col = np.array([c for r in rows for _, c in r], dtype=np.uint32)
    val = g.fr_to_mont([v for r in rows for v, _ in r]).reshape(-1, 4)
    z = [1] + [rnd.randrange(R) for _ in range(m - 1)]
    zm = g.fr_to_mont(z)
    want = [o.evaluate_constraint(r, z) for r in rows]
    ctx = ffi.Context(0)
    try:
        ctx.load_r1cs(nc, ni, m, [ptr] * 3, [col] * 3, [val] * 3, ffi.ENC_MONTGOMERY)
        for sell in (1, 0):
            ctx.set_option("spmv_sell", sell)
            az, bz, cz = ctx.r1cs_eval(zm, nc)
            assert g.fr_from_mont(az) == want and g.fr_from_mont(bz) == want and g.fr_from_mont(cz) == want, sell
        # the witness map places rows bit-reversed and pads: both reductions, both kernels, same h
        for red in (ffi.REDUCTION_LIBSNARK, ffi.REDUCTION_CIRCOM):
            ctx.set_option("spmv_sell", 1)
            h1 = ctx.witness_map(zm, red)
            ctx.set_option("spmv_sell", 0)
            h0 = ctx.witness_map(zm, red)
            assert np.array_equal(h0, h1), red
    finally:
        ctx.close()


# ---- row f-2: the key generator against the committed key bytes ------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_generator_reproduces_golden_key_bytes(gpu_ctx, name):
    """generate_parameters_with_qap on the GPU (forks/groth16/src/generator.rs:50-228; gamma = 1 and vk.delta_g1 as in the fork;
    both h_query conventions: r1cs_to_qap.rs:215-225 and qap.rs:92-107) serialises to the committed arkworks-layout key, byte for
    byte -- the same bytes the oracle's generator produced."""
    import ast
    from crescent_credentials_b200 import generator
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    td = ast.literal_eval(meta["trapdoor"]) if isinstance(meta["trapdoor"], str) else meta["trapdoor"]
    td = generator.Trapdoor(*(int(td[k], 16) for k in ("alpha", "beta", "gamma", "delta", "t")))
    pk, _ = generator.generate_parameters_with_qap(gpu_ctx, mats, td, reduction=meta["reduction"])
    assert pk.serialize_uncompressed(gpu_ctx) == pk_bytes
    assert pk.serialize_uncompressed() == pk_bytes   # same bytes through the integer marshalling path


def test_generate_random_parameters_follows_the_reference_draw_order(gpu_ctx):
    """generate_random_parameters_with_reduction (generator.rs:19-47,93): alpha, beta, delta, then t from the same StdRng stream;
    gamma = 1.  The key equals generate_parameters_with_qap under the toxic waste drawn by hand, and a proof under it verifies."""
    from crescent_credentials_b200 import generator, rng, verifier
    meta, r1cs_bytes, _ = load_golden("rand100")
    mats = load_matrices(r1cs_bytes)
    pk = generator.generate_random_parameters_with_reduction(gpu_ctx, mats, rng.StdRng.seed_from_u64(42))
    r2 = rng.StdRng.seed_from_u64(42)
    alpha, beta, delta = g.sample_fr(r2), g.sample_fr(r2), g.sample_fr(r2)
    t = generator.sample_element_outside_domain(r2, 128)
    want, _ = generator.generate_parameters_with_qap(gpu_ctx, mats, generator.Trapdoor(alpha, beta, 1, delta, t))
    assert pk.serialize_uncompressed(gpu_ctx) == want.serialize_uncompressed(gpu_ctx)
    assert g1_gamma_is_generator(pk)
    prover, ver = g.Groth16(0), verifier.Verifier(0)
    try:
        z = [int(v, 16) for v in meta["z"]]
        proof = prover.create_random_proof_with_reduction(pk, mats, mats.num_instance_variables, mats.num_constraints, z, r2)
        assert ver.verify(pk, z[1:mats.num_instance_variables], proof) is True
        assert ver.verify(pk, [z[1] + 1] + z[2:mats.num_instance_variables], proof) is False
    finally:
        prover.close()
        ver.close()


def g1_gamma_is_generator(pk) -> bool:
    """gamma = 1 (generator.rs:28): vk.gamma_g2 is the canonical G2 generator."""
    return g.g2_from_mont(np.asarray(pk.gamma_g2)) == o.G2_GEN
