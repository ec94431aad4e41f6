"""Full-size bit-parity of the hot path against the C++ oracle: BASELINE.json's configurations, not their
2^12 twins.  For S-rs256 (uniform AND circom-like witness) and S-mdl1, with the key minted at full size:

  (a) all n coefficients of h from g16_witness_map            == coracle.witness_map    (r1cs_to_qap.rs:150-213)
  (b) each of the five MSM outputs of the proof's own configuration (window tables, shared digit stages, batched-affine
      levels) read back from the prover's partial sums          == coracle.msm over the FULL length (prover.rs:63-74,256-274)
  (c) the 256 serialised proof bytes                           == coracle.prove           (prover.rs:54-136)
  (d) the same system built on the host cores (oracle/refsynth.py, what `bench.py --impl reference` proves) is the same
      arrays, and the GPU-minted key equals the CPU-minted one byte for byte
  (e) size-independent properties: the QAP identity at the trapdoor point ties h to the witness; the proof satisfies the
      verification equation in the exponent; MSM linearity at full length.

Everything on the GPU side goes through the C ABI (ffi.Context)."""
import numpy as np
import pytest

import coracle as c
import pyref as o
import refsynth
from crescent_credentials_b200 import ffi, generator, synth
from crescent_credentials_b200 import groth16 as g

pytestmark = pytest.mark.gpu

R = o.R_MOD
TD = generator.Trapdoor(alpha=0x1234567 + 11, beta=0x89ABCDE + 13, gamma=1, delta=0xFEDCBA9 + 17, t=0x5EED0000C0FFEE0000BEEF + 19)
R_INT = 0x1111222233334444555566667777888899990000AAAABBBBCCCCDDDD % R
S_INT = 0x0F0E0D0C0B0A09080706050403020100FFEEDDCCBBAA9988 % R

CASES = [("S-rs256", "uniform"), ("S-rs256", "circom"), ("S-mdl1", "uniform")]


class Full:
    """One full-size instance, its GPU-minted key, a loaded prover context (the bench's configuration) and the oracle's view."""

    def __init__(self, workload, witness):
        self.workload, self.witness = workload, witness
        self.ctx = ffi.Context(0)
        self.inst = synth.make_instance(self.ctx, workload, witness=witness)
        self.pk, self.qap = generator.generate_parameters_with_qap(self.ctx, self.inst.matrices, TD)
        m = self.inst.matrices
        self.ctx.load_r1cs(self.inst.nc, self.inst.ni, self.inst.m, m.row_ptr, m.col, m.val, m.encoding)
        self.ctx.load_pk(self.pk.arrays, self.pk.encoding, 0, 1, True)
        self.r1 = refsynth.r1cs_of(self.inst)
        self.r_m, self.s_m = g.fr_to_mont([R_INT])[0], g.fr_to_mont([S_INT])[0]
        self._h = None

    @property
    def h_oracle(self):
        if self._h is None:
            self._h = c.witness_map(self.r1, self.inst.z_mont, self.inst.n)
        return self._h

    def close(self):
        self.ctx.close()


@pytest.fixture(scope="module", params=CASES, ids=lambda p: f"{p[0]}-{p[1]}")
def full(request):
    f = Full(*request.param)
    yield f
    f.close()


def _xyzz_g1(words):
    """16 u64 words (x, y, zz, zzz Montgomery) -> affine tuple / None."""
    x, y, zz, zzz = g.fq_from_mont(np.asarray(words, dtype=np.uint64).reshape(4, 4))
    if zz == 0:
        return None
    q = o.Q_MOD
    return (x * pow(zz, -1, q) % q, y * pow(zzz, -1, q) % q)


def _xyzz_g2(words):
    v = g.fq_from_mont(np.asarray(words, dtype=np.uint64).reshape(8, 4))
    x, y, zz, zzz = (v[0], v[1]), (v[2], v[3]), (v[4], v[5]), (v[6], v[7])
    if zz == (0, 0):
        return None
    return (o.Fq2.mul(x, o.Fq2.inv(zz)), o.Fq2.mul(y, o.Fq2.inv(zzz)))


def test_witness_map_all_coefficients(full):
    """(a) every one of the n coefficients of h, bit for bit; h[n-1] == 0 for a satisfied system (generator.rs:178)."""
    h_gpu = full.ctx.witness_map(full.inst.z_mont)
    assert h_gpu.shape == full.h_oracle.shape == (full.inst.n, 4)
    assert np.array_equal(h_gpu, full.h_oracle)
    assert not h_gpu[-1].any()


def test_five_msm_outputs(full):
    """(b) the proof's own five MSMs (tables, shared digits, batched-affine levels) against the oracle over the full length."""
    inst, pk = full.inst, full.pk
    part = full.ctx.prove_shard(inst.z_mont, full.r_m, full.s_m)
    got = {"h": _xyzz_g1(part[0:16]), "l": _xyzz_g1(part[16:32]), "a": _xyzz_g1(part[32:48]), "sa": _xyzz_g1(part[48:64]),
           "rb1": _xyzz_g1(part[64:80]), "b2": _xyzz_g2(part[80:112])}
    z = inst.z_mont
    want_h = g.g1_from_mont(c.msm(1, pk.arrays["h_query"], full.h_oracle))
    want_l = g.g1_from_mont(c.msm(1, pk.arrays["l_query"], z[inst.ni:]))
    want_a = g.g1_from_mont(c.msm(1, pk.arrays["a_query"][1:], z[1:]))
    want_b1 = g.g1_from_mont(c.msm(1, pk.arrays["b_g1_query"][1:], z[1:]))
    want_b2 = g.g2_from_mont(c.msm(2, pk.arrays["b_g2_query"][1:], z[1:]))
    assert got["h"] == want_h
    assert got["l"] == want_l
    assert got["a"] == want_a
    assert got["b2"] == want_b2
    assert got["sa"] == o.G1.mul(want_a, S_INT)     # s * MSM_a, scaled per rank (prover.rs:98 by linearity)
    assert got["rb1"] == o.G1.mul(want_b1, R_INT)
    st = full.ctx.msm_stats(0)
    assert st["points"] == len(pk.arrays["h_query"]) and st["levels"] > 0  # the batched-affine path really ran


def test_proof_bytes(full):
    """(c) serialised proof == oracle's, through the public call (host witness in, proof out) and the resident call."""
    inst = full.inst
    ref, _, _ = c.prove(c.pk_struct(full.pk.arrays), full.r1, inst.z_mont, full.r_m, full.s_m)
    want = g.Proof(g.g1_from_mont(ref[0]), g.g2_from_mont(ref[1]), g.g1_from_mont(ref[2])).serialize_uncompressed()
    got = g.Proof.from_ffi(full.ctx.prove(inst.z_mont, full.r_m, full.s_m))
    assert got.serialize_uncompressed() == want
    assert g.Proof.from_ffi(full.ctx.prove_resident(full.r_m, full.s_m)).serialize_uncompressed() == want
    # (e) the verification equation in the exponent, with h(t) Z(t) taken from the QAP identity -- never from the GPU's h
    z = g.fr_from_mont(inst.z_mont)
    dot = lambda u, v: sum(x * y for x, y in zip(u, v)) % R
    za, zb, zc = dot(z, g.fr_from_mont(full.qap["a"])), dot(z, g.fr_from_mont(full.qap["b"])), dot(z, g.fr_from_mont(full.qap["c"]))
    zl = dot(z[inst.ni:], g.fr_from_mont(full.qap["l"]))
    di = pow(TD.delta, -1, R)
    A = (TD.alpha + za + R_INT * TD.delta) % R
    B = (TD.beta + zb + S_INT * TD.delta) % R
    C = (zl + (za * zb - zc) * di + S_INT * A + R_INT * B - R_INT * S_INT % R * TD.delta) % R
    assert got.a == o.G1.mul(o.G1_GEN, A) and got.b == o.G2.mul(o.G2_GEN, B) and got.c == o.G1.mul(o.G1_GEN, C)
    # the QAP identity itself, on the oracle-checked h:  (sum z a)(sum z b) - sum z c == h(t) Z(t) == delta * sum h_i hs_i
    assert (za * zb - zc) % R == TD.delta * dot(g.fr_from_mont(full.h_oracle), g.fr_from_mont(full.qap["hs"])) % R


def test_cpu_built_system_is_the_same_system(full):
    """(d) oracle/refsynth.py (the reference arm's builder) returns the arrays of synth.make_instance, and the key minted on
    the host cores equals the GPU generator's (f-2 parity at full size) -- checked on S-rs256/uniform only (one CPU key)."""
    inst = full.inst
    if (full.workload, full.witness) != ("S-rs256", "uniform"):
        pytest.skip("one configuration is enough: the CPU key takes ~20 s")
    cpu = refsynth.make_instance_cpu(inst.name, witness="uniform")
    assert np.array_equal(cpu.z_mont, inst.z_mont)
    for k in range(3):
        assert np.array_equal(cpu.matrices.row_ptr[k], inst.matrices.row_ptr[k])
        assert np.array_equal(cpu.matrices.col[k], inst.matrices.col[k])
        assert np.array_equal(cpu.matrices.val[k], inst.matrices.val[k])
    arrays, qap = refsynth.generate_parameters_cpu(cpu, TD, full.r1)
    for name, arr in arrays.items():
        assert np.array_equal(np.asarray(arr).reshape(-1), np.asarray(full.pk.arrays[name]).reshape(-1)), name
    assert np.array_equal(qap["gamma_abc_g1"], full.pk.gamma_abc_g1)


def test_msm_linearity_full_length(full):
    """(e) MSM(s1) + MSM(s2) == MSM(s1 + s2) at full length through the stand-alone entry point (no tables, own window choice)."""
    inst, pk = full.inst, full.pk
    if (full.workload, full.witness) != ("S-rs256", "circom"):
        pytest.skip("one configuration is enough")
    pts = pk.arrays["l_query"]
    s1 = inst.z_mont[inst.ni:]
    s2 = np.ascontiguousarray(np.roll(s1, 7, axis=0))
    s12 = full.ctx.field_op(ffi.FIELD_FR, ffi.OP_ADD, s1, s2)
    p1 = g.g1_from_mont(*full.ctx.msm(1, pts, s1))
    p2 = g.g1_from_mont(*full.ctx.msm(1, pts, s2))
    p12 = g.g1_from_mont(*full.ctx.msm(1, pts, s12))
    assert o.G1.add(p1, p2) == p12
    assert p1 == g.g1_from_mont(c.msm(1, pts, s1))
