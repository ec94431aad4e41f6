"""CPU tests: pin the oracles to every byte-level golden the reference tree holds for this path, to each other, and to
the committed fixtures.  (The reference holds no golden proof / H vector / MSM output -- SURVEY 8c.)"""
import json
import os
import random

import numpy as np
import pytest

import coracle as c
import pyref as o
from conftest import GOLDEN, GOLDEN_NAMES, load_golden
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200.r1cs import R1CSFile, load_matrices

# forks/circom-compat/src/zkey.rs:397-402  (snarkjs curve.G1.F.one, Montgomery LE)
FQ_ONE_BUF = bytes([157, 13, 143, 197, 141, 67, 93, 211, 61, 11, 199, 245, 40, 235, 120, 10, 44, 70, 121, 120, 111, 163, 110,
                    102, 47, 223, 7, 154, 193, 119, 10, 14])
# zkey.rs:408-415 (G1 generator, Montgomery LE x||y)
G1_BUF = bytes([157, 13, 143, 197, 141, 67, 93, 211, 61, 11, 199, 245, 40, 235, 120, 10, 44, 70, 121, 120, 111, 163, 110, 102,
                47, 223, 7, 154, 193, 119, 10, 14, 58, 27, 30, 139, 27, 135, 186, 166, 123, 22, 142, 235, 81, 214, 241, 20, 88,
                140, 242, 240, 222, 70, 221, 204, 94, 190, 15, 52, 131, 239, 20, 28])
# zkey.rs:421-431 (G2 generator, Montgomery LE x.c0||x.c1||y.c0||y.c1)
G2_BUF = bytes([38, 32, 188, 2, 209, 181, 131, 142, 114, 1, 123, 73, 53, 25, 235, 220, 223, 26, 129, 151, 71, 38, 184, 251, 59,
                80, 150, 175, 65, 56, 87, 25, 64, 97, 76, 168, 125, 115, 180, 175, 196, 216, 2, 88, 90, 221, 67, 96, 134, 47,
                160, 82, 252, 80, 233, 9, 107, 123, 234, 58, 131, 240, 254, 20, 246, 233, 107, 136, 157, 250, 157, 97, 120, 155,
                158, 245, 151, 210, 127, 254, 254, 125, 27, 35, 98, 26, 158, 255, 6, 66, 158, 174, 235, 126, 253, 40, 238, 86,
                24, 199, 86, 91, 9, 100, 187, 60, 125, 50, 34, 249, 87, 220, 118, 16, 53, 51, 190, 53, 249, 85, 130, 100, 253,
                147, 230, 160, 164, 13])
# forks/circom-compat/src/circom/r1cs_reader.rs:266-318 (the reader's worked example)
R1CS_SAMPLE_HEX = """
72316373 01000000 03000000 01000000 40000000 00000000 20000000
010000f0 93f5e143 9170b979 48e83328 5d588181 b64550b8 29a031e1 724e6430
07000000 01000000 02000000 03000000 e8030000 00000000 03000000
02000000 88020000 00000000
02000000
05000000 03000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
06000000 08000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000
00000000 02000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
02000000 14000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000 0C000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
02000000
00000000 05000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
02000000 07000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000
01000000 04000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
04000000 08000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
05000000 03000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
02000000
03000000 2C000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
06000000 06000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
00000000
01000000
06000000 04000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000
00000000 06000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
02000000 0B000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000 05000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
01000000
06000000 58020000 00000000 00000000 00000000 00000000 00000000 00000000 00000000
03000000 38000000 00000000
00000000 00000000 03000000 00000000 0a000000 00000000 0b000000 00000000 0c000000 00000000 0f000000 00000000
44010000 00000000
"""


def test_reference_byte_goldens_montgomery_encodings():
    assert o.mont_le_bytes(1, o.Q_MOD) == FQ_ONE_BUF
    assert o.mont_le_bytes(o.G1_GEN[0], o.Q_MOD) + o.mont_le_bytes(o.G1_GEN[1], o.Q_MOD) == G1_BUF
    (x0, x1), (y0, y1) = o.G2_GEN
    assert b"".join(o.mont_le_bytes(v, o.Q_MOD) for v in (x0, x1, y0, y1)) == G2_BUF
    # the host-side marshalling of the product uses the same encoding
    assert g.g1_points_to_mont([o.G1_GEN]).tobytes() == G1_BUF
    assert g.g2_points_to_mont([o.G2_GEN]).tobytes() == G2_BUF
    # and the C++ oracle agrees (to_mont of the canonical coordinates)
    canon = g.ints_to_limbs([o.G1_GEN[0], o.G1_GEN[1], x0, x1, y0, y1])
    assert c.field_op(1, 5, canon).tobytes() == G1_BUF + G2_BUF
    assert o.G1.is_on_curve(o.G1_GEN) and o.G2.is_on_curve(o.G2_GEN)
    assert o.G1.mul(o.G1_GEN, o.R_MOD - 1) == o.G1.neg(o.G1_GEN)   # group order is r
    assert o.G2.mul(o.G2_GEN, o.R_MOD - 1) == o.G2.neg(o.G2_GEN)


@pytest.mark.parametrize("reader", ["oracle", "product"])
def test_reference_r1cs_sample(reader):
    data = bytes.fromhex("".join(R1CS_SAMPLE_HEX.split()))
    if reader == "oracle":
        r = o.read_r1cs(data)
        hdr = (r["version"], r["field_size"], r["n_wires"], r["n_pub_out"], r["n_pub_in"], r["n_prv_in"], r["n_labels"],
               r["n_constraints"])
        cons, wmap, prime = r["constraints"], r["wire_mapping"], r["prime"]
    else:
        f = R1CSFile(data)
        hdr = (f.version, f.field_size, f.n_wires, f.n_pub_out, f.n_pub_in, f.n_prv_in, f.n_labels, f.n_constraints)
        cons, wmap, prime = list(f.constraints()), list(f.wire_mapping), f.prime_size
    # the assertions of r1cs_reader.rs:320-344
    assert hdr == (1, 32, 7, 1, 2, 3, 0x03E8, 3)
    assert prime == bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")
    assert len(cons) == 3 and len(cons[0][0]) == 2
    assert cons[0][0][0] == (5, 3)
    assert cons[2][1][0] == (0, 6)
    assert len(cons[1][2]) == 0
    assert len(wmap) == 7 and wmap[1] == 3


def test_r1cs_reader_error_conventions():
    data = bytearray(bytes.fromhex("".join(R1CS_SAMPLE_HEX.split())))
    for mutate, msg in ((lambda d: d.__setitem__(0, 0x00), "Invalid magic number"),
                        (lambda d: d.__setitem__(4, 0x02), "Unsupported version"),
                        (lambda d: d.__setitem__(28, 0x02), "only supports bn256")):
        d = bytearray(data)
        mutate(d)
        with pytest.raises(ValueError, match=msg):
            R1CSFile(bytes(d))
        with pytest.raises(ValueError, match=msg):
            o.read_r1cs(bytes(d))


def test_constants_rederived():
    """SURVEY appendix constants, re-derived; these are also hard-coded in csrc/fp.cuh and oracle/g16_oracle.cpp."""
    for p, inv32, inv64 in ((o.R_MOD, 0xEFFFFFFF, 0xC2E1F593EFFFFFFF), (o.Q_MOD, 0xE4866389, 0x87D20782E4866389)):
        assert (-pow(p, -1, 1 << 32)) % (1 << 32) == inv32
        assert (-pow(p, -1, 1 << 64)) % (1 << 64) == inv64
    assert (1 << 256) % o.R_MOD == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert (1 << 256) % o.Q_MOD == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert o.FR_ROOT_2_28 == 0x2A3C09F0A58A7E8500E0A7EB8EF62ABC402D111E41112ED49BD61B6E725B19F0
    assert pow(o.FR_ROOT_2_28, 1 << 28, o.R_MOD) == 1 and pow(o.FR_ROOT_2_28, 1 << 27, o.R_MOD) != 1
    assert (o.R_MOD - 1) % (1 << 28) == 0 and ((o.R_MOD - 1) >> 28) % 2 == 1   # two-adicity 28
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "crescent_credentials_b200", "csrc", "fp.cuh")).read()
    for v in ((1 << 256) % o.R_MOD, (1 << 512) % o.R_MOD, (1 << 256) % o.Q_MOD, (1 << 512) % o.Q_MOD, o.R_MOD, o.Q_MOD):
        for k in range(8):
            assert "0x%08xu" % ((v >> (32 * k)) & 0xFFFFFFFF) in src


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_fixtures_cpp_oracle_reproduces_python_oracle(name):
    """Two independent CPU implementations (big-int Python, 4x64-limb C++) agree on H and on the proof bytes of every
    committed fixture; the fixture's proof satisfies the Groth16 equation in the exponent."""
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    val_m = [c.field_op(0, 5, v) if len(v) else v for v in mats.val]
    r1 = c.r1cs_struct(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, val_m)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    arr = {k: c.field_op(1, 5, v.reshape(-1, 4)).reshape(v.shape) for k, v in pk.arrays.items()}
    z_int = [int(v, 16) for v in meta["z"]]
    red = 1 if meta["reduction"] == "circom" else 0
    proof, h, _ = c.prove(c.pk_struct(arr), r1, g.fr_to_mont(z_int), g.fr_to_mont([int(meta["r"], 16)])[0],
                          g.fr_to_mont([int(meta["s"], 16)])[0], red, want_h=True)
    assert g.fr_from_mont(h) == [int(v, 16) for v in meta["h"]]
    P = g.Proof(g.g1_from_mont(proof[0]), g.g2_from_mont(proof[1]), g.g1_from_mont(proof[2]))
    assert P.serialize_uncompressed().hex() == meta["proof_uncompressed"]
    assert P.serialize_compressed().hex() == meta["proof_compressed"]
    assert len(P.serialize_compressed()) == 128 and len(P.serialize_uncompressed()) == 256
    A, B, C = (int(v, 16) for v in meta["proof_dlog"])
    assert P.a == o.G1.mul(o.G1_GEN, A) and P.c == o.G1.mul(o.G1_GEN, C)


def test_fixture_regeneration_is_deterministic():
    """The committed 'silly' fixture is exactly what tests/golden/make_golden.py produces today."""
    meta, _, pk_bytes = load_golden("silly")
    td = o.Trapdoor(**{k: int(v, 16) for k, v in meta["trapdoor"].items()})
    m, z = o.my_silly_circuit(o.stream_fr(7, 1), o.stream_fr(7, 2))
    pk, qap = o.generate_parameters(m, td)
    assert o.ser_pk(pk, False) == pk_bytes
    proof, h, _ = o.create_proof_with_reduction_and_matrices(pk, int(meta["r"], 16), int(meta["s"], 16), m, 2, 6, z)
    assert o.ser_proof(proof, False).hex() == meta["proof_uncompressed"]
    assert [hex(v) for v in h] == meta["h"]


def test_ntt_oracles_agree_and_invert():
    for lg in (0, 1, 4, 9):
        n = 1 << lg
        vals = [o.stream_fr(0x99, i) for i in range(n)]
        d = o.Domain(n)
        x = g.fr_to_mont(vals)
        assert g.fr_from_mont(c.ntt(x)) == d.fft(vals)
        assert g.fr_from_mont(c.ntt(x, inverse=True, coset=True)) == d.coset_ifft(vals)
        assert d.ifft(d.fft(vals)) == vals and d.coset_ifft(d.coset_fft(vals)) == vals
    # evaluation semantics: fft(coeffs)[i] == poly(omega^i)
    d = o.Domain(8)
    coeffs = [3, 1, 4, 1, 5, 9, 2, 6]
    ev = d.fft(coeffs)
    for i in range(8):
        w = d.element(i)
        assert ev[i] == sum(cf * pow(w, k, o.R_MOD) for k, cf in enumerate(coeffs)) % o.R_MOD


def test_msm_oracles_agree_including_edge_cases():
    rnd = random.Random(3)
    ks = [rnd.randrange(o.R_MOD) for _ in range(64)]
    pts = c.fixed_base(1, g.fr_to_mont(ks))
    pts[5] = 0   # infinity
    sc_int = [0, 1, o.R_MOD - 1, 2] + [rnd.randrange(o.R_MOD) for _ in range(60)]
    sc = g.fr_to_mont(sc_int)
    want = o.G1.to_affine(o.G1.msm([g.g1_from_mont(p) for p in pts], sc_int))
    assert g.g1_from_mont(c.msm(1, pts, sc)) == want
    assert g.g1_from_mont(c.msm(1, pts, sc, naive=True)) == want
    assert g.g1_from_mont(c.msm(1, pts[:0], sc[:0])) is None


def test_domain_too_large_is_rejected():
    with pytest.raises(ValueError):
        o.Domain((1 << 28) + 1)


def test_witness_map_quotient_identity():
    """h really is the quotient: (A*B - C)(x) == h(x) * Z(x) at a random point."""
    m, z = o.random_satisfiable_r1cs(5, 40, 3, 30)
    h = o.witness_map_libsnark(m, 3, 40, z)
    dom = o.Domain(43)
    assert h[dom.n - 1] == 0
    x = o.stream_fr(1, 1)
    lag = dom.lagrange_at(x)
    ev = lambda rows, extra: (sum(lag[i] * o.evaluate_constraint(rows[i], z) for i in range(40)) + extra) % o.R_MOD
    a_x = ev(m.a, sum(lag[40 + i] * z[i] for i in range(3)))
    b_x, c_x = ev(m.b, 0), ev(m.c, 0)
    h_x = sum(cf * pow(x, k, o.R_MOD) for k, cf in enumerate(h)) % o.R_MOD
    assert (a_x * b_x - c_x) % o.R_MOD == h_x * dom.vanishing(x) % o.R_MOD


# ---- round 2: pins added after VERDICT r01 ------------------------------------------------------------------------------------
# circuit_setup/circuits/circomlib/circuits/pointbits.circom:38-40 -- the one in-tree copy of the 2^28-th root of unity of
# BN254 Fr: Tonelli-Shanks constants m = 28, c = 5^((r-1)/2^28), exponent (r-1)/2^28.  The oracles' (and the kernels') NTT
# root rho must be this number: omega_n = rho^(2^(28 - log n)) (ASSUMPTION "arkworks TWO_ADIC_ROOT_OF_UNITY" pinned to bytes).
POINTBITS_M = 28
POINTBITS_C = 19103219067921713944291392827692070036145651957329286315305642004821462161904
POINTBITS_T_EXP = 81540058820840996586704275553141814055101440848469862132140264610111
POINTBITS_PATH = "/root/reference/circuit_setup/circuits/circomlib/circuits/pointbits.circom"


def test_root_of_unity_pinned_to_pointbits_circom():
    assert POINTBITS_T_EXP == (o.R_MOD - 1) >> POINTBITS_M and (o.R_MOD - 1) % (1 << POINTBITS_M) == 0
    assert o.FR_ROOT_2_28 == POINTBITS_C == pow(o.FR_GENERATOR, POINTBITS_T_EXP, o.R_MOD)
    assert pow(POINTBITS_C, 1 << 27, o.R_MOD) == o.R_MOD - 1          # primitive: order exactly 2^28
    for log_n in (1, 12, 21, 22, 28):
        assert o.Domain(1 << log_n).element(1) == pow(POINTBITS_C, 1 << (28 - log_n), o.R_MOD)
    # the C++ oracle's transform uses the same root: NTT of the delta at index 1 is [omega^k]
    n = 16
    e1 = g.fr_to_mont([0, 1] + [0] * (n - 2))
    w = pow(POINTBITS_C, 1 << (28 - 4), o.R_MOD)
    assert g.fr_from_mont(c.ntt(e1)) == [pow(w, k, o.R_MOD) for k in range(n)]
    if os.path.exists(POINTBITS_PATH):  # in the build container the literal is re-read from the reference tree itself
        src = open(POINTBITS_PATH).read().splitlines()
        assert f"var m = {POINTBITS_M};" in src[37] and f"var c = {POINTBITS_C};" in src[38]
        assert str(POINTBITS_T_EXP) in src[39]


def _td_of(meta):
    import ast
    td = ast.literal_eval(meta["trapdoor"]) if isinstance(meta["trapdoor"], str) else meta["trapdoor"]
    return o.Trapdoor(*(int(td[k], 16) for k in ("alpha", "beta", "gamma", "delta", "t")))


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if "circom" not in n])
def test_cpp_generator_reproduces_golden_key_bytes(name):
    """generate_parameters_with_qap on the host cores (oracle/refsynth.py over libg16oracle.so: what bench.py's reference arm
    mints its key with) == the committed arkworks-layout key bytes (generator.rs:50-228, the fork's gamma = 1, delta_g1 in vk)."""
    import types
    import refsynth
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    td = _td_of(meta)
    nc, ni = mats.num_constraints, mats.num_instance_variables
    m = ni + mats.num_witness_variables
    n = 1
    while n < nc + ni:
        n <<= 1
    inst = types.SimpleNamespace(matrices=mats, nc=nc, ni=ni, m=m, n=n)
    arrays, qap = refsynth.generate_parameters_cpu(inst, td)
    pk = g.ProvingKey(arrays, 0)
    pk.gamma_g2, pk.gamma_abc_g1 = qap["gamma_g2"], qap["gamma_abc_g1"]
    assert pk.serialize_uncompressed() == pk_bytes


def test_pk_serialisation_round_trip_all_fixtures():
    for name in GOLDEN_NAMES:
        _, _, pk_bytes = load_golden(name)
        assert g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes).serialize_uncompressed() == pk_bytes, name
    assert g._ser_points_uncompressed(np.zeros((2, 8), dtype=np.uint64), 1) == (bytes(63) + b"\x40") * 2   # infinity flag


def test_cpu_instance_builder_is_satisfied_and_deterministic():
    """oracle/refsynth.make_instance_cpu (the reference arm's builder): every row satisfied, h[n-1] == 0, same arrays twice."""
    import refsynth
    a = refsynth.make_instance_cpu("S-2^12", seed=0x5A0CE)
    b = refsynth.make_instance_cpu("S-2^12", seed=0x5A0CE)
    assert np.array_equal(a.z_mont, b.z_mont) and all(np.array_equal(x, y) for x, y in zip(a.matrices.val, b.matrices.val))
    r1 = refsynth.r1cs_of(a)
    az, bz, cz = c.r1cs_eval(r1, a.z_mont)
    assert np.array_equal(c.field_op(0, 0, az, bz), cz)
    h = c.witness_map(r1, a.z_mont, a.n)
    assert not h[-1].any() and h.any()
    # broadcast / power-table helpers of the oracle against big integers
    x = g.fr_to_mont([3, 5, o.R_MOD - 1])
    assert g.fr_from_mont(c.field_op(0, 8, x, g.fr_to_mont([7]))) == [21, 35, (o.R_MOD - 7) % o.R_MOD]
    assert g.fr_from_mont(c.field_op(0, 9, x, g.fr_to_mont([7]))) == [10, 12, 6]
    assert g.fr_from_mont(c.pow_table(g.fr_to_mont([3])[0], g.fr_to_mont([2])[0], 70)) == [2 * pow(3, i, o.R_MOD) % o.R_MOD for i in range(70)]


def test_bench_exponent_check_is_not_vacuous():
    """bench.check_proof_in_exponent (the gate in front of every bench number) on a CPU-built instance and key: it accepts the
    oracle's proof and rejects (i) a proof whose C carries a wrong h -- one coefficient off, every MSM right: the case the
    round-1 check let through because it took h(t)Z(t) from the prover's own h --, (ii) a proof for another witness,
    (iii) a proof made with another r."""
    import sys
    import types
    import refsynth
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    inst = refsynth.make_instance_cpu("S-2^12", seed=0x5A0CE)
    td = types.SimpleNamespace(**bench.TRAPDOOR)
    arrays, qap = refsynth.generate_parameters_cpu(inst, td)
    r_int, s_int = bench.R_INT % o.R_MOD, bench.S_INT % o.R_MOD
    r_m, s_m = g.fr_to_mont([r_int])[0], g.fr_to_mont([s_int])[0]
    r1, pk = refsynth.r1cs_of(inst), c.pk_struct(arrays)

    def proof_of(z_mont, r_mont):
        pr, _, _ = c.prove(pk, r1, z_mont, r_mont, s_m, threads=4)
        return g.Proof(g.g1_from_mont(pr[0]), g.g2_from_mont(pr[1]), g.g1_from_mont(pr[2]))

    good = proof_of(inst.z_mont, r_m)
    assert bench.check_proof_in_exponent(good, inst, qap, td, r_int, s_int)
    # (i) h[5] off by one: C moves by h_query[5], A and B stay right
    wrong_h = g.Proof(good.a, good.b, o.G1.add(good.c, g.g1_from_mont(arrays["h_query"][5])))
    assert not bench.check_proof_in_exponent(wrong_h, inst, qap, td, r_int, s_int)
    # (ii) another (still satisfying) witness is another proof: rebuild the instance from another seed, prove it under its
    # own key, check it against THIS instance
    other = refsynth.make_instance_cpu("S-2^12", seed=0x5A0CF)
    arrays2, _ = refsynth.generate_parameters_cpu(other, td)
    pr2, _, _ = c.prove(c.pk_struct(arrays2), refsynth.r1cs_of(other), other.z_mont, r_m, s_m, threads=4)
    foreign = g.Proof(g.g1_from_mont(pr2[0]), g.g2_from_mont(pr2[1]), g.g1_from_mont(pr2[2]))
    assert not bench.check_proof_in_exponent(foreign, inst, qap, td, r_int, s_int)
    # (iii) another r
    assert not bench.check_proof_in_exponent(proof_of(inst.z_mont, g.fr_to_mont([r_int + 1])[0]), inst, qap, td, r_int, s_int)
