"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, refuses to compute
without a device (no CPU fallback), and the host-compiled Montgomery code equals the oracle."""
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

import pyref as o
from crescent_credentials_b200 import ffi
from crescent_credentials_b200 import groth16 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "g16_b200.h")).read()
    declared = set(re.findall(r"\b(g16_[a-z0-9_]+)\s*\(", hdr))
    lib = ffi.load_library()
    assert declared, "no declarations found"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"libg16b200.so does not export {sym}"
    assert declared == set(ffi.EXPORTS), declared ^ set(ffi.EXPORTS)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(ffi.ProofOut) == 8 * 8 + 16 * 8 + 8 * 8 + 16
    assert ctypes.sizeof(ffi.Partial) == ffi.PARTIAL_U64 * 8 == 896
    assert ctypes.sizeof(ffi.Timings) == 17 * 4


def test_no_cpu_fallback_without_device():
    lib = ffi.load_library()
    if lib.g16_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(ffi.G16Error) as e:
        ffi.Context(0)
    assert e.value.code == ffi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "crescent_credentials_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                # comments may cite the oracle; code may not import, include, link or dlopen it
                assert not re.search(r"^\s*(import|from)\s+(pyref|coracle|oracle|pairing)\b", src, re.M), fn
                assert not re.search(r"#include\s*[<\"][^>\"]*oracle", src), fn
                assert "libg16oracle" not in src and "CDLL(" not in src.replace("C.CDLL(LIB_PATH)", ""), fn


@pytest.fixture(scope="module")
def host_fp():
    out = os.path.join(ROOT, "tests", "_host_fp.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "host_fp_shim.cpp")])
    return ctypes.CDLL(out)


def _call(fn, op, a, b, n=32):
    A, B, R = ctypes.create_string_buffer(a, n), ctypes.create_string_buffer(b, n), ctypes.create_string_buffer(n)
    fn(op, A, B, R)
    return R.raw


@pytest.mark.parametrize("name,p", [("fr", o.R_MOD), ("fq", o.Q_MOD)])
def test_device_montgomery_code_on_host_matches_oracle(host_fp, name, p):
    """csrc/fp.cuh compiled for the host (carry chains emulated in C): same control flow as the CUDA kernels."""
    fn = getattr(host_fp, f"host_{name}_op")
    rnd = random.Random(1)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (p - 1) // 2, (p + 1) // 2, (1 << 253) % p]
    vals = edge + [rnd.randrange(p) for _ in range(1500)]
    rinv = pow(1 << 256, -1, p)
    le = lambda v: v.to_bytes(32, "little")
    dec = lambda b: int.from_bytes(b, "little")
    for k, a in enumerate(vals):
        others = edge if k < len(edge) else [vals[(k * 7 + 3) % len(vals)]]
        for b in others:
            assert dec(_call(fn, 0, le(a), le(b))) == a * b * rinv % p
            assert dec(_call(fn, 8, le(a), le(b))) == a * b * rinv % p   # Karatsuba + separate reduction variant
            assert dec(_call(fn, 1, le(a), le(b))) == (a + b) % p
            assert dec(_call(fn, 2, le(a), le(b))) == (a - b) % p
        assert dec(_call(fn, 3, le(a), le(0))) == (-a) % p
        assert dec(_call(fn, 5, le(a), le(0))) == (a << 256) % p
        assert dec(_call(fn, 6, le(a), le(0))) == a * rinv % p
    for a in vals[1:30]:
        assert dec(_call(fn, 4, le((a << 256) % p), le(0))) == (pow(a, -1, p) << 256) % p
    # binary-GCD inverse (the one the batched-affine MSM uses): same answers, and 0 -> 0
    for a in vals[1:400]:
        assert dec(_call(fn, 9, le((a << 256) % p), le(0))) == (pow(a, -1, p) << 256) % p
    assert dec(_call(fn, 9, le(0), le(0))) == 0
    # divsteps ("safegcd") inverse -- what k_ba_invert and the normalisations call: edge values, small values, 0 -> 0
    small = [rnd.randrange(1, 1 << k) for k in range(1, 254, 3)]
    for a in vals[1:1200] + small:
        assert dec(_call(fn, 10, le((a << 256) % p), le(0))) == (pow(a, -1, p) << 256) % p
    assert dec(_call(fn, 10, le(0), le(0))) == 0


def test_windowed_scalar_mul_on_host_matches_oracle(host_fp):
    """ec.cuh scalar_mul_window (k_scale_point: s * MSM_a, r * MSM_b1) against the oracle and the plain double-and-add: edge
    scalars (0, 1, 8, 9, all-ones nibbles, r - 1, 2^255-ish carries into the 65th digit) and random ones."""
    rnd = random.Random(9)
    q = o.Q_MOD
    enc = lambda P: b"".join(((v << 256) % q).to_bytes(32, "little") for v in P)
    rinv = pow(1 << 256, -1, q)
    def dec(b):
        x, y = (int.from_bytes(b[i:i + 32], "little") * rinv % q for i in (0, 32))
        return None if x == 0 and y == 0 else (x, y)
    P = o.G1.mul(o.G1_GEN, 0xC0FFEE1234567)
    ks = [0, 1, 2, 7, 8, 9, 15, 16, 0x88888888, (1 << 256) - 1, int("9" * 64, 16), int("8" * 64, 16), o.R_MOD - 1, o.R_MOD, 1 << 255]
    ks += [rnd.randrange(1 << 256) for _ in range(12)] + [rnd.randrange(o.R_MOD) for _ in range(12)]
    for k in ks:
        outs = []
        for which in (0, 1):
            A = ctypes.create_string_buffer(enc(P), 64)
            K = ctypes.create_string_buffer(k.to_bytes(32, "little"), 32)
            R = ctypes.create_string_buffer(64)
            host_fp.host_g1_scalar_mul(A, K, which, R)
            outs.append(dec(R.raw))
        want = o.G1.to_affine(o.G1.jmul(o.G1.to_jac(P), k))
        assert outs[0] == want and outs[1] == want, hex(k)


def test_glv_split_on_host_matches_oracle(host_fp):
    """glv.cuh: the constants (beta, lambda, the lattice basis behind the rounding constants), the decomposition
    k = k1 + k2 * lambda (mod r) with short halves, and k * P = k1 * P + k2 * phi(P) against the oracle (what k_scale_point runs)."""
    r, q = o.R_MOD, o.Q_MOD
    beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    assert pow(beta, 3, q) == 1 and beta != 1 and pow(lam, 3, r) == 1 and lam != 1
    P = o.G1.mul(o.G1_GEN, 0xC0FFEE1234567)
    assert o.G1.mul(P, lam) == (beta * P[0] % q, P[1])                      # phi(P) = lambda * P
    rnd = random.Random(10)
    ks = [0, 1, 2, r - 1, r - 2, lam, r - lam, r // 2, 1 << 253, (1 << 128) - 1, 1 << 128] + [rnd.randrange(r) for _ in range(3000)]
    for k in ks:
        K = ctypes.create_string_buffer(k.to_bytes(32, "little"), 32)
        out = (ctypes.c_uint32 * 13)()
        host_fp.host_glv_decompose(K, out)
        k1 = sum(out[i] << (32 * i) for i in range(5)) * (-1 if out[5] else 1)
        k2 = sum(out[6 + i] << (32 * i) for i in range(5)) * (-1 if out[11] else 1)
        assert out[12] == 1 and abs(k1) < 1 << 128 and abs(k2) < 1 << 128, hex(k)
        assert (k1 + k2 * lam) % r == k, hex(k)
    enc = lambda Pt: b"".join(((v << 256) % q).to_bytes(32, "little") for v in Pt)
    rinv = pow(1 << 256, -1, q)
    for k in ks[:11] + ks[-20:]:
        A = ctypes.create_string_buffer(enc(P), 64)
        K = ctypes.create_string_buffer(k.to_bytes(32, "little"), 32)
        R = ctypes.create_string_buffer(64)
        host_fp.host_g1_scalar_mul_glv(A, K, R)
        x, y = (int.from_bytes(R.raw[i:i + 32], "little") * rinv % q for i in (0, 32))
        got = None if x == 0 and y == 0 else (x, y)
        assert got == o.G1.to_affine(o.G1.jmul(o.G1.to_jac(P), k)), hex(k)


def test_device_fq2_code_on_host_matches_oracle(host_fp):
    rnd = random.Random(2)
    enc = lambda x: b"".join(((v << 256) % o.Q_MOD).to_bytes(32, "little") for v in x)
    rinv = pow(1 << 256, -1, o.Q_MOD)
    dec = lambda bs: tuple(int.from_bytes(bs[i:i + 32], "little") * rinv % o.Q_MOD for i in (0, 32))
    for _ in range(300):
        a = (rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD))
        b = (rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD))
        assert dec(_call(host_fp.host_fq2_op, 0, enc(a), enc(b), 64)) == o.Fq2.mul(a, b)
        assert dec(_call(host_fp.host_fq2_op, 7, enc(a), enc(b), 64)) == o.Fq2.sqr(a)
        assert dec(_call(host_fp.host_fq2_op, 4, enc(a), enc(b), 64)) == o.Fq2.inv(a)


def test_lockstep_triple_product_on_host_matches_oracle():
    """fp.cuh mul_cios3 (three Montgomery products row-interleaved; Fq2 products use it under -DG16_FQ2_MUL3, off by default):
    the host build with the switch on must give the oracle's Fq2 products, and the whole pairing built on it the oracle's GT."""
    import random
    import pairing as P
    out = os.path.join(ROOT, "tests", "_host_fp_mul3.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DG16_FQ2_MUL3", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "host_fp_shim.cpp")])
    lib = ctypes.CDLL(out)
    rnd = random.Random(33)
    enc = lambda v: o.mont_le_bytes(v[0], o.Q_MOD) + o.mont_le_bytes(v[1], o.Q_MOD)
    dec = lambda b: (o.from_mont(int.from_bytes(b[:32], "little"), o.Q_MOD), o.from_mont(int.from_bytes(b[32:], "little"), o.Q_MOD))
    edge = [0, 1, o.Q_MOD - 1, (o.Q_MOD - 1) // 2]
    cases = [((x, y), (y, x)) for x in edge for y in edge] + [((rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD)),
                                                               (rnd.randrange(o.Q_MOD), rnd.randrange(o.Q_MOD))) for _ in range(40)]
    for a, b in cases:
        assert dec(_call(lib.host_fq2_op, 0, enc(a), enc(b), 64)) == o.Fq2.mul(a, b)
    outp = os.path.join(ROOT, "tests", "_host_pairing_mul3.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DG16_FQ2_MUL3", "-shared", "-fPIC", "-o", outp, os.path.join(ROOT, "tests", "host_pairing_shim.cpp")])
    hp = ctypes.CDLL(outp)
    from test_pairing_host import enc_g1, enc_g2, dec_f12
    p, q = o.G1.mul(o.G1_GEN, o.stream_fr(0x3A, 1)), o.G2.mul(o.G2_GEN, o.stream_fr(0x3A, 2))
    act = (ctypes.c_int * 3)(1, 0, 0)
    zero = ctypes.create_string_buffer(91 * 192)
    f, gt = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
    pts = enc_g1(p) + bytes(128)
    hp.host_miller3(ctypes.create_string_buffer(pts, len(pts)), act, ctypes.create_string_buffer(enc_g2(q), 128), zero, zero, f)
    assert hp.host_final_exp(f, gt) == 1
    assert dec_f12(gt.raw) == P.pairing(p, q)


def test_serialisation_mirror_matches_oracle():
    rnd = random.Random(9)
    for _ in range(20):
        P = o.G1.mul(o.G1_GEN, rnd.randrange(1, o.R_MOD))
        Q = o.G2.mul(o.G2_GEN, rnd.randrange(1, o.R_MOD))
        for comp in (True, False):
            assert g._ser_g1(P, comp) == o.ser_g1(P, comp) and g._ser_g2(Q, comp) == o.ser_g2(Q, comp)
        assert o.deser_g1_uncompressed(o.ser_g1(P, False)) == P and o.deser_g2_uncompressed(o.ser_g2(Q, False)) == Q
    assert g._ser_g1(None, True) == o.ser_g1(None, True) and g._ser_g2(None, False) == o.ser_g2(None, False)
    pr = g.Proof(None, None, None)
    assert pr.serialize_compressed()[31] == 0x40 and len(pr.serialize_uncompressed()) == 256


def test_fr_sampling_shape():
    """Fr::rand: rejection sampling of 254-bit values, accepted integer is the Montgomery representation."""
    class Fixed:
        def __init__(self, words):
            self.w = list(words)

        def getrandbits(self, k):
            return self.w.pop(0)
    top_too_big = [0xFFFFFFFFFFFFFFFF] * 4          # masks to 2^254 - 1 >= r: rejected
    ok = [5, 0, 0, 0]
    assert g.sample_fr(Fixed(top_too_big + ok)) == 5 * pow(1 << 256, -1, o.R_MOD) % o.R_MOD
