"""ctypes wrapper of oracle/libg16oracle.so (g16_oracle.cpp).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libg16oracle.so")
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


class OcPk(C.Structure):
    _fields_ = [("a_query", C.c_void_p), ("a_len", C.c_size_t), ("b_g1_query", C.c_void_p), ("b_g1_len", C.c_size_t),
                ("b_g2_query", C.c_void_p), ("b_g2_len", C.c_size_t), ("h_query", C.c_void_p), ("h_len", C.c_size_t),
                ("l_query", C.c_void_p), ("l_len", C.c_size_t), ("alpha_g1", C.c_void_p), ("beta_g1", C.c_void_p),
                ("delta_g1", C.c_void_p), ("beta_g2", C.c_void_p), ("delta_g2", C.c_void_p)]


class OcR1cs(C.Structure):
    _fields_ = [("nc", C.c_uint64), ("ni", C.c_uint64), ("m", C.c_uint64), ("row_ptr", C.c_void_p * 3),
                ("col", C.c_void_p * 3), ("val", C.c_void_p * 3)]


class OcProof(C.Structure):
    _fields_ = [("a", C.c_uint64 * 8), ("b", C.c_uint64 * 16), ("c", C.c_uint64 * 8)]


class OcTimings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("witness_map_s", "msm_h_s", "msm_l_s", "msm_a_s", "msm_b_g1_s", "msm_b_g2_s", "total_s")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oc_hardware_threads.restype = C.c_int
    return _lib


def hardware_threads() -> int:
    return lib().oc_hardware_threads()


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def field_op(field, op, a, b=None, threads=0):
    """op codes of include/g16_b200.h (G16_OP_*): 0 mul, 1 add, 2 sub, 3 neg, 4 inv, 5 to_mont, 6 from_mont, 7 sqr,
    8 mul by ONE broadcast element b, 9 add ONE broadcast element b."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        assert b.shape[0] == (1 if op in (8, 9) else a.shape[0])
    lib().oc_field_op_mt(field, op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]), threads or hardware_threads())
    return out


def pow_table(base, scale, n, threads=0):
    """out[i] = scale * base^i (Fr, Montgomery)."""
    base = np.ascontiguousarray(base, dtype=np.uint64)
    scale = np.ascontiguousarray(scale, dtype=np.uint64)
    out = np.empty((n, 4), dtype=np.uint64)
    lib().oc_pow_table(_p(base), _p(scale), C.c_size_t(n), _p(out), threads or hardware_threads())
    return out


def ntt(data, inverse=False, coset=False, threads=0):
    d = np.array(data, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    n = d.shape[0]
    rc = lib().oc_ntt(_p(d), n.bit_length() - 1, int(inverse), int(coset), threads or hardware_threads())
    assert rc == 0
    return d


def r1cs_struct(nc, ni, m, row_ptr, col, val):
    """val must be Montgomery limbs."""
    r = OcR1cs()
    r.nc, r.ni, r.m = nc, ni, m
    keep = []
    for k in range(3):
        rp = np.ascontiguousarray(row_ptr[k], dtype=np.uint64)
        cc = np.ascontiguousarray(col[k], dtype=np.uint32)
        vv = np.ascontiguousarray(val[k], dtype=np.uint64)
        keep += [rp, cc, vv]
        r.row_ptr[k], r.col[k], r.val[k] = rp.ctypes.data, cc.ctypes.data, vv.ctypes.data
    r._keep = keep
    return r


def witness_map(r, z, n, reduction=0, threads=0):
    z = np.ascontiguousarray(z, dtype=np.uint64)
    h = np.zeros((n, 4), dtype=np.uint64)
    rc = lib().oc_witness_map(C.byref(r), _p(z), reduction, _p(h), threads or hardware_threads())
    if rc:
        raise RuntimeError(f"oc_witness_map rc={rc}")
    return h


def r1cs_eval(r, z, threads=0):
    z = np.ascontiguousarray(z, dtype=np.uint64)
    out = [np.zeros((r.nc, 4), dtype=np.uint64) for _ in range(3)]
    lib().oc_r1cs_eval(C.byref(r), _p(z), _p(out[0]), _p(out[1]), _p(out[2]), threads or hardware_threads())
    return out


def msm(group, points, scalars, naive=False, threads=0):
    pw = 8 if group == 1 else 16
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, pw)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(points.shape[0], scalars.shape[0])
    out = np.zeros(pw, dtype=np.uint64)
    fn = lib().oc_msm_g1 if group == 1 else lib().oc_msm_g2
    fn(_p(points), _p(scalars), C.c_size_t(n), _p(out), int(naive), threads or hardware_threads())
    return out


def fixed_base(group, scalars, threads=0):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((scalars.shape[0], 8 if group == 1 else 16), dtype=np.uint64)
    lib().oc_fixed_base(group, _p(scalars), C.c_size_t(scalars.shape[0]), _p(out), threads or hardware_threads())
    return out


def instance_map(r, t_mont, threads=0):
    t_mont = np.ascontiguousarray(t_mont, dtype=np.uint64)
    a, b, c = (np.zeros((r.m, 4), dtype=np.uint64) for _ in range(3))
    zt = np.zeros(4, dtype=np.uint64)
    n = C.c_uint64(0)
    rc = lib().oc_instance_map(C.byref(r), _p(t_mont), _p(a), _p(b), _p(c), _p(zt), C.byref(n), threads or hardware_threads())
    if rc:
        raise RuntimeError(f"oc_instance_map rc={rc}")
    return a, b, c, zt, int(n.value)


def pk_struct(arrays):
    """arrays as in crescent_credentials_b200.ffi.Context.load_pk, Montgomery limbs."""
    pk = OcPk()
    keep = {}
    for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
        w = 16 if name == "b_g2_query" else 8
        a = np.ascontiguousarray(arrays[name], dtype=np.uint64).reshape(-1, w)
        keep[name] = a
        setattr(pk, name, a.ctypes.data)
        setattr(pk, name.replace("_query", "_len"), a.shape[0])
    for name in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
        a = np.ascontiguousarray(arrays[name], dtype=np.uint64).reshape(-1)
        keep[name] = a
        setattr(pk, name, a.ctypes.data)
    pk._keep = keep
    return pk


def prove(pk, r, z, r_mont, s_mont, reduction=0, want_h=False, threads=0):
    z = np.ascontiguousarray(z, dtype=np.uint64)
    r_mont = np.ascontiguousarray(r_mont, dtype=np.uint64)
    s_mont = np.ascontiguousarray(s_mont, dtype=np.uint64)
    out = OcProof()
    tm = OcTimings()
    n = 1
    while n < r.nc + r.ni:
        n <<= 1
    h = np.zeros((n, 4), dtype=np.uint64) if want_h else None
    rc = lib().oc_prove(C.byref(pk), C.byref(r), _p(z), _p(r_mont), _p(s_mont), reduction, C.byref(out), _p(h), C.byref(tm),
                        threads or hardware_threads())
    if rc:
        raise RuntimeError(f"oc_prove rc={rc}")
    proof = (np.array(out.a[:], dtype=np.uint64), np.array(out.b[:], dtype=np.uint64), np.array(out.c[:], dtype=np.uint64))
    return proof, h, tm.as_dict()


# ---- the verifier (row f-4): pairing / prepare_verifying_key / prepare_inputs / verify_proof -----------------------------------------
class OcVk(C.Structure):
    _fields_ = [("alpha_g1", C.c_void_p), ("beta_g2", C.c_void_p), ("gamma_g2", C.c_void_p), ("delta_g2", C.c_void_p),
                ("gamma_abc_g1", C.c_void_p), ("gamma_abc_len", C.c_size_t)]


def vk_struct(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1):
    """Montgomery limb arrays (8 / 16 / 16 / 16 words, (len, 8)); the returned struct keeps them alive."""
    keep = [np.ascontiguousarray(a, dtype=np.uint64) for a in (alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1)]
    v = OcVk(*[a.ctypes.data for a in keep], keep[4].reshape(-1, 8).shape[0])
    v._keep = keep
    return v


def pairing(g1_points, g2_points, threads=0):
    p = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, 8)
    q = np.ascontiguousarray(g2_points, dtype=np.uint64).reshape(-1, 16)
    out = np.zeros((p.shape[0], 48), dtype=np.uint64)
    lib().oc_pairing(_p(p), _p(q), C.c_size_t(p.shape[0]), _p(out), C.c_int(threads))
    return out


def prepare_vk(vk: OcVk):
    out = np.zeros(48, dtype=np.uint64)
    lib().oc_prepare_vk(C.byref(vk), _p(out))
    return out


def prepare_inputs(vk: OcVk, inputs, n, threads=0):
    x = np.ascontiguousarray(inputs, dtype=np.uint64)
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oc_prepare_inputs(C.byref(vk), _p(x) if x.size else None, C.c_size_t(n), _p(out), C.c_int(threads))
    return out


def verify(vk: OcVk, proofs, inputs, n, threads=0):
    """proofs: (n, 32) uint64 = a | b | c Montgomery affine.  Returns (verdicts uint8[n], seconds of the per-proof part)."""
    pr = np.ascontiguousarray(proofs, dtype=np.uint64).reshape(n, 32)
    x = np.ascontiguousarray(inputs, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint8)
    f = lib().oc_verify
    f.restype = C.c_double
    secs = f(C.byref(vk), _p(pr), _p(x) if x.size else None, C.c_size_t(n), _p(out), C.c_int(threads))
    return out, secs
