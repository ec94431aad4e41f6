"""ctypes binding of libg16b200.so (include/g16_b200.h).  No compute happens in Python and there is no fallback:
if the CUDA library is missing or no device is visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G16_LIB", os.path.join(_HERE, "libg16b200.so"))  # G16_LIB: kernel-variant experiments

G16_OK = 0
ERR_DEGREE_TOO_LARGE, ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE, ERR_VANISHING_ZERO = 1, 2, 3, 4, 5, 6
REDUCTION_LIBSNARK, REDUCTION_CIRCOM = 0, 1
WM_PART_A, WM_PART_B, WM_PART_C, WM_PART_FINAL = 1, 2, 4, 8
ENC_MONTGOMERY, ENC_CANONICAL = 0, 1
FIELD_FR, FIELD_FQ, FIELD_FQ2 = 0, 1, 2
OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_INV, OP_TO_MONT, OP_FROM_MONT, OP_SQR, OP_MUL_BCAST, OP_ADD_BCAST = range(10)
PARTIAL_U64 = 5 * 16 + 32

# every symbol include/g16_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "g16_device_count", "g16_ctx_create", "g16_ctx_destroy", "g16_last_error", "g16_version", "g16_ctx_load_pk",
    "g16_ctx_load_r1cs", "g16_prove", "g16_upload_witness", "g16_prove_resident", "g16_prove_shard",
    "g16_prove_combine", "g16_partial_dev", "g16_prove_shard_dev", "g16_prove_combine_dev", "g16_witness_map",
    "g16_domain_size", "g16_get_timings", "g16_msm_g1", "g16_msm_g2", "g16_msm_set_bases", "g16_msm_set_bases_dev",
    "g16_msm_run_dev", "g16_msm_window_bits", "g16_ntt", "g16_ntt_dev", "g16_field_op", "g16_fixed_base_g1", "g16_fixed_base_g2",
    "g16_fixed_base_g1_dev", "g16_fixed_base_g2_dev", "g16_r1cs_eval", "g16_dev_alloc", "g16_dev_free",
    "g16_dev_upload", "g16_dev_download", "g16_sync", "g16_bench_int_pipe", "g16_launch_count", "g16_set_option",
    "g16_pow_table", "g16_copy_partial_dev", "g16_prove_prepare", "g16_get_msm_stats", "g16_ctx_load_pk_ranges",
    "g16_prove_shard_begin_dev", "g16_prove_shard_finish_dev", "g16_copy_h_dev",
    "g16_witness_map_part_dev", "g16_wm_vector_copy_dev", "g16_upload_witness_async", "g16_upload_witness_dev", "g16_memcpy_h2d_async", "g16_msm_copy_result_dev", "g16_msm_combine_dev", "g16_graph_stats",
    "g16_ctx_load_vk", "g16_vk_alpha_beta", "g16_prepare_inputs", "g16_verify_batch", "g16_verify_batch_prepared",
    "g16_verify_batch_dev", "g16_pairing", "g16_host_alloc", "g16_host_free",
]
VERDICT_REJECT, VERDICT_ACCEPT, VERDICT_UNEXPECTED_IDENTITY = 0, 1, 2

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


class PkView(C.Structure):
    _fields_ = [
        ("a_query", _u64p), ("a_len", C.c_size_t),
        ("b_g1_query", _u64p), ("b_g1_len", C.c_size_t),
        ("b_g2_query", _u64p), ("b_g2_len", C.c_size_t),
        ("h_query", _u64p), ("h_len", C.c_size_t),
        ("l_query", _u64p), ("l_len", C.c_size_t),
        ("alpha_g1", _u64p), ("beta_g1", _u64p), ("delta_g1", _u64p), ("beta_g2", _u64p), ("delta_g2", _u64p),
        ("encoding", C.c_int),
    ]


class R1csView(C.Structure):
    _fields_ = [
        ("num_constraints", C.c_uint64), ("num_instance", C.c_uint64), ("num_wires", C.c_uint64),
        ("row_ptr", _u64p * 3), ("col", _u32p * 3), ("val", _u64p * 3), ("encoding", C.c_int),
    ]


class VkView(C.Structure):
    _fields_ = [("alpha_g1", _u64p), ("beta_g2", _u64p), ("gamma_g2", _u64p), ("delta_g2", _u64p),
                ("gamma_abc_g1", _u64p), ("gamma_abc_len", C.c_size_t), ("encoding", C.c_int)]


class ProofOut(C.Structure):
    _fields_ = [("a", C.c_uint64 * 8), ("b", C.c_uint64 * 16), ("c", C.c_uint64 * 8),
                ("a_inf", C.c_int32), ("b_inf", C.c_int32), ("c_inf", C.c_int32), ("_pad", C.c_int32)]


class Partial(C.Structure):
    _fields_ = [("w", C.c_uint64 * PARTIAL_U64)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "witness_map_ms", "msm_h_ms", "msm_l_ms", "msm_a_ms",
                                         "msm_b_g1_ms", "msm_b_g2_ms", "assemble_ms", "total_ms")] + [
        ("acc_ms", C.c_float * 5), ("assemble_kernel_ms", C.c_float), ("h_wait_ms", C.c_float), ("h_start_ms", C.c_float)]

    def as_dict(self):
        d = {n: float(getattr(self, n)) for n, _ in self._fields_[:9]}
        d["acc_ms"] = dict(zip(("h", "l", "a", "b_g1", "b_g2"), (float(x) for x in self.acc_ms)))
        d["assemble_kernel_ms"] = float(self.assemble_kernel_ms)
        d["h_wait_ms"] = float(self.h_wait_ms)
        d["h_start_ms"] = float(self.h_start_ms)
        return d


class G16Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libg16b200 error {code}: {msg}")
        self.code = code


class PolynomialDegreeTooLarge(G16Error):
    """SynthesisError::PolynomialDegreeTooLarge (forks/groth16/src/r1cs_to_qap.rs:156-157)."""


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.g16_last_error.restype = C.c_char_p
    lib.g16_last_error.argtypes = [C.c_void_p]
    lib.g16_version.restype = C.c_char_p
    lib.g16_launch_count.restype = C.c_uint64
    lib.g16_launch_count.argtypes = [C.c_void_p]
    lib.g16_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.g16_ctx_destroy.argtypes = [C.c_void_p]
    lib.g16_ctx_destroy.restype = None
    lib.g16_ctx_load_pk.argtypes = [C.c_void_p, C.POINTER(PkView), C.c_int, C.c_int, C.c_int]
    lib.g16_ctx_load_pk_ranges.argtypes = [C.c_void_p, C.POINTER(PkView), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.g16_prove_shard_begin_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.g16_prove_shard_finish_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.g16_copy_h_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.g16_ctx_load_r1cs.argtypes = [C.c_void_p, C.POINTER(R1csView)]
    lib.g16_prove.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(ProofOut)]
    lib.g16_upload_witness.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_upload_witness_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.g16_upload_witness_dev.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_memcpy_h2d_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.g16_witness_map_part_dev.argtypes = [C.c_void_p, C.c_int]
    lib.g16_wm_vector_copy_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    lib.g16_prove_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(ProofOut)]
    lib.g16_prove_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Partial)]
    lib.g16_prove_combine.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(ProofOut)]
    lib.g16_partial_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.g16_prove_shard_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.g16_copy_partial_dev.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_prove_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.g16_get_msm_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.g16_prove_combine_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(ProofOut)]
    lib.g16_witness_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.g16_domain_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    lib.g16_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
    lib.g16_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
    lib.g16_msm_g2.argtypes = lib.g16_msm_g1.argtypes
    lib.g16_msm_set_bases.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    lib.g16_msm_set_bases_dev.argtypes = lib.g16_msm_set_bases.argtypes
    lib.g16_msm_run_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
    lib.g16_msm_window_bits.argtypes = [C.c_size_t, C.c_int]
    lib.g16_msm_copy_result_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.g16_msm_combine_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.g16_graph_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_ntt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int]
    lib.g16_ntt_dev.argtypes = lib.g16_ntt.argtypes
    lib.g16_field_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    for n in ("g16_fixed_base_g1", "g16_fixed_base_g2", "g16_fixed_base_g1_dev", "g16_fixed_base_g2_dev"):
        getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.g16_r1cs_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.g16_dev_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.g16_dev_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.g16_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.g16_dev_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.g16_sync.argtypes = [C.c_void_p]
    lib.g16_bench_int_pipe.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    lib.g16_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.g16_pow_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.g16_ctx_load_vk.argtypes = [C.c_void_p, C.POINTER(VkView)]
    lib.g16_vk_alpha_beta.argtypes = [C.c_void_p, C.c_void_p]
    lib.g16_prepare_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.g16_verify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.g16_verify_batch_prepared.argtypes = lib.g16_verify_batch.argtypes
    lib.g16_verify_batch_dev.argtypes = lib.g16_verify_batch.argtypes
    lib.g16_pairing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    _lib = lib
    return lib


def memcpy_h2d_async(dst_dev: int, src_host: int, nbytes: int, stream: int):
    """cudaMemcpyAsync(host -> device) on a raw stream handle, through the runtime the library is linked with (used by the
    sharded host glue for page-locked witness chunks that are addressed by pointer, not by a torch tensor)."""
    lib = load_library()
    rc = lib.g16_memcpy_h2d_async(C.c_void_p(dst_dev), C.c_void_p(src_host), C.c_size_t(nbytes), C.c_void_p(stream))
    if rc != G16_OK:
        raise G16Error(rc, "cudaMemcpyAsync(host -> device) failed")


def _ptr(a):
    """numpy array / int address / None -> void* (arrays must be C-contiguous)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)


def msm_window_bits(n: int, precompute: bool = True) -> int:
    """The Pippenger window the library picks for n bases (g16_msm_window_bits; no device needed)."""
    return int(load_library().g16_msm_window_bits(n, int(precompute)))


class Context:
    """Owns one g16_ctx (one CUDA device).  `stream` is an optional raw cudaStream_t (e.g. torch's current stream)."""

    def __init__(self, device: int = 0, stream: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.g16_ctx_create(C.byref(h), device, C.c_void_p(stream) if stream else None)
        if rc != G16_OK:
            raise G16Error(rc, self.lib.g16_last_error(None).decode())
        self.h = h
        self.device = device
        self.vk_inputs = None
        self.vk_generation = 0   # bumped by every load_vk: lets several Verifiers share one context safely
        self._wires = None       # num_wires of the loaded R1CS (size checks of z)

    def close(self):
        if getattr(self, "h", None):
            self.lib.g16_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != G16_OK:
            msg = self.lib.g16_last_error(self.h).decode()
            if rc == ERR_DEGREE_TOO_LARGE:
                raise PolynomialDegreeTooLarge(rc, msg)
            raise G16Error(rc, msg)

    # -- building blocks -------------------------------------------------------------------------------------
    def field_op(self, field: int, op: int, a: np.ndarray, b: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        w = 8 if field == FIELD_FQ2 else 4
        n = a.size // w
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
        self.check(self.lib.g16_field_op(self.h, field, op, _ptr(a), _ptr(b), _ptr(out), n))
        return out

    def ntt(self, data: np.ndarray, inverse: bool = False, coset: bool = False) -> np.ndarray:
        d = np.array(data, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
        n = d.shape[0]
        log_n = max(n.bit_length() - 1, 0)
        if n == 0 or (1 << log_n) != n:
            raise G16Error(ERR_BAD_ARG, "NTT size must be a power of two")
        self.check(self.lib.g16_ntt(self.h, _ptr(d), log_n, int(inverse), int(coset)))
        return d

    def msm(self, group: int, points: np.ndarray, scalars: np.ndarray):
        pw = 8 if group == 1 else 16
        points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, pw)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = min(points.shape[0], scalars.shape[0])  # msm_bigint zips (prover.rs:66)
        out = np.zeros(pw, dtype=np.uint64)
        inf = C.c_int(0)
        fn = self.lib.g16_msm_g1 if group == 1 else self.lib.g16_msm_g2
        self.check(fn(self.h, _ptr(points), _ptr(scalars), n, _ptr(out), C.byref(inf)))
        return out, bool(inf.value)

    # resident MSM slots (bases stay on the device; the synthetic sweep and the sharded stand-alone MSM)
    def msm_set_bases_dev(self, slot: int, group: int, points_dev: int, n: int, window_bits: int = 0, precompute: bool = False):
        self.check(self.lib.g16_msm_set_bases_dev(self.h, slot, group, C.c_void_p(points_dev), n, window_bits, int(precompute)))
        self._slot_group = getattr(self, "_slot_group", {})
        self._slot_group[slot] = group

    def msm_run_dev(self, slot: int, scalars_dev: int, n: int, want_result: bool = True):
        """Runs the MSM of `slot` over n device scalars; want_result: returns (affine words, is_infinity), else queues only."""
        if not want_result:
            self.check(self.lib.g16_msm_run_dev(self.h, slot, C.c_void_p(scalars_dev), n, None, None))
            return None
        out = np.zeros(16, dtype=np.uint64)
        inf = C.c_int(0)
        self.check(self.lib.g16_msm_run_dev(self.h, slot, C.c_void_p(scalars_dev), n, _ptr(out), C.byref(inf)))
        return out[:8 * getattr(self, "_slot_group", {}).get(slot, 2)].copy(), bool(inf.value)

    def msm_copy_result_dev(self, slot: int, dst_dev: int):
        self.check(self.lib.g16_msm_copy_result_dev(self.h, slot, C.c_void_p(dst_dev)))

    def msm_combine_dev(self, group: int, partials_dev: int, count: int):
        out = np.zeros(8 if group == 1 else 16, dtype=np.uint64)
        inf = C.c_int(0)
        self.check(self.lib.g16_msm_combine_dev(self.h, group, C.c_void_p(partials_dev), count, _ptr(out), C.byref(inf)))
        return out, bool(inf.value)

    def fixed_base_dev(self, group: int, scalars_dev: int, n: int, out_dev: int):
        fn = self.lib.g16_fixed_base_g1_dev if group == 1 else self.lib.g16_fixed_base_g2_dev
        self.check(fn(self.h, C.c_void_p(scalars_dev), n, C.c_void_p(out_dev)))

    def ntt_dev(self, data_dev: int, log_n: int, inverse: bool = False, coset: bool = False):
        self.check(self.lib.g16_ntt_dev(self.h, C.c_void_p(data_dev), log_n, int(inverse), int(coset)))

    def fixed_base(self, group: int, scalars: np.ndarray) -> np.ndarray:
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        pw = 8 if group == 1 else 16
        out = np.zeros((scalars.shape[0], pw), dtype=np.uint64)
        fn = self.lib.g16_fixed_base_g1 if group == 1 else self.lib.g16_fixed_base_g2
        self.check(fn(self.h, _ptr(scalars), scalars.shape[0], _ptr(out)))
        return out

    def set_option(self, key: str, value: int):
        self.check(self.lib.g16_set_option(self.h, key.encode(), int(value)))

    def pow_table(self, base: np.ndarray, scale: np.ndarray, n: int) -> np.ndarray:
        base = np.ascontiguousarray(base, dtype=np.uint64)
        scale = np.ascontiguousarray(scale, dtype=np.uint64)
        out = np.empty((n, 4), dtype=np.uint64)
        self.check(self.lib.g16_pow_table(self.h, _ptr(base), _ptr(scale), n, _ptr(out)))
        return out

    def bench_int_pipe(self, which: int) -> float:
        v = C.c_double(0)
        self.check(self.lib.g16_bench_int_pipe(self.h, which, C.byref(v)))
        return v.value

    def graph_stats(self) -> dict:
        out = np.zeros(3, dtype=np.uint64)
        self.check(self.lib.g16_graph_stats(self.h, _ptr(out)))
        return {"replays": int(out[0]), "captures": int(out[1]), "fallbacks": int(out[2])}

    def launch_count(self) -> int:
        return int(self.lib.g16_launch_count(self.h))

    def sync(self):
        self.check(self.lib.g16_sync(self.h))

    # -- device memory ----------------------------------------------------------------------------------------
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.g16_dev_alloc(self.h, nbytes, C.byref(p)))
        return int(p.value)

    def host_alloc(self, nbytes: int) -> int:
        """Page-locked host memory (g16_host_alloc); returns the address."""
        p = C.c_void_p()
        self.check(self.lib.g16_host_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, p: int):
        self.check(self.lib.g16_host_free(self.h, C.c_void_p(p)))

    def dev_free(self, p: int):
        self.check(self.lib.g16_dev_free(self.h, C.c_void_p(p)))

    def dev_upload(self, p: int, a: np.ndarray):
        a = np.ascontiguousarray(a)
        self.check(self.lib.g16_dev_upload(self.h, C.c_void_p(p), _ptr(a), a.nbytes))

    def dev_download(self, p: int, out: np.ndarray):
        self.check(self.lib.g16_dev_download(self.h, _ptr(out), C.c_void_p(p), out.nbytes))
        return out

    # -- the hot path -------------------------------------------------------------------------------------------
    def load_r1cs(self, nc: int, ni: int, m: int, row_ptr, col, val, encoding=ENC_MONTGOMERY):
        v = R1csView()
        v.num_constraints, v.num_instance, v.num_wires, v.encoding = nc, ni, m, encoding
        keep = []
        for k in range(3):
            rp = np.ascontiguousarray(row_ptr[k], dtype=np.uint64)
            cc = np.ascontiguousarray(col[k], dtype=np.uint32)
            vv = np.ascontiguousarray(val[k], dtype=np.uint64)
            keep += [rp, cc, vv]
            v.row_ptr[k] = C.cast(rp.ctypes.data, _u64p)
            v.col[k] = C.cast(cc.ctypes.data, _u32p)
            v.val[k] = C.cast(vv.ctypes.data, _u64p)
        self._wires = None
        self.check(self.lib.g16_ctx_load_r1cs(self.h, C.byref(v)))
        self._wires = int(m)

    def load_pk(self, arrays: dict, encoding=ENC_MONTGOMERY, shard_rank=0, shard_count=1, precompute=False,
                h_range=None, z_range=None):
        """arrays: a_query, b_g1_query, b_g2_query, h_query, l_query (n x 8 / n x 16 uint64) and the single points
        alpha_g1, beta_g1, delta_g1 (8), beta_g2, delta_g2 (16).  h_range / z_range: explicit [lo, hi) point ranges of
        this rank (g16_ctx_load_pk_ranges) instead of the uniform split."""
        v = PkView()
        keep = {}
        for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
            w = 16 if name == "b_g2_query" else 8
            a = np.ascontiguousarray(arrays[name], dtype=np.uint64).reshape(-1, w)
            keep[name] = a
            setattr(v, name, C.cast(a.ctypes.data, _u64p))
            setattr(v, name.replace("_query", "_len"), a.shape[0])
        for name in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
            a = np.ascontiguousarray(arrays[name], dtype=np.uint64).reshape(-1)
            keep[name] = a
            setattr(v, name, C.cast(a.ctypes.data, _u64p))
        v.encoding = encoding
        if h_range is None and z_range is None:
            self.check(self.lib.g16_ctx_load_pk(self.h, C.byref(v), shard_rank, shard_count, int(precompute)))
        else:
            hr = np.array(h_range, dtype=np.uint64)
            zr = np.array(z_range, dtype=np.uint64)
            self.check(self.lib.g16_ctx_load_pk_ranges(self.h, C.byref(v), shard_rank, shard_count, _ptr(hr), _ptr(zr), int(precompute)))

    def domain_size(self) -> int:
        n = C.c_size_t(0)
        self.check(self.lib.g16_domain_size(self.h, C.byref(n)))
        return int(n.value)

    def _check_z(self, z):
        """A numpy witness must hold exactly num_wires x 4 words (the C side copies m * 32 bytes from the pointer)."""
        if isinstance(z, np.ndarray) and self._wires is not None and z.size != self._wires * 4:
            raise G16Error(ERR_BAD_ARG, f"witness holds {z.size} words, the loaded R1CS has {self._wires} wires x 4")

    def _check_words(self, what: str, a, words: int):
        if isinstance(a, np.ndarray) and a.size * a.itemsize != words * 8:
            raise G16Error(ERR_BAD_ARG, f"{what}: buffer of {a.size * a.itemsize} bytes, expected {words * 8}")

    def witness_map(self, z: np.ndarray, reduction=REDUCTION_LIBSNARK) -> np.ndarray:
        z = np.ascontiguousarray(z, dtype=np.uint64)
        self._check_z(z)
        n = self.domain_size()
        h = np.empty((n, 4), dtype=np.uint64)
        got = C.c_size_t(0)
        self.check(self.lib.g16_witness_map(self.h, _ptr(z), reduction, _ptr(h), n, C.byref(got)))
        return h

    def r1cs_eval(self, z: np.ndarray, nc: int):
        z = np.ascontiguousarray(z, dtype=np.uint64)
        out = [np.zeros((nc, 4), dtype=np.uint64) for _ in range(3)]
        self.check(self.lib.g16_r1cs_eval(self.h, _ptr(z), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    def prove(self, z, r: np.ndarray, s: np.ndarray, reduction=REDUCTION_LIBSNARK) -> ProofOut:
        """z: numpy (m x 4 uint64) or a raw host address (e.g. pinned memory) of m Montgomery elements."""
        if isinstance(z, np.ndarray):
            z = np.ascontiguousarray(z, dtype=np.uint64)
            self._check_z(z)
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        out = ProofOut()
        self.check(self.lib.g16_prove(self.h, _ptr(z), _ptr(r), _ptr(s), reduction, C.byref(out)))
        return out

    def upload_witness(self, z):
        if isinstance(z, np.ndarray):
            z = np.ascontiguousarray(z, dtype=np.uint64)
            self._check_z(z)
        self.check(self.lib.g16_upload_witness(self.h, _ptr(z)))

    def upload_witness_async(self, z, shard_only: bool = False):
        """Stream-ordered upload (no synchronisation): z must stay valid until the context's stream has passed the copy, so
        pass page-locked memory (an address, or an array that outlives the proof).  shard_only: just this rank's slice."""
        if isinstance(z, np.ndarray):
            z = np.ascontiguousarray(z, dtype=np.uint64)
            self._check_z(z)
        self.check(self.lib.g16_upload_witness_async(self.h, _ptr(z), int(shard_only)))

    def upload_witness_dev(self, z_dev: int):
        """The whole witness from device memory (stream-ordered device-to-device copy)."""
        self.check(self.lib.g16_upload_witness_dev(self.h, C.c_void_p(z_dev)))

    def prove_resident(self, r, s, reduction=REDUCTION_LIBSNARK) -> ProofOut:
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        out = ProofOut()
        self.check(self.lib.g16_prove_resident(self.h, _ptr(r), _ptr(s), reduction, C.byref(out)))
        return out

    def prove_shard(self, z, r, s, reduction=REDUCTION_LIBSNARK) -> np.ndarray:
        if isinstance(z, np.ndarray):
            z = np.ascontiguousarray(z, dtype=np.uint64)
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        p = Partial()
        self.check(self.lib.g16_prove_shard(self.h, _ptr(z), _ptr(r), _ptr(s), reduction, C.byref(p)))
        return np.frombuffer(bytes(p), dtype=np.uint64).copy()

    def prove_shard_dev(self, r, s, reduction=REDUCTION_LIBSNARK):
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        self.check(self.lib.g16_prove_shard_dev(self.h, _ptr(r), _ptr(s), reduction))

    def prove_shard_begin_dev(self, r, s, reduction=REDUCTION_LIBSNARK, run_witness_map=True):
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        self.check(self.lib.g16_prove_shard_begin_dev(self.h, _ptr(r), _ptr(s), reduction, int(run_witness_map)))

    def prove_shard_finish_dev(self, h_dev: int = 0, h_first: int = 0, h_count: int = 0):
        """h_dev: device pointer holding h[h_first, h_first + h_count) (0 = this context's own witness-map output)."""
        self.check(self.lib.g16_prove_shard_finish_dev(self.h, C.c_void_p(h_dev) if h_dev else None, h_first, h_count))

    def witness_map_part_dev(self, parts: int):
        """LibsnarkReduction witness map in parts (mask of WM_PART_*), stream-ordered; the whole witness must be resident."""
        self.check(self.lib.g16_witness_map_part_dev(self.h, int(parts)))

    def wm_vector_copy_dev(self, which: int, ext_dev: int, capacity_elems: int, to_ctx: bool):
        """Copies the context's a (0), b (1) or c (2) vector from / to caller-owned device memory (stream-ordered)."""
        self.check(self.lib.g16_wm_vector_copy_dev(self.h, which, C.c_void_p(ext_dev), capacity_elems, int(to_ctx)))

    def copy_h_dev(self, dst_dev: int, capacity_elems: int):
        self.check(self.lib.g16_copy_h_dev(self.h, C.c_void_p(dst_dev), capacity_elems))

    def msm_stats(self, which: int) -> dict:
        """Geometry of the last run of one of the proof's MSMs (0 = h, 1 = l, 2 = a, 3 = b_g1, 4 = b_g2)."""
        out = np.zeros(16, dtype=np.uint64)
        self.check(self.lib.g16_get_msm_stats(self.h, which, _ptr(out), 16))
        levels = int(out[3])
        return {"points": int(out[0]), "window_bits": int(out[1]), "windows": int(out[2]), "levels": levels,
                "buckets": int(out[4]), "shared_digits": bool(out[5]), "tail_tasks": int(out[6]),
                "level_points": [int(v) for v in out[8:8 + levels]]}

    def prove_prepare(self, r, s):
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        self.check(self.lib.g16_prove_prepare(self.h, _ptr(r), _ptr(s)))

    def copy_partial_dev(self, dst_dev: int):
        self.check(self.lib.g16_copy_partial_dev(self.h, C.c_void_p(dst_dev)))

    def partial_dev(self):
        p = C.c_void_p()
        n = C.c_size_t()
        self.check(self.lib.g16_partial_dev(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def prove_combine(self, partials: np.ndarray, r, s) -> ProofOut:
        partials = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, PARTIAL_U64)
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        out = ProofOut()
        self.check(self.lib.g16_prove_combine(self.h, _ptr(partials), partials.shape[0], _ptr(r), _ptr(s), C.byref(out)))
        return out

    def prove_combine_dev(self, dev_partials: int, count: int, r, s) -> ProofOut:
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        out = ProofOut()
        self.check(self.lib.g16_prove_combine_dev(self.h, C.c_void_p(dev_partials), count, _ptr(r), _ptr(s), C.byref(out)))
        return out

    # ---- verification (row f-4) -------------------------------------------------------------------------------------------
    def load_vk(self, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1, encoding=ENC_MONTGOMERY):
        """prepare_verifying_key on the device.  Arrays of uint64 limbs: 8 / 16 / 16 / 16 words and (len, 8)."""
        keep = [np.ascontiguousarray(a, dtype=np.uint64) for a in (alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1)]
        abc = keep[4].reshape(-1, 8)
        v = VkView()
        v.alpha_g1, v.beta_g2, v.gamma_g2, v.delta_g2 = (k.ctypes.data_as(_u64p) for k in keep[:4])
        v.gamma_abc_g1 = abc.ctypes.data_as(_u64p)
        v.gamma_abc_len = abc.shape[0]
        v.encoding = encoding
        self.vk_inputs = None
        self.vk_generation += 1
        self.check(self.lib.g16_ctx_load_vk(self.h, C.byref(v)))
        self.vk_inputs = abc.shape[0] - 1

    def _check_vk_args(self, proofs, inputs, n: int):
        """Sizes of the host buffers of the verify calls against n and the loaded key (the C side reads n * 272 proof bytes
        and n * vk_inputs * 32 input bytes from the raw pointers)."""
        if self.vk_inputs is None:
            raise G16Error(ERR_BAD_ARG, "no verifying key loaded")
        if proofs is not None:
            have = proofs.size * proofs.itemsize if isinstance(proofs, np.ndarray) else C.sizeof(proofs)
            if have != n * C.sizeof(ProofOut):
                raise G16Error(ERR_BAD_ARG, f"proofs buffer of {have} bytes, expected {n} x {C.sizeof(ProofOut)}")
        if inputs is not None:
            self._check_words("public inputs", inputs, n * self.vk_inputs * 4)

    def vk_alpha_beta(self) -> np.ndarray:
        out = np.zeros(48, dtype=np.uint64)
        self.check(self.lib.g16_vk_alpha_beta(self.h, _ptr(out)))
        return out

    def prepare_inputs(self, public_inputs: np.ndarray, n: int) -> np.ndarray:
        x = np.ascontiguousarray(public_inputs, dtype=np.uint64)
        self._check_vk_args(None, x, n)
        out = np.zeros((n, 8), dtype=np.uint64)
        self.check(self.lib.g16_prepare_inputs(self.h, _ptr(x) if x.size else None, n, _ptr(out)))
        return out

    def verify_batch(self, proofs, public_inputs: np.ndarray, n: int) -> np.ndarray:
        """proofs: ctypes array of ProofOut (g16_proof) or a uint8 numpy buffer of n * 272 bytes."""
        x = np.ascontiguousarray(public_inputs, dtype=np.uint64)
        self._check_vk_args(proofs, x, n)
        out = np.zeros(n, dtype=np.uint8)
        pp = _ptr(proofs) if isinstance(proofs, np.ndarray) else C.cast(proofs, C.c_void_p)
        self.check(self.lib.g16_verify_batch(self.h, pp, _ptr(x) if x.size else None, n, _ptr(out)))
        return out

    def verify_batch_prepared(self, proofs, prepared: np.ndarray, n: int) -> np.ndarray:
        x = np.ascontiguousarray(prepared, dtype=np.uint64)
        self._check_vk_args(proofs, None, n)
        self._check_words("prepared inputs", x, n * 8)
        out = np.zeros(n, dtype=np.uint8)
        pp = _ptr(proofs) if isinstance(proofs, np.ndarray) else C.cast(proofs, C.c_void_p)
        self.check(self.lib.g16_verify_batch_prepared(self.h, pp, _ptr(x), n, _ptr(out)))
        return out

    def verify_batch_dev(self, proofs_dev: int, inputs_dev: int, n: int, verdict_dev: int):
        self.check(self.lib.g16_verify_batch_dev(self.h, _ptr(proofs_dev), _ptr(inputs_dev) if inputs_dev else None, n, _ptr(verdict_dev)))

    def pairing(self, g1_points: np.ndarray, g2_points: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, 8)
        q = np.ascontiguousarray(g2_points, dtype=np.uint64).reshape(-1, 16)
        assert p.shape[0] == q.shape[0]
        out = np.zeros((p.shape[0], 48), dtype=np.uint64)
        self.check(self.lib.g16_pairing(self.h, _ptr(p), _ptr(q), p.shape[0], _ptr(out)))
        return out

    def timings(self) -> dict:
        t = Timings()
        self.check(self.lib.g16_get_timings(self.h, C.byref(t)))
        return t.as_dict()
