"""Host-side mirror of the ark-groth16 API surface Crescent uses (forks/groth16/src/{lib,prover,r1cs_to_qap,
data_structures}.rs), sitting directly on the C ABI.  Names and argument meaning follow the reference:

    Groth16.create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment)
        -> prover.rs:26-51
    Groth16.create_random_proof_with_reduction / create_proof_with_reduction_no_zk   -> prover.rs:142-172
    LibsnarkReduction / CircomReduction .witness_map_from_matrices                   -> r1cs_to_qap.rs:150-213, qap.rs:25-90
    ProvingKey / VerifyingKey / Proof (+ ark-serialize canonical (de)serialisation)  -> data_structures.rs

Everything arithmetic runs in libg16b200.so on the GPU.  Python integers are used only to marshal bytes
(Montgomery <-> canonical encodings, serialisation flag bits) -- there is no CPU prover in here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import ffi

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_MASK64 = (1 << 64) - 1
_RINV_R = pow(1 << 256, -1, R_MOD)
_RINV_Q = pow(1 << 256, -1, Q_MOD)


# ---- byte marshalling ---------------------------------------------------------------------------------------------
def ints_to_limbs(vals: Sequence[int]) -> np.ndarray:
    """list of ints < 2^256 -> (n, 4) uint64 little-endian limbs."""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def limbs_to_ints(a: np.ndarray) -> List[int]:
    raw = np.ascontiguousarray(a, dtype="<u8").tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def fr_to_mont(vals: Sequence[int]) -> np.ndarray:
    return ints_to_limbs([(v << 256) % R_MOD for v in vals])


def fr_from_mont(a: np.ndarray) -> List[int]:
    return [v * _RINV_R % R_MOD for v in limbs_to_ints(a)]


def fq_to_mont(vals: Sequence[int]) -> np.ndarray:
    return ints_to_limbs([(v << 256) % Q_MOD for v in vals])


def fq_from_mont(a: np.ndarray) -> List[int]:
    return [v * _RINV_Q % Q_MOD for v in limbs_to_ints(a)]


def g1_points_to_mont(points: Sequence[Optional[Tuple[int, int]]]) -> np.ndarray:
    """affine (x, y) tuples / None for infinity -> (n, 8) uint64, infinity = zeros."""
    flat = []
    for p in points:
        flat += [0, 0] if p is None else [p[0], p[1]]
    return fq_to_mont(flat).reshape(-1, 8)


def g2_points_to_mont(points) -> np.ndarray:
    flat = []
    for p in points:
        flat += [0, 0, 0, 0] if p is None else [p[0][0], p[0][1], p[1][0], p[1][1]]
    return fq_to_mont(flat).reshape(-1, 16)


def g1_from_mont(a: np.ndarray, inf: bool = False):
    if inf:
        return None
    v = fq_from_mont(np.asarray(a).reshape(-1, 4))
    return None if v[0] == 0 and v[1] == 0 else (v[0], v[1])


def g2_from_mont(a: np.ndarray, inf: bool = False):
    if inf:
        return None
    v = fq_from_mont(np.asarray(a).reshape(-1, 4))
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))


# ---- ark-serialize 0.4 canonical encoding (little-endian; SW flags in the two top bits of the last byte) -------------
def _fq_neg_flag(y: int) -> int:
    return 0x80 if y > (Q_MOD - y) % Q_MOD else 0


def _fq2_neg_flag(y) -> int:
    ny = ((-y[0]) % Q_MOD, (-y[1]) % Q_MOD)
    return 0x80 if (y[1], y[0]) > (ny[1], ny[0]) else 0


def _ser_g1(p, compressed: bool) -> bytes:
    if p is None:
        b = bytearray(32 if compressed else 64)
        b[-1] |= 0x40
        return bytes(b)
    b = bytearray(p[0].to_bytes(32, "little") + (b"" if compressed else p[1].to_bytes(32, "little")))
    b[-1] |= _fq_neg_flag(p[1])
    return bytes(b)


def _ser_g2(p, compressed: bool) -> bytes:
    if p is None:
        b = bytearray(64 if compressed else 128)
        b[-1] |= 0x40
        return bytes(b)
    x, y = p
    b = bytearray(x[0].to_bytes(32, "little") + x[1].to_bytes(32, "little"))
    if not compressed:
        b += y[0].to_bytes(32, "little") + y[1].to_bytes(32, "little")
    b[-1] |= _fq2_neg_flag(y)
    return bytes(b)


@dataclass
class Proof:
    """Proof{a, b, c} (data_structures.rs:7-14); coordinates are canonical integers, None = infinity."""
    a: Optional[Tuple[int, int]]
    b: Optional[Tuple[Tuple[int, int], Tuple[int, int]]]
    c: Optional[Tuple[int, int]]

    def serialize_compressed(self) -> bytes:  # 128 bytes
        return _ser_g1(self.a, True) + _ser_g2(self.b, True) + _ser_g1(self.c, True)

    def serialize_uncompressed(self) -> bytes:  # 256 bytes; what creds/src/utils.rs:140-152 persists
        return _ser_g1(self.a, False) + _ser_g2(self.b, False) + _ser_g1(self.c, False)

    @staticmethod
    def from_ffi(p: ffi.ProofOut) -> "Proof":
        return Proof(g1_from_mont(np.array(p.a[:], dtype=np.uint64), bool(p.a_inf)),
                     g2_from_mont(np.array(p.b[:], dtype=np.uint64), bool(p.b_inf)),
                     g1_from_mont(np.array(p.c[:], dtype=np.uint64), bool(p.c_inf)))


@dataclass
class ConstraintMatrices:
    """ark-relations ConstraintMatrices<Fr> flattened to CSR.  Coefficients are Montgomery limbs."""
    num_instance_variables: int
    num_witness_variables: int
    num_constraints: int
    row_ptr: List[np.ndarray]  # 3 x uint64[nc + 1]
    col: List[np.ndarray]      # 3 x uint32[nnz]
    val: List[np.ndarray]      # 3 x uint64[nnz, 4]
    encoding: int = ffi.ENC_MONTGOMERY

    @property
    def a_num_non_zero(self):
        return int(self.row_ptr[0][-1])

    @property
    def b_num_non_zero(self):
        return int(self.row_ptr[1][-1])

    @property
    def c_num_non_zero(self):
        return int(self.row_ptr[2][-1])

    @staticmethod
    def from_rows(num_instance: int, num_witness: int, a, b, c) -> "ConstraintMatrices":
        """a, b, c: Vec<Vec<(Fr, usize)>> as lists of rows of (coeff_int, column)."""
        rp, cc, vv = [], [], []
        for rows in (a, b, c):
            ptr = np.zeros(len(rows) + 1, dtype=np.uint64)
            cols, vals = [], []
            for i, row in enumerate(rows):
                for coeff, idx in row:
                    cols.append(idx)
                    vals.append(coeff % R_MOD)
                ptr[i + 1] = len(cols)
            rp.append(ptr)
            cc.append(np.array(cols, dtype=np.uint32))
            vv.append(fr_to_mont(vals) if vals else np.zeros((0, 4), dtype=np.uint64))
        return ConstraintMatrices(num_instance, num_witness, len(a), rp, cc, vv)


@dataclass
class VerifyingKey:
    alpha_g1: object = None
    beta_g2: object = None
    gamma_g2: object = None
    delta_g1: object = None  # fork-only field (data_structures.rs:39)
    delta_g2: object = None
    gamma_abc_g1: list = field(default_factory=list)


class ProvingKey:
    """ProvingKey<Bn254> (data_structures.rs:101-118) held as packed limb arrays ready for the C ABI."""

    def __init__(self, arrays: dict, encoding: int, vk: Optional[VerifyingKey] = None):
        self.arrays = arrays
        self.encoding = encoding
        self.vk = vk

    @staticmethod
    def from_points(vk_alpha_g1, beta_g1, delta_g1, vk_beta_g2, vk_delta_g2, a_query, b_g1_query, b_g2_query, h_query,
                    l_query, vk: Optional[VerifyingKey] = None) -> "ProvingKey":
        arr = dict(
            alpha_g1=g1_points_to_mont([vk_alpha_g1]).reshape(-1), beta_g1=g1_points_to_mont([beta_g1]).reshape(-1),
            delta_g1=g1_points_to_mont([delta_g1]).reshape(-1), beta_g2=g2_points_to_mont([vk_beta_g2]).reshape(-1),
            delta_g2=g2_points_to_mont([vk_delta_g2]).reshape(-1), a_query=g1_points_to_mont(a_query),
            b_g1_query=g1_points_to_mont(b_g1_query), b_g2_query=g2_points_to_mont(b_g2_query),
            h_query=g1_points_to_mont(h_query), l_query=g1_points_to_mont(l_query))
        return ProvingKey(arr, ffi.ENC_MONTGOMERY, vk)

    @staticmethod
    def deserialize_uncompressed_unchecked(buf: bytes) -> "ProvingKey":
        """Reads arkworks' uncompressed ProvingKey bytes -- what Crescent keeps in cache/prover_params.bin
        (creds/src/utils.rs:179-189, layout data_structures.rs:31-44,101-118).  Coordinates stay canonical
        little-endian words (flag bits cleared, infinity -> zeros); the library converts to Montgomery on the GPU."""
        mv = memoryview(buf)
        off = 0

        def g1s(count):
            nonlocal off
            a = np.frombuffer(mv[off:off + 64 * count], dtype="<u8").reshape(count, 8).copy()
            off += 64 * count
            flags = (a[:, 7] >> np.uint64(62)).astype(np.uint8)
            a[:, 7] &= np.uint64((1 << 62) - 1)
            a[(flags & 1) == 1] = 0
            return a

        def g2s(count):
            nonlocal off
            a = np.frombuffer(mv[off:off + 128 * count], dtype="<u8").reshape(count, 16).copy()
            off += 128 * count
            flags = (a[:, 15] >> np.uint64(62)).astype(np.uint8)
            a[:, 15] &= np.uint64((1 << 62) - 1)
            a[(flags & 1) == 1] = 0
            return a

        def vec_len():
            nonlocal off
            (n,) = struct.unpack_from("<Q", mv, off)
            off += 8
            return n

        arr = {}
        arr["alpha_g1"] = g1s(1).reshape(-1)
        arr["beta_g2"] = g2s(1).reshape(-1)
        gamma_g2 = g2s(1)
        vk_delta_g1 = g1s(1)
        arr["delta_g2"] = g2s(1).reshape(-1)
        gamma_abc = g1s(vec_len())
        arr["beta_g1"] = g1s(1).reshape(-1)
        arr["delta_g1"] = g1s(1).reshape(-1)
        arr["a_query"] = g1s(vec_len())
        arr["b_g1_query"] = g1s(vec_len())
        arr["b_g2_query"] = g2s(vec_len())
        arr["h_query"] = g1s(vec_len())
        arr["l_query"] = g1s(vec_len())
        if off != len(buf):
            raise ValueError(f"trailing bytes in ProvingKey: parsed {off} of {len(buf)}")
        pk = ProvingKey(arr, ffi.ENC_CANONICAL)
        pk.raw_vk = dict(gamma_g2=gamma_g2, delta_g1=vk_delta_g1, gamma_abc_g1=gamma_abc)
        return pk


    # -- CanonicalSerialize (uncompressed), the bytes creds/src/utils.rs:140-152 writes to prover_params.bin ------------------
    def vk_parts(self):
        """(gamma_g2 (16,), vk.delta_g1 (8,), gamma_abc_g1 (len, 8)) in this key's encoding, or None when the key carries no
        verifying-key part (a hand-assembled prover-only key)."""
        if hasattr(self, "raw_vk"):
            rv = self.raw_vk
            return (np.asarray(rv["gamma_g2"]).reshape(-1), np.asarray(rv["delta_g1"]).reshape(-1),
                    np.asarray(rv["gamma_abc_g1"]).reshape(-1, 8))
        if getattr(self, "gamma_g2", None) is not None and getattr(self, "gamma_abc_g1", None) is not None:
            return (np.asarray(self.gamma_g2).reshape(-1), np.asarray(self.arrays["delta_g1"]).reshape(-1),
                    np.asarray(self.gamma_abc_g1).reshape(-1, 8))
        return None

    def serialize_uncompressed(self, ctx: Optional["ffi.Context"] = None) -> bytes:
        """ark-serialize 0.4 uncompressed bytes of ProvingKey{vk{alpha_g1, beta_g2, gamma_g2, delta_g1, delta_g2, gamma_abc_g1},
        beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query} (data_structures.rs:31-44,101-118): coordinates
        as 32-byte little-endian canonical integers, the y-sign / infinity flags in the two top bits of the last byte, every
        Vec prefixed by its u64 length.  `ctx` (optional) converts Montgomery words on the GPU; without it the conversion is
        byte marshalling with Python integers (fine for the test-size keys)."""
        parts = self.vk_parts()
        if parts is None:
            raise ValueError("this ProvingKey has no verifying-key part (gamma_g2, gamma_abc_g1) to serialise")
        gamma_g2, vk_delta_g1, gamma_abc = parts
        canon = lambda a: _to_canonical_words(np.asarray(a, dtype=np.uint64), self.encoding, ctx)
        g1 = lambda a: _ser_points_uncompressed(canon(a).reshape(-1, 8), 1)
        g2 = lambda a: _ser_points_uncompressed(canon(a).reshape(-1, 16), 2)
        vec = lambda a, f, w: struct.pack("<Q", np.asarray(a).reshape(-1, w).shape[0]) + f(a)
        A = self.arrays
        return b"".join([g1(A["alpha_g1"]), g2(A["beta_g2"]), g2(gamma_g2), g1(vk_delta_g1), g2(A["delta_g2"]),
                         vec(gamma_abc, g1, 8), g1(A["beta_g1"]), g1(A["delta_g1"]), vec(A["a_query"], g1, 8),
                         vec(A["b_g1_query"], g1, 8), vec(A["b_g2_query"], g2, 16), vec(A["h_query"], g1, 8),
                         vec(A["l_query"], g1, 8)])


_Q_HALF = (Q_MOD - 1) // 2


def _to_canonical_words(a: np.ndarray, encoding: int, ctx=None) -> np.ndarray:
    if encoding == ffi.ENC_CANONICAL or a.size == 0:
        return np.ascontiguousarray(a, dtype=np.uint64)
    flat = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    if ctx is not None:
        return ctx.field_op(ffi.FIELD_FQ, ffi.OP_FROM_MONT, flat).reshape(a.shape)
    return ints_to_limbs(fq_from_mont(flat)).reshape(a.shape)


def _gt_half(y: np.ndarray) -> np.ndarray:
    """y (n, 4) canonical little-endian words -> bool[n]: y > (q - 1) / 2, i.e. y is the larger of {y, -y}."""
    half = [(_Q_HALF >> (64 * k)) & _MASK64 for k in range(4)]
    gt = np.zeros(y.shape[0], dtype=bool)
    eq = np.ones(y.shape[0], dtype=bool)
    for k in (3, 2, 1, 0):
        h = np.uint64(half[k])
        gt |= eq & (y[:, k] > h)
        eq &= y[:, k] == h
    return gt


def _ser_points_uncompressed(pts: np.ndarray, group: int) -> bytes:
    """(n, 8) / (n, 16) canonical words, infinity = all zero -> n * 64 / n * 128 bytes with the SW flags (ark-ec 0.4
    SWFlags: bit 7 = y is the lexicographically larger root, bit 6 = infinity) on the last byte."""
    out = np.array(pts, dtype="<u8", order="C", copy=True)
    n = out.shape[0]
    if n == 0:
        return b""
    inf = ~out.any(axis=1)
    if group == 1:
        neg = _gt_half(out[:, 4:8])
    else:  # Fq2 ordering: c1 first, then c0
        c0, c1 = out[:, 8:12], out[:, 12:16]
        neg = _gt_half(c1) | (~c1.any(axis=1) & _gt_half(c0))
    last = out.shape[1] - 1
    out[neg & ~inf, last] |= np.uint64(1 << 63)
    out[inf, last] |= np.uint64(1 << 62)
    return out.tobytes()


class LibsnarkReduction:
    """R1CSToQAP implementation selector (forks/groth16/src/r1cs_to_qap.rs:100-226); the default of Groth16<E, QAP>
    (forks/groth16/src/lib.rs:55) and what Crescent uses (creds/src/lib.rs:229,283)."""
    ID = ffi.REDUCTION_LIBSNARK


class CircomReduction:
    """forks/circom-compat/src/circom/qap.rs:15-108."""
    ID = ffi.REDUCTION_CIRCOM


def sample_fr(rng) -> int:
    """Fr::rand shape (ark-ff 0.4): 4 x next_u64, top limb masked to 254 bits, rejection above the modulus; the accepted
    integer is the Montgomery representation.  `rng` needs a .getrandbits(64)-style next_u64.  r and s are the
    zero-knowledge blinders of the proof (prover.rs:151-152): production callers must pass a CSPRNG -- `secrets.SystemRandom()`
    or the ChaCha `rng.StdRng` mirror seeded from OS entropy; `random.Random` is for reproducible tests only."""
    while True:
        v = 0
        for k in range(4):
            v |= rng.getrandbits(64) << (64 * k)
        v &= (1 << 254) - 1
        if v < R_MOD:
            return v * _RINV_R % R_MOD


class Groth16:
    """Groth16<Bn254, QAP> prover bound to one GPU.  The context keeps device copies of the proving key queries and the
    CSR matrices (keyed by object identity) so that back-to-back proofs pay only for the witness upload."""

    def __init__(self, device: int = 0, qap=LibsnarkReduction, stream: int = 0, shard_rank: int = 0, shard_count: int = 1,
                 precompute: bool = False):
        self.ctx = ffi.Context(device, stream)
        self.qap = qap
        self.shard_rank, self.shard_count, self.precompute = shard_rank, shard_count, precompute
        # the loaded objects themselves, not their id(): CPython reuses an id once an object is freed, and a different key at
        # the same address must not be mistaken for the resident one
        self._pk = None
        self._matrices = None

    def close(self):
        self.ctx.close()

    # -- resident state ----------------------------------------------------------------------------------------------
    def _ensure_matrices(self, matrices: ConstraintMatrices, num_inputs: int, num_constraints: int):
        if num_inputs != matrices.num_instance_variables or num_constraints != matrices.num_constraints:
            raise ffi.G16Error(ffi.ERR_BAD_ARG, "num_inputs / num_constraints disagree with the matrices")
        if self._matrices is not matrices:
            m = matrices.num_instance_variables + matrices.num_witness_variables
            self._matrices = None
            self.ctx.load_r1cs(matrices.num_constraints, matrices.num_instance_variables, m, matrices.row_ptr, matrices.col,
                               matrices.val, matrices.encoding)
            self._matrices = matrices

    def _ensure_pk(self, pk: ProvingKey):
        if self._pk is not pk:
            self._pk = None
            self.ctx.load_pk(pk.arrays, pk.encoding, self.shard_rank, self.shard_count, self.precompute)
            self._pk = pk

    @staticmethod
    def _assignment(full_assignment) -> np.ndarray:
        if isinstance(full_assignment, np.ndarray):
            return np.ascontiguousarray(full_assignment, dtype=np.uint64).reshape(-1, 4)
        return fr_to_mont(full_assignment)

    # -- reference API ---------------------------------------------------------------------------------------------------
    def witness_map_from_matrices(self, matrices, num_inputs, num_constraints, full_assignment) -> List[int]:
        self._ensure_matrices(matrices, num_inputs, num_constraints)
        z = self._assignment(full_assignment)
        if z.shape[0] != matrices.num_instance_variables + matrices.num_witness_variables:
            raise ffi.G16Error(ffi.ERR_BAD_ARG, "full_assignment length != number of wires")
        return fr_from_mont(self.ctx.witness_map(z, self.qap.ID))

    def create_proof_with_reduction_and_matrices(self, pk: ProvingKey, r: int, s: int, matrices: ConstraintMatrices,
                                                 num_inputs: int, num_constraints: int, full_assignment) -> Proof:
        self._ensure_matrices(matrices, num_inputs, num_constraints)
        self._ensure_pk(pk)
        z = self._assignment(full_assignment)
        if z.shape[0] != matrices.num_instance_variables + matrices.num_witness_variables:
            raise ffi.G16Error(ffi.ERR_BAD_ARG, "full_assignment length != number of wires")
        rr, ss = fr_to_mont([r % R_MOD])[0], fr_to_mont([s % R_MOD])[0]
        if self.shard_count == 1:
            return Proof.from_ffi(self.ctx.prove(z, rr, ss, self.qap.ID))
        raise ffi.G16Error(ffi.ERR_BAD_ARG, "sharded prover: use crescent_credentials_b200.sharded.ShardedProver")

    def create_random_proof_with_reduction(self, pk, matrices, num_inputs, num_constraints, full_assignment, rng) -> Proof:
        """prover.rs:142-154: r then s sampled from rng, then the deterministic prover."""
        r = sample_fr(rng)
        s = sample_fr(rng)
        return self.create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment)

    prove = create_random_proof_with_reduction  # SNARK::prove (forks/groth16/src/lib.rs:76-82)

    def create_proof_with_reduction_no_zk(self, pk, matrices, num_inputs, num_constraints, full_assignment) -> Proof:
        """prover.rs:159-172 (r = s = 0; the B-in-G1 MSM is skipped as at prover.rs:102)."""
        return self.create_proof_with_reduction_and_matrices(pk, 0, 0, matrices, num_inputs, num_constraints, full_assignment)

    def timings(self) -> dict:
        return self.ctx.timings()
