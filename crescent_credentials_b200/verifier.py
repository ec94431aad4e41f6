"""Host-side mirror of the reference's Groth16 verifier over the C ABI ("next" row f-4).

Same names, argument meaning and error behaviour as forks/groth16/src/verifier.rs:
    prepare_verifying_key(vk)                                   :13-20  -> Verifier.prepare_verifying_key
    Groth16::prepare_inputs(pvk, public_inputs)                 :25-39  -> Verifier.prepare_inputs          (MalformedVerifyingKey)
    Groth16::verify_proof_with_prepared_inputs(pvk, proof, g_ic):44-65  -> Verifier.verify_proof_with_prepared_inputs
    Groth16::verify_proof(pvk, proof, public_inputs)            :69-76  -> Verifier.verify_proof            (UnexpectedIdentity)
plus the batched form the device is for: Verifier.verify_proofs(pvk, proofs, inputs) -> one reference verdict per proof.
All arithmetic (pairings, the fixed-base multiplications of prepare_inputs) runs in libg16b200.so on the GPU; there is no
CPU fallback and nothing here imports the oracle."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import ffi
from .groth16 import (Proof, ProvingKey, R_MOD, VerifyingKey, fq_from_mont, fr_to_mont, g1_from_mont, g1_points_to_mont,
                      g2_points_to_mont)


class MalformedVerifyingKey(ffi.G16Error):
    """SynthesisError::MalformedVerifyingKey (verifier.rs:29-31)."""

    def __init__(self, msg="public input count != gamma_abc_g1.len() - 1"):
        super().__init__(ffi.ERR_BAD_ARG, msg)


class UnexpectedIdentity(ffi.G16Error):
    """SynthesisError::UnexpectedIdentity (verifier.rs:62: final_exponentiation returned None)."""

    def __init__(self, msg="final exponentiation of zero"):
        super().__init__(ffi.ERR_BAD_ARG, msg)


@dataclass
class VkArrays:
    """A verifying key as packed limb arrays ready for the C ABI (encoding: ffi.ENC_*)."""
    alpha_g1: np.ndarray
    beta_g2: np.ndarray
    gamma_g2: np.ndarray
    delta_g2: np.ndarray
    gamma_abc_g1: np.ndarray  # (len, 8)
    encoding: int = ffi.ENC_MONTGOMERY

    @staticmethod
    def from_vk(vk: VerifyingKey) -> "VkArrays":
        """VerifyingKey with affine integer points (None = infinity)."""
        return VkArrays(g1_points_to_mont([vk.alpha_g1]).reshape(-1), g2_points_to_mont([vk.beta_g2]).reshape(-1),
                        g2_points_to_mont([vk.gamma_g2]).reshape(-1), g2_points_to_mont([vk.delta_g2]).reshape(-1),
                        g1_points_to_mont(vk.gamma_abc_g1))

    @staticmethod
    def from_pk(pk: ProvingKey) -> "VkArrays":
        """pk.vk of a key read by ProvingKey.deserialize_uncompressed_unchecked (canonical words) or minted by
        generator.generate_parameters_with_qap (Montgomery words)."""
        if hasattr(pk, "raw_vk"):
            return VkArrays(pk.arrays["alpha_g1"], pk.arrays["beta_g2"], pk.raw_vk["gamma_g2"].reshape(-1), pk.arrays["delta_g2"],
                            pk.raw_vk["gamma_abc_g1"], pk.encoding)
        return VkArrays(pk.arrays["alpha_g1"], pk.arrays["beta_g2"], np.asarray(pk.gamma_g2).reshape(-1), pk.arrays["delta_g2"],
                        np.asarray(pk.gamma_abc_g1).reshape(-1, 8), pk.encoding)


@dataclass
class PreparedVerifyingKey:
    """PreparedVerifyingKey (data_structures.rs:62-72).  gamma_g2_neg_pc / delta_g2_neg_pc live on the device; what the host
    keeps is alpha_g1_beta_g2 as 12 canonical Fq integers in ark-serialize order."""
    vk: VkArrays
    alpha_g1_beta_g2: List[int]

    @property
    def num_public_inputs(self) -> int:
        return self.vk.gamma_abc_g1.shape[0] - 1

    def alpha_g1_beta_g2_bytes(self) -> bytes:
        """ark-serialize of the Fq12: 12 x 32 little-endian bytes."""
        return b"".join(int(v).to_bytes(32, "little") for v in self.alpha_g1_beta_g2)


def proofs_to_ffi(proofs: Sequence[Proof]):
    arr = (ffi.ProofOut * len(proofs))()
    for i, p in enumerate(proofs):
        a = g1_points_to_mont([p.a]).reshape(-1)
        b = g2_points_to_mont([p.b]).reshape(-1)
        c = g1_points_to_mont([p.c]).reshape(-1)
        arr[i].a[:] = [int(v) for v in a]
        arr[i].b[:] = [int(v) for v in b]
        arr[i].c[:] = [int(v) for v in c]
        arr[i].a_inf, arr[i].b_inf, arr[i].c_inf = int(p.a is None), int(p.b is None), int(p.c is None)
    return arr


class Verifier:
    """Groth16<Bn254> verification bound to one GPU.  The context keeps the prepared key (keyed by object identity)."""

    def __init__(self, device: int = 0, stream: int = 0, ctx: Optional[ffi.Context] = None):
        self.ctx = ctx if ctx is not None else ffi.Context(device, stream)
        self._own = ctx is None
        # the prepared key this Verifier last loaded (the object, not its id()) and the context's load generation at that
        # moment: another Verifier / caller sharing the context may have replaced the device key since
        self._pvk = None
        self._gen = -1

    def close(self):
        if self._own:
            self.ctx.close()

    # -- verifier.rs:13-20 -------------------------------------------------------------------------------------------------
    def prepare_verifying_key(self, vk) -> PreparedVerifyingKey:
        arrays = vk if isinstance(vk, VkArrays) else (VkArrays.from_pk(vk) if isinstance(vk, ProvingKey) else VkArrays.from_vk(vk))
        self._pvk = None
        self.ctx.load_vk(arrays.alpha_g1, arrays.beta_g2, arrays.gamma_g2, arrays.delta_g2, arrays.gamma_abc_g1, arrays.encoding)
        pvk = PreparedVerifyingKey(arrays, fq_from_mont(self.ctx.vk_alpha_beta().reshape(-1, 4)))
        self._pvk, self._gen = pvk, self.ctx.vk_generation
        return pvk

    def _ensure(self, pvk: PreparedVerifyingKey):
        if self._pvk is not pvk or self._gen != self.ctx.vk_generation:
            a = pvk.vk
            self._pvk = None
            self.ctx.load_vk(a.alpha_g1, a.beta_g2, a.gamma_g2, a.delta_g2, a.gamma_abc_g1, a.encoding)
            self._pvk, self._gen = pvk, self.ctx.vk_generation

    @staticmethod
    def _inputs(pvk: PreparedVerifyingKey, inputs_list) -> np.ndarray:
        k = pvk.num_public_inputs
        flat = []
        for x in inputs_list:
            if len(x) != k:
                raise MalformedVerifyingKey()
            flat += [int(v) % R_MOD for v in x]
        return fr_to_mont(flat).reshape(len(inputs_list), k, 4) if flat else np.zeros((len(inputs_list), 0, 4), dtype=np.uint64)

    # -- verifier.rs:25-39 -------------------------------------------------------------------------------------------------
    def prepare_inputs(self, pvk: PreparedVerifyingKey, public_inputs: Sequence[int]):
        """gamma_abc_g1[0] + sum x_i * gamma_abc_g1[i + 1] as an affine point (None = infinity)."""
        self._ensure(pvk)
        x = self._inputs(pvk, [public_inputs])
        return g1_from_mont(self.ctx.prepare_inputs(x, 1)[0])

    # -- verifier.rs:44-65 -------------------------------------------------------------------------------------------------
    def verify_proof_with_prepared_inputs(self, pvk: PreparedVerifyingKey, proof: Proof, prepared_inputs) -> bool:
        self._ensure(pvk)
        v = self.ctx.verify_batch_prepared(proofs_to_ffi([proof]), g1_points_to_mont([prepared_inputs]), 1)
        return self._verdicts(v)[0]

    # -- verifier.rs:69-76 -------------------------------------------------------------------------------------------------
    def verify_proof(self, pvk: PreparedVerifyingKey, proof: Proof, public_inputs: Sequence[int]) -> bool:
        return self.verify_proofs(pvk, [proof], [public_inputs])[0]

    # -- SNARK trait names (forks/groth16/src/lib.rs:84-96) ---------------------------------------------------------------------
    def process_vk(self, circuit_vk) -> PreparedVerifyingKey:
        return self.prepare_verifying_key(circuit_vk)

    def verify_with_processed_vk(self, circuit_pvk: PreparedVerifyingKey, x: Sequence[int], proof: Proof) -> bool:
        return self.verify_proof(circuit_pvk, proof, x)

    def verify(self, vk, x: Sequence[int], proof: Proof) -> bool:
        """SNARK::verify: process_vk + verify_with_processed_vk."""
        return self.verify_with_processed_vk(self.process_vk(vk), x, proof)

    # -- the batched form: n independent verify_proof calls in one launch ------------------------------------------------------
    def verify_proofs(self, pvk: PreparedVerifyingKey, proofs: Sequence[Proof], public_inputs: Sequence[Sequence[int]]) -> List[bool]:
        if len(proofs) != len(public_inputs):
            raise ffi.G16Error(ffi.ERR_BAD_ARG, "one public-input vector per proof")
        self._ensure(pvk)
        if not proofs:
            return []
        x = self._inputs(pvk, public_inputs)
        return self._verdicts(self.ctx.verify_batch(proofs_to_ffi(proofs), x, len(proofs)), batched=len(proofs) > 1)

    @staticmethod
    def _verdicts(v: np.ndarray, batched: bool = False) -> List[bool]:
        """Single-proof calls raise UnexpectedIdentity like verifier.rs:62; in a batch one such proof must not discard the
        other verdicts, so it reads as rejected (False) there."""
        if not batched and (v == ffi.VERDICT_UNEXPECTED_IDENTITY).any():
            raise UnexpectedIdentity()
        return [bool(t == ffi.VERDICT_ACCEPT) for t in v]


def verify_proofs_replicated(verifier, pvk, proofs: Sequence, public_inputs: Sequence, rank: int, world: int,
                             device=None) -> List[bool]:
    """Verification over several GPUs.  The unit of work is one (proof, inputs) pair and pairs are independent (verifier.rs:44-65
    reads nothing but the pair and the key), so there is no exchange step to put a collective on: every rank is a replica
    holding the same prepared key, verifies the contiguous slice shard_range(n, rank, world) of the batch on its own GPU with
    `verifier.verify_proofs`, and ONE all_gather of the verdict bytes (n bytes in total) hands every rank the full list in
    proof order.  Call collectively (torch.distributed initialised; `device` = where the gathered tensor lives: the rank's
    CUDA device under NCCL, None / "cpu" under gloo)."""
    import torch
    import torch.distributed as dist
    from .sharded import shard_range
    n = len(proofs)
    if n != len(public_inputs):
        raise ffi.G16Error(ffi.ERR_BAD_ARG, "one public-input vector per proof")
    if world <= 1:
        return verifier.verify_proofs(pvk, proofs, public_inputs)
    lo, hi = shard_range(n, rank, world)
    mine = verifier.verify_proofs(pvk, proofs[lo:hi], public_inputs[lo:hi]) if hi > lo else []
    chunk = -(-n // world) if n else 0          # slices differ by at most one pair: pad to the longest
    if chunk == 0:
        return []
    buf = torch.full((chunk,), 255, dtype=torch.uint8)
    if mine:
        buf[:len(mine)] = torch.tensor([1 if v else 0 for v in mine], dtype=torch.uint8)
    buf = buf.to(device) if device is not None else buf
    out = torch.empty((world * chunk,), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().view(world, chunk)
    verdicts: List[bool] = []
    for r in range(world):
        rlo, rhi = shard_range(n, r, world)
        row = out[r, :rhi - rlo].tolist()
        if any(v > 1 for v in row):
            raise ffi.G16Error(ffi.ERR_BAD_ARG, f"rank {r} returned no verdict for part of its slice")
        verdicts += [v == 1 for v in row]
    return verdicts
