"""MSM-sharded proving over the GPUs of one box: one process per GPU, every rank holds the contiguous point
range [rank*N/G, (rank+1)*N/G) of each of the five queries, runs the witness map and its five partial MSMs, and the
G partial results (768 bytes per rank: 4 x G1 XYZZ + 1 x G2 XYZZ) are gathered with ONE small collective
(torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests).  Rank 0 adds the partials and assembles.
There is no other data-path collective: the witness goes host -> each GPU directly."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import ffi
from .groth16 import Proof, ProvingKey, ConstraintMatrices, fr_to_mont, R_MOD


def shard_range(total: int, rank: int, world: int):
    """Same split as g16_ctx_load_pk: [total*rank/world, total*(rank+1)/world)."""
    return total * rank // world, total * (rank + 1) // world


def gather_partials(mine, world: int):
    """all_gather of one partial per rank -> (world, PARTIAL_U64) tensor in rank order (device follows `mine`)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world, ffi.PARTIAL_U64), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), mine.contiguous())
    return out


class ShardedProver:
    """Groth16 prover whose MSMs are sharded over torch.distributed ranks (call collectively on every rank)."""

    def __init__(self, pk: ProvingKey, matrices: ConstraintMatrices, device: int, rank: int, world: int, stream: int = 0,
                 precompute: bool = False):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.ctx = ffi.Context(device, stream)
        m = matrices.num_instance_variables + matrices.num_witness_variables
        self.ctx.load_r1cs(matrices.num_constraints, matrices.num_instance_variables, m, matrices.row_ptr, matrices.col,
                           matrices.val, matrices.encoding)
        self.ctx.load_pk(pk.arrays, pk.encoding, rank, world, precompute)
        self.mine = torch.zeros((ffi.PARTIAL_U64,), dtype=torch.int64, device=f"cuda:{device}")

    def prove(self, z_mont, r: int, s: int, reduction=ffi.REDUCTION_LIBSNARK) -> Optional[Proof]:
        """z_mont: (m, 4) uint64 Montgomery witness (same on every rank).  Returns the proof on rank 0, None elsewhere."""
        rr, ss = fr_to_mont([r % R_MOD])[0], fr_to_mont([s % R_MOD])[0]
        self.ctx.upload_witness(z_mont)
        if self.rank == 0:
            self.ctx.prove_prepare(rr, ss)   # overlaps r*delta, s*delta, ... with the shard MSMs
        self.ctx.prove_shard_dev(rr, ss, reduction)
        self.ctx.copy_partial_dev(self.mine.data_ptr())
        allp = gather_partials(self.mine, self.world)
        if self.rank != 0:
            return None
        return Proof.from_ffi(self.ctx.prove_combine_dev(allp.data_ptr(), self.world, rr, ss))

    def close(self):
        self.ctx.close()
