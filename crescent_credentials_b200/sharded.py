"""MSM-sharded proving over the GPUs of one box: one process per GPU, every rank holds a contiguous point
range of each of the five queries and runs its five partial MSMs; the G partial results (896 bytes per rank: 5 x G1 XYZZ
+ 1 x G2 XYZZ) are gathered with ONE small collective (torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU
tests) and rank 0 adds the partials and assembles.

Two plans:
  uniform    every rank takes [rank*N/G, (rank+1)*N/G) of every query and runs the whole witness map itself (no h exchange).
  staggered  (default for G > 1) the witness map does not shard (north star: "NTT and witness_map stay on one GPU"), so
             replicating it puts ~4 ms on every rank's critical path.  Instead rank 0 alone runs it while the other ranks
             spend that time on a LARGER share of the z-only MSMs (a, l, b_g1, b_g2); rank 0 then scatters the h
             coefficients (n*32/G bytes per peer, one NCCL scatter over NVLink) and every rank finishes with its h-MSM
             range.  Rank 0's share f0 of the wire MSMs balances  wm + f0*Z  against  (1 - f0)*Z/(G - 1).
The witness goes host -> each GPU directly in both plans; a rank that does not run the witness map uploads only the slice of z
its wire MSMs read (g16_upload_witness_async(shard_only)), stream-ordered, without a host synchronisation."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import ffi
from .groth16 import Proof, ProvingKey, ConstraintMatrices, fr_to_mont, R_MOD

# witness-map time over the time of the four z-only MSMs on one B200; only the balance of the staggered plan depends on it,
# never a result.  0.28 in round 1 (4.25 ms vs ~15 ms); at N = 2 that left rank 0 done at 14.4 ms and rank 1 at 16.2 ms
# (profiles/r02_bench_n2.json), i.e. rank 0 should take 0.40 of the wire MSMs instead of 0.36: 0.20.  The smaller automatic
# window then made the wire MSMs ~9 % cheaper: with 0.20 rank 0 was done at 14.46 ms and rank 1 at 13.87 ms at N = 2
# (profiles/r02_bench_n2_final.json: 3.23 + (2 f0 - 1) Z = 0.59 with f0 = 0.40 gives Z = 13.2 ms, ratio 0.245); the N = 4 stage
# timings of the same day point at ~0.23.  0.235: rank 0 takes 0.38 of the wire MSMs at N = 2, 0.074 at N = 4, none from N = 5 on.
WM_OVER_Z = 0.235


def shard_range(total: int, rank: int, world: int):
    """Same split as g16_ctx_load_pk: [total*rank/world, total*(rank+1)/world)."""
    return total * rank // world, total * (rank + 1) // world


@dataclass
class ShardPlan:
    world: int
    wm_rank: int                       # rank that runs the witness map; -1 = every rank (uniform plan)
    h_chunk: int                       # staggered: h coefficients per rank in the scatter (equal chunks)
    h_ranges: List[Tuple[int, int]]    # per rank [lo, hi) of h_query
    z_ranges: List[Tuple[int, int]]    # per rank [lo, hi) of a_query[1..] / b_g1_query[1..] / b_g2_query[1..]
    wm2_rank: int = -1                 # "wm_split": this rank runs the b pipeline of the witness map for wm_rank ...
    wm3_rank: int = -1                 # ... and this one the c pipeline (both beside their own wire MSMs)

    @property
    def staggered(self) -> bool:
        return self.wm_rank >= 0


def uniform_plan(h_len: int, m1: int, world: int) -> ShardPlan:
    return ShardPlan(world, -1, 0, [shard_range(h_len, r, world) for r in range(world)],
                     [shard_range(m1, r, world) for r in range(world)])


def rank0_wire_share(world: int, wm_over_z: float = WM_OVER_Z) -> float:
    """f0 with  wm + f0*Z == (1 - f0)*Z/(G - 1)  (clamped at 0: from ~5 ranks on rank 0 takes no wire MSM work at all)."""
    if world <= 1:
        return 1.0
    per_other = 1.0 / (world - 1)
    return max(0.0, (per_other - wm_over_z) / (1.0 + per_other))


def staggered_plan(h_len: int, m1: int, world: int, rank0_share: Optional[float] = None,
                   wm_split: Optional[bool] = None) -> ShardPlan:
    """wm_split (default off; needs world >= 3): the a, b and c pipelines of the witness map are independent until the last
    transform, so rank 1 computes the b pipeline and rank 2 the c pipeline beside their own wire MSMs (those chains are
    latency-bound and leave the multiplier mostly idle), each sends its vector point to point over NVLink (32 n bytes) and rank 0
    -- which computes the a pipeline meanwhile -- runs the last transform.  Measured at 8 GPUs in DESIGN.md section 4."""
    if world == 1:
        return uniform_plan(h_len, m1, 1)
    split = bool(wm_split) and world >= 3
    f0 = rank0_wire_share(world) if rank0_share is None else min(max(float(rank0_share), 0.0), 1.0)
    cut = int(round(m1 * f0))
    rest = m1 - cut
    z = [(0, cut)] + [(cut + rest * (r - 1) // (world - 1), cut + rest * r // (world - 1)) for r in range(1, world)]
    chunk = -(-h_len // world)
    h = [(min(r * chunk, h_len), min((r + 1) * chunk, h_len)) for r in range(world)]
    return ShardPlan(world, 0, chunk, h, z, wm2_rank=1 if split else -1, wm3_rank=2 if split else -1)


def gather_partials(mine, world: int):
    """all_gather of one partial per rank -> (world, PARTIAL_U64) tensor in rank order (device follows `mine`)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world, ffi.PARTIAL_U64), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), mine.contiguous())
    return out


def scatter_h(h_all, h_mine, plan: ShardPlan, rank: int):
    """Rank plan.wm_rank holds h in h_all ((>= world * h_chunk, 4) int64); every rank receives its chunk in h_mine."""
    import torch.distributed as dist
    c = plan.h_chunk
    chunks = [h_all[r * c:(r + 1) * c].view(-1) for r in range(plan.world)] if rank == plan.wm_rank else None
    dist.scatter(h_mine.view(-1), chunks, src=plan.wm_rank)


class ShardedProver:
    """Groth16 prover whose MSMs are sharded over torch.distributed ranks (call collectively on every rank).

    Stream discipline: the library's asynchronous entry points (`*_dev`) run on the context's main stream and the
    torch.distributed collectives on torch's *current* stream; the two are only ordered if they are the same stream.  The
    prover therefore owns (or is given) ONE torch.cuda.Stream, hands its raw handle to the context as the main stream, and
    issues every collective inside `torch.cuda.stream(self.stream)`: scatter -> h MSM -> partial copy -> all_gather ->
    combine is a single stream-ordered chain, with no host synchronisation before the proof read-back on rank 0."""

    def __init__(self, pk: ProvingKey, matrices: ConstraintMatrices, device: int, rank: int, world: int, stream=None,
                 precompute: bool = False, plan: Optional[ShardPlan] = None, ctx: Optional[ffi.Context] = None,
                 gather_upload: bool = True):
        """stream: a torch.cuda.Stream (default: a new one on `device`).  ctx: an existing context that was created on that
        stream's handle (its R1CS / key are loaded here); default: a new context."""
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        dev = torch.device("cuda", device)
        if ctx is not None and stream is None:
            raise ValueError("ShardedProver(ctx=...) needs the torch stream the context was created on")
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        if not isinstance(self.stream, torch.cuda.Stream):
            raise TypeError("stream must be a torch.cuda.Stream: collectives and library calls have to share it")
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else ffi.Context(device, self.stream.cuda_stream)
        m = matrices.num_instance_variables + matrices.num_witness_variables
        self.m = m
        self.ctx.load_r1cs(matrices.num_constraints, matrices.num_instance_variables, m, matrices.row_ptr, matrices.col,
                           matrices.val, matrices.encoding)
        h_len = int(np.asarray(pk.arrays["h_query"]).reshape(-1, 8).shape[0])
        m1 = int(np.asarray(pk.arrays["a_query"]).reshape(-1, 8).shape[0]) - 1
        self.plan = plan if plan is not None else staggered_plan(h_len, m1, world)
        assert self.plan.world == world
        self.ctx.load_pk(pk.arrays, pk.encoding, rank, world, precompute, h_range=self.plan.h_ranges[rank],
                         z_range=self.plan.z_ranges[rank])
        with torch.cuda.stream(self.stream):
            self.mine = torch.zeros((ffi.PARTIAL_U64,), dtype=torch.int64, device=dev)
            self.gathered = torch.zeros((world, ffi.PARTIAL_U64), dtype=torch.int64, device=dev)
            self.h_all = self.h_mine = None
            if self.plan.staggered:
                n = self.ctx.domain_size()
                cap = max(n, world * self.plan.h_chunk)
                if rank == self.plan.wm_rank:
                    self.h_all = torch.zeros((cap, 4), dtype=torch.int64, device=dev)
                self.h_mine = torch.zeros((self.plan.h_chunk, 4), dtype=torch.int64, device=dev)
            # gathered upload (staggered plan): every rank's 1/G chunk of z and the assembled vector
            self.gather_upload = bool(gather_upload) and world > 1 and self.plan.staggered
            self.z_chunk_len = -(-m // world)
            self.z_chunk = self.z_all = None
            if self.gather_upload:
                self.z_chunk = torch.zeros((self.z_chunk_len, 4), dtype=torch.int64, device=dev)
                self.z_all = torch.zeros((world * self.z_chunk_len, 4), dtype=torch.int64, device=dev)
            # wm_split: the b and c vectors travel from ranks wm2 / wm3 to rank wm through these
            self.vb = self.vc = None
            if self.plan.wm2_rank >= 0 and rank in (self.plan.wm_rank, self.plan.wm2_rank, self.plan.wm3_rank):
                n = self.ctx.domain_size()
                if rank in (self.plan.wm_rank, self.plan.wm2_rank):
                    self.vb = torch.zeros((n, 4), dtype=torch.int64, device=dev)
                if rank in (self.plan.wm_rank, self.plan.wm3_rank):
                    self.vc = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        self.stream.synchronize()
        self.z_pin = None          # page-locked staging of the witness for prove()
        self._z_done = None        # event: the last upload from z_pin has been consumed

    @property
    def runs_witness_map(self) -> bool:
        """True on a rank that needs the WHOLE witness on its device."""
        return (not self.plan.staggered) or self.rank in (self.plan.wm_rank, self.plan.wm2_rank, self.plan.wm3_rank)

    def upload_witness(self, z_host):
        """Stream-ordered upload of the witness (the same z on every rank).  z_host: a page-locked host ADDRESS (int) of m x 4
        u64 words that stays valid until the proof is done, or a numpy array (staged through an internal page-locked buffer).

        Staggered plan: the rank that runs the witness map needs all of z -- 32 m bytes over ONE PCIe link would sit at the head
        of its critical path (46 MB: 0.8 ms for S-rs256) -- so every rank uploads 1/G of z over its own link and one NCCL
        all_gather over NVLink assembles the vector on every rank (`gather_upload`, default on for G > 1; 2.07 -> 0.5 ms
        between the device-timed and the end-to-end proof at 8 GPUs).  Without it, a rank that does not run the witness map
        uploads only the slice of z its wire MSMs read.  Uniform plan: every rank runs the map and uploads all of z."""
        torch = self.torch
        if isinstance(z_host, np.ndarray):
            z = np.ascontiguousarray(z_host, dtype=np.uint64).reshape(-1, 4)
            if z.shape[0] != self.m:
                raise ffi.G16Error(ffi.ERR_BAD_ARG, f"witness has {z.shape[0]} elements, the R1CS {self.m} wires")
            if self.z_pin is None:
                self.z_pin = torch.empty((self.m, 4), dtype=torch.int64).pin_memory()
                self._z_done = torch.cuda.Event()
            else:
                self._z_done.synchronize()   # the previous proof's copy must have left the staging buffer
            self.z_pin.numpy()[:] = z.view(np.int64)
            addr = self.z_pin.data_ptr()
        else:
            addr = int(z_host)
        if not self.plan.staggered:
            self.ctx.upload_witness_async(addr, shard_only=False)
        elif self.gather_upload:
            import torch.distributed as dist
            c = self.z_chunk_len
            lo = min(self.rank * c, self.m)
            cnt = min(c, self.m - lo)
            with torch.cuda.stream(self.stream):
                if cnt:   # my 1/G of z over my own PCIe link (a raw pinned address: copy through the runtime on this stream)
                    ffi.memcpy_h2d_async(self.z_chunk.data_ptr(), addr + lo * 32, cnt * 32, self.stream.cuda_stream)
                dist.all_gather_into_tensor(self.z_all.view(-1), self.z_chunk.view(-1))   # 32 m bytes per rank over NVLink
                self.ctx.upload_witness_dev(self.z_all.data_ptr())
        else:
            self.ctx.upload_witness_async(addr, shard_only=not self.runs_witness_map)
        if isinstance(z_host, np.ndarray):
            with torch.cuda.stream(self.stream):
                self._z_done.record()

    def prove_resident(self, rr, ss, reduction=ffi.REDUCTION_LIBSNARK):
        """Witness already uploaded (upload_witness); (rr, ss) Montgomery.  Returns the raw proof on rank 0, None elsewhere."""
        import torch.distributed as dist
        ctx, plan, rank = self.ctx, self.plan, self.rank
        with self.torch.cuda.stream(self.stream):
            if rank == 0:
                ctx.prove_prepare(rr, ss)   # overlaps r*delta, s*delta, ... with the shard MSMs
            if plan.staggered:
                owner = rank == plan.wm_rank
                split = plan.wm2_rank >= 0 and reduction == ffi.REDUCTION_LIBSNARK
                ctx.prove_shard_begin_dev(rr, ss, reduction, run_witness_map=owner and not split)
                if split and owner:          # a pipeline here, b and c arrive from the helpers, then the last transform
                    n = self.vb.shape[0]
                    ctx.witness_map_part_dev(ffi.WM_PART_A)
                    dist.recv(self.vc.view(-1), src=plan.wm3_rank)   # c is ready first (one transform)
                    ctx.wm_vector_copy_dev(2, self.vc.data_ptr(), n, to_ctx=True)
                    dist.recv(self.vb.view(-1), src=plan.wm2_rank)
                    ctx.wm_vector_copy_dev(1, self.vb.data_ptr(), n, to_ctx=True)
                    ctx.witness_map_part_dev(ffi.WM_PART_FINAL)
                elif split and rank == plan.wm2_rank:
                    n = self.vb.shape[0]
                    ctx.witness_map_part_dev(ffi.WM_PART_B)
                    ctx.wm_vector_copy_dev(1, self.vb.data_ptr(), n, to_ctx=False)
                    dist.send(self.vb.view(-1), dst=plan.wm_rank)
                elif split and rank == plan.wm3_rank:
                    n = self.vc.shape[0]
                    ctx.witness_map_part_dev(ffi.WM_PART_C)
                    ctx.wm_vector_copy_dev(2, self.vc.data_ptr(), n, to_ctx=False)
                    dist.send(self.vc.view(-1), dst=plan.wm_rank)
                if owner:
                    ctx.copy_h_dev(self.h_all.data_ptr(), self.h_all.shape[0])
                scatter_h(self.h_all, self.h_mine, plan, rank)   # the one exchange step: n*32/G bytes to every peer
                if owner:
                    ctx.prove_shard_finish_dev()
                else:
                    ctx.prove_shard_finish_dev(self.h_mine.data_ptr(), plan.h_ranges[rank][0], plan.h_chunk)
            else:
                ctx.prove_shard_dev(rr, ss, reduction)
            ctx.copy_partial_dev(self.mine.data_ptr())
            dist.all_gather_into_tensor(self.gathered.view(-1), self.mine)
            if rank != 0:
                return None
            return ctx.prove_combine_dev(self.gathered.data_ptr(), self.world, rr, ss)

    def prove(self, z_mont, r: int, s: int, reduction=ffi.REDUCTION_LIBSNARK) -> Optional[Proof]:
        """z_mont: (m, 4) uint64 Montgomery witness (same on every rank).  Returns the proof on rank 0, None elsewhere."""
        rr, ss = fr_to_mont([r % R_MOD])[0], fr_to_mont([s % R_MOD])[0]
        self.upload_witness(z_mont)
        raw = self.prove_resident(rr, ss, reduction)
        return None if raw is None else Proof.from_ffi(raw)

    def close(self):
        self.stream.synchronize()
        if self._own_ctx:
            self.ctx.close()
