"""GPU tests of the sharded path on ONE device: G contexts play the G ranks (same arithmetic as the multi-process
run; the NCCL gather itself is exercised by bench.py --gpus N and, on CPU, by the gloo test)."""
import numpy as np
import pytest

from conftest import load_golden
from crescent_credentials_b200 import ffi
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200.r1cs import load_matrices

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,world", [("rand300", 2), ("rand300", 3), ("dummy924_nozk", 4), ("rand100", 8)])
def test_sharded_prove_matches_golden(name, world):
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
    parts, ctxs = [], []
    try:
        for rank in range(world):
            ctx = ffi.Context(0)
            ctxs.append(ctx)
            ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
            ctx.load_pk(pk.arrays, pk.encoding, rank, world)
        r, s = g.fr_to_mont([int(meta["r"], 16)])[0], g.fr_to_mont([int(meta["s"], 16)])[0]
        for ctx in ctxs:
            parts.append(ctx.prove_shard(z, r, s))
        proof = g.Proof.from_ffi(ctxs[0].prove_combine(np.stack(parts), r, s))
        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        with pytest.raises(ffi.G16Error):   # a sharded context refuses the single-GPU entry point
            ctxs[0].prove(z, r, s)
    finally:
        for ctx in ctxs:
            ctx.close()


@pytest.mark.parametrize("name,world,share", [("rand300", 2, None), ("rand300", 3, 0.0), ("dummy924_nozk", 4, 0.3),
                                              ("rand100", 8, None), ("rand100_circom", 2, 1.0)])
def test_staggered_plan_matches_golden(name, world, share):
    """Staggered plan (sharded.py): only rank 0 runs the witness map, the other ranks take a larger share of the wire MSMs
    and receive their h chunk; G contexts on one device play the ranks and a device copy plays the NCCL scatter.
    share = 0.0 leaves rank 0 with EMPTY a / l / b ranges, share = 1.0 leaves every other rank with empty ones."""
    from crescent_credentials_b200.sharded import staggered_plan
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
    reduction = ffi.REDUCTION_CIRCOM if meta.get("reduction") == "circom" else ffi.REDUCTION_LIBSNARK
    h_len = np.asarray(pk.arrays["h_query"]).reshape(-1, 8).shape[0]
    m1 = np.asarray(pk.arrays["a_query"]).reshape(-1, 8).shape[0] - 1
    plan = staggered_plan(h_len, m1, world, share)
    assert plan.z_ranges[0][0] == 0 and plan.z_ranges[-1][1] == m1 and plan.h_ranges[-1][1] == h_len
    parts, ctxs = [], []
    try:
        for rank in range(world):
            ctx = ffi.Context(0)
            ctxs.append(ctx)
            ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
            ctx.load_pk(pk.arrays, pk.encoding, rank, world, h_range=plan.h_ranges[rank], z_range=plan.z_ranges[rank])
            # a rank that does not run the witness map uploads only the slice of z its wire MSMs read (stream-ordered, no sync)
            ctx.upload_witness_async(z, shard_only=(rank != 0))
        r, s = g.fr_to_mont([int(meta["r"], 16)])[0], g.fr_to_mont([int(meta["s"], 16)])[0]
        n = ctxs[0].domain_size()
        cap = max(n, world * plan.h_chunk)
        h_all = ctxs[0].dev_alloc(cap * 32)
        for rank, ctx in enumerate(ctxs):
            ctx.prove_shard_begin_dev(r, s, reduction, run_witness_map=(rank == 0))
        ctxs[0].copy_h_dev(h_all, cap)
        ctxs[0].sync()
        for rank, ctx in enumerate(ctxs):
            if rank == 0:
                ctx.prove_shard_finish_dev()
            else:   # the chunk this rank would receive from the scatter
                ctx.prove_shard_finish_dev(h_all + rank * plan.h_chunk * 32, plan.h_ranges[rank][0], plan.h_chunk)
            ptr, nbytes = ctx.partial_dev()
            ctx.sync()
            parts.append(ctx.dev_download(ptr, np.zeros(nbytes // 8, dtype=np.uint64)))
        proof = g.Proof.from_ffi(ctxs[0].prove_combine(np.stack(parts), r, s))
        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        with pytest.raises(ffi.G16Error):   # finish without begin
            ctxs[0].prove_shard_finish_dev()
        if world > 1:
            with pytest.raises(ffi.G16Error):   # the witness map needs the whole witness, this rank only holds its slice
                ctxs[1].prove_shard_begin_dev(r, s, reduction, run_witness_map=True)
        ctxs[0].prove_shard_begin_dev(r, s, reduction, run_witness_map=False)
        with pytest.raises(ffi.G16Error):   # h chunk that does not cover the rank's range
            ctxs[0].prove_shard_finish_dev(h_all, plan.h_ranges[0][1] + 1, 1)
        ctxs[0].dev_free(h_all)
    finally:
        for ctx in ctxs:
            ctx.close()


def test_prepare_then_other_prove_does_not_reuse_stale_scalars():
    """g16_prove_prepare(r, s) followed by a full prove with OTHER scalars on the same context, then combine(r, s): the
    assembly must not pick up the (r2, s2) points the full prove left in the shared slot (advisor finding, api.cu)."""
    meta, r1cs_bytes, pk_bytes = load_golden("rand300")
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
    r, s = g.fr_to_mont([int(meta["r"], 16)])[0], g.fr_to_mont([int(meta["s"], 16)])[0]
    r2, s2 = g.fr_to_mont([12345])[0], g.fr_to_mont([67890])[0]
    shard, full = ffi.Context(0), ffi.Context(0)
    try:
        for ctx in (shard, full):
            ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
            ctx.load_pk(pk.arrays, pk.encoding, 0, 1)
        part = shard.prove_shard(z, r, s)
        full.prove_prepare(r, s)
        other = g.Proof.from_ffi(full.prove(z, r2, s2))
        assert other.serialize_uncompressed().hex() != meta["proof_uncompressed"]
        proof = g.Proof.from_ffi(full.prove_combine(part.reshape(1, -1), r, s))
        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
    finally:
        shard.close()
        full.close()


def test_sharded_prover_class_over_nccl():
    """The ShardedProver class itself, one process per GPU over NCCL (2 ranks): staggered and uniform plans, numpy and
    page-locked witnesses, proofs back to back; rank 0 compares the bytes with the golden fixture.  Needs two GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    ranks = "4" if torch.cuda.device_count() >= 4 else "2"   # 4 ranks also exercise the wm_split plan
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", ranks, "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(here, "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARDED_WORKER_OK" in out.stdout


def test_witness_from_device_memory():
    """g16_upload_witness_dev (the landing step of the gathered upload of the sharded path): the witness copied from another
    device buffer gives the golden proof through the resident entry point."""
    meta, r1cs_bytes, pk_bytes = load_golden("rand300")
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
    r, s = g.fr_to_mont([int(meta["r"], 16)])[0], g.fr_to_mont([int(meta["s"], 16)])[0]
    ctx = ffi.Context(0)
    try:
        ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
        ctx.load_pk(pk.arrays, pk.encoding, 0, 1)
        d = ctx.dev_alloc(z.nbytes)
        ctx.dev_upload(d, z)
        ctx.upload_witness_dev(d)
        for _ in range(3):   # eager, captured, replayed
            proof = g.Proof.from_ffi(ctx.prove_resident(r, s))
            assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"]
        st = ctx.graph_stats()
        assert st["fallbacks"] == 0 and st["captures"] >= 1 and st["replays"] >= 2
        ctx.set_option("graph", 0)   # eager launches: same bytes
        assert g.Proof.from_ffi(ctx.prove_resident(r, s)).serialize_uncompressed().hex() == meta["proof_uncompressed"]
        ctx.dev_free(d)
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["rand300", "dummy924_nozk", "silly"])
def test_witness_map_in_parts_equals_the_whole(name):
    """g16_witness_map_part_dev: A | B | C | FINAL on one context, and the "wm_split" choreography of sharded.py on two
    contexts (B and C computed by a helper, carried over with g16_wm_vector_copy_dev, A and FINAL on the owner), both give
    the h of g16_witness_map bit for bit."""
    meta, r1cs_bytes, _ = load_golden(name)
    mats = load_matrices(r1cs_bytes)
    wires = mats.num_instance_variables + mats.num_witness_variables
    z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
    owner, helper = ffi.Context(0), ffi.Context(0)
    try:
        for ctx in (owner, helper):
            ctx.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
        n = owner.domain_size()
        want = owner.witness_map(z, ffi.REDUCTION_LIBSNARK)
        assert g.fr_from_mont(want) == [int(v, 16) for v in meta["h"]]
        out = owner.dev_alloc(n * 32)
        for rep in range(3):   # eager, captured, replayed
            owner.upload_witness(z)
            owner.witness_map_part_dev(ffi.WM_PART_A | ffi.WM_PART_B | ffi.WM_PART_C | ffi.WM_PART_FINAL)
            owner.copy_h_dev(out, n)
            owner.sync()
            assert np.array_equal(owner.dev_download(out, np.zeros((n, 4), dtype=np.uint64)), want), rep
        vb, vc = helper.dev_alloc(n * 32), helper.dev_alloc(n * 32)
        for rep in range(3):
            helper.upload_witness(z)
            owner.upload_witness(z)
            helper.witness_map_part_dev(ffi.WM_PART_B)
            helper.wm_vector_copy_dev(1, vb, n, to_ctx=False)
            helper.witness_map_part_dev(ffi.WM_PART_C)
            helper.wm_vector_copy_dev(2, vc, n, to_ctx=False)
            helper.sync()
            owner.witness_map_part_dev(ffi.WM_PART_A)
            owner.wm_vector_copy_dev(1, vb, n, to_ctx=True)
            owner.wm_vector_copy_dev(2, vc, n, to_ctx=True)
            owner.witness_map_part_dev(ffi.WM_PART_FINAL)
            owner.copy_h_dev(out, n)
            owner.sync()
            assert np.array_equal(owner.dev_download(out, np.zeros((n, 4), dtype=np.uint64)), want), rep
        with pytest.raises(ffi.G16Error):
            owner.witness_map_part_dev(0)
        fresh = ffi.Context(0)
        try:
            fresh.load_r1cs(mats.num_constraints, mats.num_instance_variables, wires, mats.row_ptr, mats.col, mats.val, mats.encoding)
            with pytest.raises(ffi.G16Error):   # no witness on the device
                fresh.witness_map_part_dev(ffi.WM_PART_A)
        finally:
            fresh.close()
    finally:
        owner.close()
        helper.close()
