"""CPU tests of host-side logic: R1CS -> CSR flattening, proving-key (de)serialisation, synthetic generator determinism,
MSM shard partitioning, and the 2-rank (gloo) gather used by the sharded prover."""
import os
import sys

import numpy as np
import pytest

import coracle as c
import pyref as o
from conftest import load_golden
from crescent_credentials_b200 import ffi, generator, synth
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200.r1cs import load_matrices
from crescent_credentials_b200.sharded import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_csr_flattening_sums_duplicates_and_drops_zeros():
    cons = [([(1, 5), (1, 7), (2, 3)], [(0, 1)], [(3, o.R_MOD - 4), (3, 4)]),   # dup wire 1 -> 12; wire 3 cancels to 0
            ([], [(2, 2)], [])]
    data = o.write_r1cs(4, 1, 0, 2, cons)
    m = load_matrices(data)
    assert m.num_instance_variables == 2 and m.num_witness_variables == 2 and m.num_constraints == 2
    assert list(m.row_ptr[0]) == [0, 2, 2] and list(m.col[0]) == [1, 2]
    assert g.limbs_to_ints(m.val[0]) == [12, 3]
    assert list(m.row_ptr[2]) == [0, 0, 0]
    om = o.r1cs_to_matrices(o.read_r1cs(data))
    assert om.a == [[(12, 1), (3, 2)], []] and om.c == [[], []]


def test_pk_deserialisation_round_trip():
    meta, _, pk_bytes = load_golden("rand100")
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    assert pk.encoding == ffi.ENC_CANONICAL
    n = meta["domain_size"]
    assert pk.arrays["h_query"].shape == (n - 1, 8)           # h_query has n-1 points (generator.rs:174-179)
    assert pk.arrays["a_query"].shape[0] == meta["num_instance"] + meta["num_witness"]
    assert pk.arrays["l_query"].shape[0] == meta["num_witness"]
    # canonical coordinates lie on the curve; infinities are zeros
    for row in pk.arrays["b_g1_query"]:
        x, y = g.limbs_to_ints(row.reshape(2, 4))
        assert (x, y) == (0, 0) or o.G1.is_on_curve((x, y))
    with pytest.raises(ValueError):
        g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes + b"\0")


def test_synthetic_streams_are_deterministic_and_match_scalar_splitmix():
    a = synth.stream(0xC0FFEE, 0x10, 1000)
    assert np.array_equal(a, synth.stream(0xC0FFEE, 0x10, 1000))
    base = (0xC0FFEE ^ (0x10 << 40)) & 0xFFFFFFFFFFFFFFFF
    assert int(a[17]) == o.splitmix64(17 ^ base)
    z = synth.witness_canonical(1, 5000, "circom")
    ints = g.limbs_to_ints(z)
    assert ints[0] == 1 and all(v < o.R_MOD for v in ints)
    frac_bits = sum(v in (0, 1) for v in ints) / len(ints)
    assert 0.8 < frac_bits < 0.9
    rp, col, val = synth._matrix(3, 0x10, 2000, 1900, 4.7, 1, 0)
    assert rp[-1] == len(col) == len(val) and col.max() < 1900
    assert 3.5 < float(rp[-1]) / 2000 < 6.0
    assert all(v < o.R_MOD for v in g.limbs_to_ints(val[:500]))


def test_transpose_csr():
    rp = np.array([0, 2, 3, 3], dtype=np.uint64)
    col = np.array([2, 0, 2], dtype=np.uint32)
    val = np.arange(12, dtype=np.uint64).reshape(3, 4)
    tp, tc, tv = generator.transpose_csr(3, 4, rp, col, val)
    assert list(tp) == [0, 1, 1, 3, 3] and list(tc) == [0, 0, 1]
    assert np.array_equal(tv, val[[1, 0, 2]])


def test_shard_ranges_partition_every_query():
    for total in (0, 1, 7, 1_449_999, 2_097_151):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from crescent_credentials_b200.sharded import gather_partials
    # each rank contributes a recognisable 96-word partial; rank 0 must see them in rank order
    mine = torch.full((ffi.PARTIAL_U64,), rank + 1, dtype=torch.int64)
    allp = gather_partials(mine, world)
    if rank == 0:
        q.put([int(allp[r, 0]) for r in range(world)] + [tuple(allp.shape)])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_of_partials():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [1, 2, (2, ffi.PARTIAL_U64)]


def test_staggered_plan_partitions_every_query():
    """sharded.staggered_plan: rank 0 (the witness-map rank) gets the share f0 of the wire MSMs, the rest is split evenly;
    h is cut in equal scatter chunks.  Whatever f0, the ranges must partition the queries."""
    from crescent_credentials_b200.sharded import rank0_wire_share, staggered_plan
    for h_len, m1 in ((0, 1), (7, 5), (1023, 923), (2_097_151, 1_449_999)):
        for world in (1, 2, 3, 4, 8):
            for share in (None, 0.0, 0.25, 1.0):
                plan = staggered_plan(h_len, m1, world, share)
                for spans, total in ((plan.z_ranges, m1), (plan.h_ranges, h_len)):
                    assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == total
                    assert all(lo <= hi for lo, hi in spans)
                    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
                if world > 1:
                    assert plan.wm_rank == 0 and plan.h_chunk * world >= h_len
                    assert all(hi - lo <= plan.h_chunk and lo == min(r * plan.h_chunk, h_len) for r, (lo, hi) in enumerate(plan.h_ranges))
                    if share is not None:
                        assert plan.z_ranges[0][1] == round(m1 * share)
    # the balance point: wm + f0 Z == (1 - f0) Z / (G - 1); no wire work for rank 0 once Z / (G - 1) < wm
    assert rank0_wire_share(1) == 1.0
    f2 = rank0_wire_share(2, 0.25)
    assert abs((0.25 + f2) - (1 - f2)) < 1e-12
    assert rank0_wire_share(8, 0.25) == 0.0


def _gloo_scatter_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from crescent_credentials_b200.sharded import scatter_h, staggered_plan
    h_len = 13
    plan = staggered_plan(h_len, 9, world)
    h_all = None
    if rank == plan.wm_rank:   # n = 16 coefficients of 4 words; coefficient i is [4i, 4i+1, 4i+2, 4i+3]
        h_all = torch.arange(max(16, world * plan.h_chunk) * 4, dtype=torch.int64).view(-1, 4)
    h_mine = torch.zeros((plan.h_chunk, 4), dtype=torch.int64)
    scatter_h(h_all, h_mine, plan, rank)
    lo, hi = plan.h_ranges[rank]
    q.put((rank, lo, hi, [int(v) for v in h_mine[:hi - lo, 0]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_scatter_of_h_chunks():
    """The one extra collective of the staggered plan: every rank receives exactly the h coefficients of its h range."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_scatter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [(lo, hi) for _, lo, hi, _ in got] == [(0, 7), (7, 13)]
    for _, lo, hi, first_words in got:
        assert first_words == [4 * i for i in range(lo, hi)]


def test_sharded_msm_partials_add_up_on_cpu():
    """The sharding identity the multi-GPU path relies on: MSM over [0,N) == sum of MSMs over the rank ranges
    (checked with the CPU oracle; the GPU version of this test is in test_gpu_sharded.py)."""
    ks = [o.stream_fr(0x51, i) for i in range(97)]
    pts = c.fixed_base(1, g.fr_to_mont(ks))
    sc_int = [o.stream_fr(0x52, i) for i in range(97)]
    sc = g.fr_to_mont(sc_int)
    full = g.g1_from_mont(c.msm(1, pts, sc))
    acc = None
    for r in range(4):
        lo, hi = shard_range(97, r, 4)
        acc = o.G1.add(acc, g.g1_from_mont(c.msm(1, pts[lo:hi], sc[lo:hi])))
    assert acc == full


# ---- rand 0.8 StdRng mirror (crescent_credentials_b200/rng.py) ---------------------------------------------------------
def test_chacha_block_rfc8439_vector():
    """RFC 8439 section 2.3.2 (20 rounds) pins the quarter round and the column/diagonal schedule."""
    import struct
    from crescent_credentials_b200 import rng
    key = list(struct.unpack("<8I", bytes(range(32))))
    out = rng.chacha_block(list(rng._SIGMA) + key + [1, 0x09000000, 0x4A000000, 0], 20)
    assert out == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                   0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_chacha12_zero_key_keystream():
    """ChaCha12, 256-bit zero key, zero IV, block 0 (eSTREAM test vector TC1): the 12-round StdRng core."""
    import struct
    from crescent_credentials_b200 import rng
    r = rng.StdRng(bytes(32))
    ks = b"".join(struct.pack("<I", r.next_u32()) for _ in range(16))
    assert ks.hex() == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                        "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")


def test_stdrng_word_order_and_buffer_straddle():
    from crescent_credentials_b200 import rng
    a, b = rng.test_rng(), rng.test_rng()
    words = [a.next_u32() for _ in range(140)]
    # next_u64 = (low word first) pairs of the same stream
    assert [b.next_u64() for _ in range(3)] == [words[2 * i] | (words[2 * i + 1] << 32) for i in range(3)]
    # BlockRng rule at the end of the 64-word buffer: the last word is the low half, the next buffer's first the high half
    c = rng.test_rng()
    for _ in range(63):
        c.next_u32()
    assert c.next_u64() == words[63] | (words[64] << 32)
    assert c.next_u32() == words[65]
    # the 64-bit block counter keeps running across buffers (words 64.. come from blocks 4..7)
    d = rng.test_rng()
    for _ in range(64):
        d.next_u32()
    assert d.counter == 4 and d.next_u64() == words[64] | (words[65] << 32)
    assert rng.StdRng.seed_from_u64(42).next_u64() == rng.StdRng.seed_from_u64(42).next_u64()
    assert rng.StdRng.seed_from_u64(42).next_u64() != rng.StdRng.seed_from_u64(43).next_u64()


def test_sample_fr_with_stdrng_follows_fr_rand():
    """Fr::rand (ark-ff 0.4): 4 x next_u64 limb 0 first, top limb masked to 254 bits, reject >= r, accepted integer is the
    Montgomery representation."""
    from crescent_credentials_b200 import rng
    a, b = rng.test_rng(), rng.test_rng()
    got = [g.sample_fr(a) for _ in range(6)]
    want = []
    while len(want) < 6:
        v = sum(b.next_u64() << (64 * k) for k in range(4)) & ((1 << 254) - 1)
        if v < o.R_MOD:
            want.append(v * pow(1 << 256, -1, o.R_MOD) % o.R_MOD)
    assert got == want and len(set(got)) == 6


def test_cpu_sample_scaling_follows_arkworks_window_rule():
    """bench.py scales the CPU sample by ark-ec's addition count: window rule c = bit_length(n)*69//100 + 2 (3 below 32
    pairs), ceil(254/c) windows of n + 2^c additions, one rayon task per window."""
    import bench
    assert bench.ark_msm_additions(0) == 0
    assert bench.ark_msm_additions(16) == 85 * (16 + 2 * 4)                  # c = 3, 85 windows
    n = 1 << 15                                                              # bit_length 16 -> c = 13, 20 windows
    assert bench.ark_msm_additions(n) == 20 * (n + 2 * 4096)
    full = 1_450_000                                                         # bit_length 21 -> c = 16, 16 windows
    assert bench.ark_msm_additions(full) == 16 * (full + 2 * 32768)
    assert bench.ark_msm_additions(full, threads=16) == full + 2 * 32768     # one round of windows on 16 threads
    # a 12.5 % sample pays more additions per point than the full MSM: the scale factor stays below len/sample
    k = full // 8
    assert 1.0 < bench.cpu_scale(full, k, 16) < full / k


def test_vectorised_r1cs_loader_equals_the_rowwise_definition():
    """Ragged rows, empty rows, repeated wires (summing to non-zero and to zero), explicit zero coefficients, unsorted wires,
    an unreduced coefficient: the array-based loader must produce exactly the CSR of the term-by-term one."""
    from crescent_credentials_b200.r1cs import _load_matrices_rowwise
    import random
    rnd = random.Random(4242)
    nw, nc = 37, 60
    cons = []
    for i in range(nc):
        row = []
        for k in range(3):
            n = rnd.choice([0, 0, 1, 2, 3, 5, 9])
            lc = [(rnd.randrange(nw), rnd.choice([1, o.R_MOD - 1, rnd.randrange(o.R_MOD), 0 if rnd.random() < 0.1 else 7])) for _ in range(n)]
            if lc and rnd.random() < 0.2:
                w, v = lc[0]
                lc.append((w, (o.R_MOD - v) % o.R_MOD if rnd.random() < 0.5 else 3))  # repeat a wire: cancels or sums
            rnd.shuffle(lc)
            row.append(lc)
        cons.append(tuple(row))
    data = bytearray(o.write_r1cs(nw, 2, 1, nw - 4, cons))
    a, b = load_matrices(bytes(data)), _load_matrices_rowwise(bytes(data))
    for k in range(3):
        assert np.array_equal(a.row_ptr[k], b.row_ptr[k]) and np.array_equal(a.col[k], b.col[k]) and np.array_equal(a.val[k], b.val[k])
    assert (a.num_instance_variables, a.num_witness_variables, a.num_constraints) == (b.num_instance_variables, b.num_witness_variables, nc)
    for name in ("rand300", "dummy924_nozk", "silly"):
        _, r1cs_bytes, _ = load_golden(name)
        a, b = load_matrices(r1cs_bytes), _load_matrices_rowwise(r1cs_bytes)
        for k in range(3):
            assert np.array_equal(a.row_ptr[k], b.row_ptr[k]) and np.array_equal(a.col[k], b.col[k]) and np.array_equal(a.val[k], b.val[k])


def test_automatic_window_rule():
    """g16_msm_window_bits: with window tables round(log2 n) - 3 from 2^17 points on -- the sizes measured on the GPU
    (profiles/r02_ab_shard_window*.log, r02_ab_window_n1.log: one rank of 8 / 4 / 2 and the whole S-rs256 / S-mdl1 proof) --
    log2 n - 1 below; without tables log2 n - 5 capped at 16; never below 4.  The table count follows from it."""
    import math
    measured_best = {207_126: 15, 262_144: 15, 414_252: 16, 524_288: 16, 828_504: 17, 1_048_576: 17, 1_449_999: 17,
                     2_097_151: 18, 2_899_999: 18, 4_194_303: 19}
    for n, c in measured_best.items():
        assert ffi.msm_window_bits(n, True) == c, n
    for n in (1 << 17, 150_000, (1 << 26) + 5, 1 << 30):
        assert ffi.msm_window_bits(n, True) == min(20, round(math.log2(n)) - 3)
    for lg in range(6, 17):
        assert ffi.msm_window_bits(1 << lg, True) == max(4, lg - 1)
        assert ffi.msm_window_bits((1 << lg) + 1, False) == max(4, min(16, lg - 5))
    assert ffi.msm_window_bits(1, True) == 4 and ffi.msm_window_bits(0, False) == 4
    assert ffi.msm_window_bits(1 << 26, False) == 16
    # monotone in n
    prev = 0
    for n in range(1 << 12, 1 << 23, 37_123):
        c = ffi.msm_window_bits(n, True)
        assert c >= prev or (n >= (1 << 17) and prev - c <= 1 and n < (1 << 17) + 37_123 * 2)
        prev = c


class _StubVerifier:
    """Stands in for verifier.Verifier on a CPU box: accepts a 'proof' iff it is an even number; records what it was given."""

    def __init__(self):
        self.seen = []

    def verify_proofs(self, pvk, proofs, public_inputs):
        assert len(proofs) == len(public_inputs) and all(x == [p, p + 1] for p, x in zip(proofs, public_inputs))
        self.seen += list(proofs)
        return [p % 2 == 0 for p in proofs]


def _gloo_verify_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from crescent_credentials_b200.verifier import verify_proofs_replicated
    got = {}
    for n in (0, 1, 2, 5, 16, 17):
        v = _StubVerifier()
        proofs = [7 * i + 3 for i in range(n)]
        out = verify_proofs_replicated(v, None, proofs, [[p, p + 1] for p in proofs], rank, world)
        got[n] = (out, v.seen)
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_replicated_verification_partitions_by_proof(world):
    """verifier.verify_proofs_replicated over gloo: every rank verifies only its slice, every rank ends with the full verdict
    list in proof order (uneven and empty slices included)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_gloo_verify_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for n in (0, 1, 2, 5, 16, 17):
        proofs = [7 * i + 3 for i in range(n)]
        seen_all = []
        for r in range(world):
            out, seen = res[r][n]
            assert out == [p % 2 == 0 for p in proofs], (n, r)
            lo, hi = shard_range(n, r, world)
            assert seen == proofs[lo:hi]
            seen_all += seen
        assert seen_all == proofs


def test_bench_traffic_records_match_by_geometry():
    """bench.ncu_traffic: `roofline.traffic` comes from the committed ncu capture whose slot count is this run's (one record
    per window geometry); any other geometry reads as no capture (null in the bench line), never a scaled guess."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    recs = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["S-rs256/uniform"]
    assert isinstance(recs, list) and len(recs) >= 2
    for rec in recs:
        got, src = bench.ncu_traffic("S-rs256", "uniform", rec["slots"])
        assert got == rec["dram_read_bytes"] + rec["dram_write_bytes"] and src == rec["source"]
        assert os.path.exists(os.path.join(ROOT, src.split(":")[0])), src     # the capture summary it cites is committed
        assert 232 * rec["slots"] < got < 2 * 232 * rec["slots"]                # above the algorithmic bytes, below 2x
    assert bench.ncu_traffic("S-rs256", "uniform", recs[0]["slots"] * 2) == (None, None)
    assert bench.ncu_traffic("S-mdl1", "uniform", recs[0]["slots"]) == (None, None)


def test_bench_whole_prove_budget_is_the_surveys():
    """bench.survey_prove_budget / whole_prove_roofline: SURVEY 8d's per-proof budget for S-rs256 -- (2.1 + 4.35) M G1 points x 160
    + 1.45 M G2 points x 480 = 1.73 G Fq products, 0.17 G Fr products, >= ~27 ms at 70 G mul/s -- and the fraction arithmetic."""
    sys.path.insert(0, ROOT)
    import bench
    b = bench.survey_prove_budget(1 << 21, 1_450_000, 24, 16_047_951)
    assert abs(b["fq_mul"] - 1.73e9) < 0.01e9 and abs(b["fr_mul"] - 0.177e9) < 0.01e9
    assert 26.5 < b["mul"] / 70e9 * 1e3 < 27.5
    cfg = {"domain": 1 << 21, "wires": 1_450_000, "nnz": 16_047_951}
    w = bench.whole_prove_roofline(cfg, 24, 26.0, 65.0, 1)
    assert abs(w["algorithmic_gmul_per_s"] - b["mul"] / 26.0e-3 / 1e9) < 1e-6
    assert abs(w["frac_of_mul_peak"] - w["algorithmic_gmul_per_s"] / 65.0) < 1e-12 and 1.0 < w["frac_of_mul_peak"] < 1.2
    assert abs(bench.whole_prove_roofline(cfg, 24, 26.0, 65.0, 8)["frac_of_mul_peak"] * 8 - w["frac_of_mul_peak"]) < 1e-12
    assert bench.whole_prove_roofline(cfg, 24, 26.0, None, 1)["frac_of_mul_peak"] is None
