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
