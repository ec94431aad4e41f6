"""One rank of tests/test_gpu_sharded.py::test_sharded_prover_class_over_nccl (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from conftest import load_golden  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402
from crescent_credentials_b200 import sharded  # noqa: E402
from crescent_credentials_b200.r1cs import load_matrices  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name in ("rand300", "dummy924_nozk"):
        meta, r1cs_bytes, pk_bytes = load_golden(name)
        mats = load_matrices(r1cs_bytes)
        pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
        z = g.fr_to_mont([int(v, 16) for v in meta["z"]])
        r, s = int(meta["r"], 16), int(meta["s"], 16)
        h_len = pk.arrays["h_query"].shape[0]
        m1 = pk.arrays["a_query"].shape[0] - 1
        plans = [None, sharded.uniform_plan(h_len, m1, world), sharded.staggered_plan(h_len, m1, world, 0.0)]
        if world >= 3:   # the b and c pipelines of the witness map on ranks 1 and 2, point-to-point to rank 0
            plans.append(sharded.staggered_plan(h_len, m1, world, wm_split=True))
        for plan in plans:
            # default constructor arguments: the class makes its own stream and context (the documented call)
            prover = sharded.ShardedProver(pk, mats, local, rank, world, plan=plan, precompute=(plan is None))
            try:
                for rep in range(3):   # back to back, no synchronisation in between on ranks != 0
                    proof = prover.prove(z, r, s)
                    if rank == 0:
                        assert proof.serialize_uncompressed().hex() == meta["proof_uncompressed"], (name, rep)
                    else:
                        assert proof is None
                # page-locked witness passed by address
                z_pin = torch.from_numpy(z.view(np.int64).copy()).pin_memory()
                prover.upload_witness(z_pin.data_ptr())
                raw = prover.prove_resident(g.fr_to_mont([r])[0], g.fr_to_mont([s])[0])
                if rank == 0:
                    assert g.Proof.from_ffi(raw).serialize_uncompressed().hex() == meta["proof_uncompressed"]
            finally:
                prover.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_WORKER_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
