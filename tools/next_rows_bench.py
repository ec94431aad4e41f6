"""Scope rows f-1 / f-2 / f-3 at full size, through the reference's FILE formats, plus the whole drop-in flow end to end:

  1. build the S-rs256 instance and mint its proving key on the GPU (f-2: generator, timed)
  2. prove in memory (Python host) -> proof P1
  3. write the instance as an iden3 .r1cs file and the key as arkworks' uncompressed ProvingKey bytes (g16_cli r1cs-write /
     pk-write: ~0.6 GB each), the witness as canonical 32-byte words
  4. f-1: g16_cli r1cs-bench (R1CSFile::read + to_matrices, once per circuit)
  5. f-3: deserialize_uncompressed_unchecked + g16_ctx_load_pk (canonical words converted and window tables built on the GPU)
  6. the compiled host end to end: g16_cli prove --r1cs --pk --witness -> proof P2;  P2 must equal P1 byte for byte
  7. f-4 at full size: the device verifier accepts P1 under the key read back from the bytes and rejects a changed input
One JSON line on stdout."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crescent_credentials_b200 import ffi, generator, synth  # noqa: E402
from crescent_credentials_b200 import groth16 as g  # noqa: E402
from crescent_credentials_b200 import verifier as v  # noqa: E402

CLI = os.path.join(ROOT, "crescent_credentials_b200", "host", "cpp", "g16_cli")


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "S-rs256"
    out = {"what": "next_rows_full_size", "workload": workload}
    tmp = tempfile.mkdtemp(prefix="g16_next_")
    ctx = ffi.Context(0)
    t0 = time.time()
    inst = synth.make_instance(ctx, workload)
    out["instance_build_s"] = round(time.time() - t0, 2)
    td = generator.Trapdoor(alpha=0x1234567 + 11, beta=0x89ABCDE + 13, gamma=1, delta=0xFEDCBA9 + 17, t=0x5EED0000C0FFEE0000BEEF + 19)
    t0 = time.time()
    pk, _ = generator.generate_parameters_with_qap(ctx, inst.matrices, td)
    ctx.sync()
    out["f2_generator_s"] = round(time.time() - t0, 2)
    out["f2_points"] = {k: int(np.asarray(pk.arrays[k]).reshape(-1, 16 if k == "b_g2_query" else 8).shape[0])
                        for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")}
    r, s = 0x1111222233334444555566667777888899990000AAAABBBBCCCCDDDD % g.R_MOD, 0x0F0E0D0C0B0A09080706050403020100FFEEDDCCBBAA9988 % g.R_MOD
    prover = g.Groth16(0, precompute=True)
    t0 = time.time()
    p1 = prover.create_proof_with_reduction_and_matrices(pk, r, s, inst.matrices, inst.ni, inst.nc, inst.z_mont)
    out["python_first_prove_s"] = round(time.time() - t0, 2)
    prover.close()

    # ---- the files -------------------------------------------------------------------------------------------------------
    t0 = time.time()
    mats = inst.matrices
    for k in range(3):
        np.ascontiguousarray(mats.row_ptr[k], dtype="<u8").tofile(os.path.join(tmp, f"m.{k}.ptr"))
        np.ascontiguousarray(mats.col[k], dtype="<u4").tofile(os.path.join(tmp, f"m.{k}.col"))
        np.ascontiguousarray(mats.val[k], dtype="<u8").tofile(os.path.join(tmp, f"m.{k}.val"))
    canon = lambda a: ctx.field_op(ffi.FIELD_FQ, ffi.OP_FROM_MONT, np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4))
    dumps = dict(alpha_g1=pk.arrays["alpha_g1"], beta_g2=pk.arrays["beta_g2"], gamma_g2=pk.gamma_g2, vk_delta_g1=pk.arrays["delta_g1"],
                 delta_g2=pk.arrays["delta_g2"], gamma_abc_g1=pk.gamma_abc_g1, beta_g1=pk.arrays["beta_g1"], delta_g1=pk.arrays["delta_g1"],
                 a_query=pk.arrays["a_query"], b_g1_query=pk.arrays["b_g1_query"], b_g2_query=pk.arrays["b_g2_query"],
                 h_query=pk.arrays["h_query"], l_query=pk.arrays["l_query"])
    for k, a in dumps.items():
        canon(a).tofile(os.path.join(tmp, f"pk.{k}.bin"))
    ctx.field_op(ffi.FIELD_FR, ffi.OP_FROM_MONT, inst.z_mont).tofile(os.path.join(tmp, "z.bin"))
    run = lambda *a: subprocess.run([CLI, *map(str, a)], capture_output=True, text=True, timeout=900, check=True)
    r1cs_path, pk_path = os.path.join(tmp, "main_c.r1cs"), os.path.join(tmp, "prover_params.bin")
    out["r1cs_file_bytes"] = int(run("r1cs-write", "--prefix", os.path.join(tmp, "m"), "--nc", inst.nc, "--nwires", inst.m, "--ninputs", inst.ni,
                                     "--out", r1cs_path).stdout)
    out["pk_file_bytes"] = int(run("pk-write", "--prefix", os.path.join(tmp, "pk"), "--out", pk_path).stdout)
    out["write_files_s"] = round(time.time() - t0, 2)
    ctx.close()

    # ---- f-1 ----------------------------------------------------------------------------------------------------------------
    out["f1_r1cs_to_csr"] = json.loads(run("r1cs-bench", r1cs_path).stdout)

    # ---- f-3 (Python host: numpy reader + device-side conversion) ---------------------------------------------------------------
    t0 = time.time()
    pk_bytes = open(pk_path, "rb").read()
    t1 = time.time()
    pk2 = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    t2 = time.time()
    c2 = ffi.Context(0)
    c2.load_pk(pk2.arrays, pk2.encoding, 0, 1, True)
    c2.sync()
    t3 = time.time()
    c2.close()
    out["f3_pk_load"] = {"read_file_s": round(t1 - t0, 2), "deserialize_unchecked_s": round(t2 - t1, 2),
                         "upload_convert_tables_s": round(t3 - t2, 2), "bytes": len(pk_bytes)}
    del pk_bytes

    # ---- the compiled host, files in, proof out ------------------------------------------------------------------------------
    t0 = time.time()
    p = run("prove", "--r1cs", r1cs_path, "--pk", pk_path, "--witness", os.path.join(tmp, "z.bin"), "--r", hex(r), "--s", hex(s),
            "--out", os.path.join(tmp, "proof.bin"), "--repeat", 3)
    out["cli_prove_wall_s"] = round(time.time() - t0, 2)
    for line in p.stderr.splitlines():
        if line.startswith("g16_cli timings:"):
            out["cli_timings_ms"] = json.loads(line.split(":", 1)[1])
    out["cli_proof_equals_python_proof"] = p.stdout.strip() == p1.serialize_compressed().hex()
    p_pg = run("prove", "--r1cs", r1cs_path, "--pk", pk_path, "--witness", os.path.join(tmp, "z.bin"), "--r", hex(r), "--s", hex(s),
               "--out", os.path.join(tmp, "proof_pg.bin"), "--repeat", 3, "--pageable")
    for line in p_pg.stderr.splitlines():
        if line.startswith("g16_cli timings:"):
            out["cli_timings_pageable_ms"] = json.loads(line.split(":", 1)[1])
    out["cli_pageable_proof_equals"] = p_pg.stdout.strip() == p.stdout.strip()
    out["cli_proof_file_equals"] = open(os.path.join(tmp, "proof.bin"), "rb").read() == p1.serialize_uncompressed()

    # ---- f-4 at full size -------------------------------------------------------------------------------------------------------
    ver = v.Verifier(0)
    t0 = time.time()
    pvk = ver.prepare_verifying_key(pk2)
    public = g.fr_from_mont(inst.z_mont[1:inst.ni])
    good = ver.verify_proof(pvk, p1, public)
    bad = ver.verify_proof(pvk, p1, [(public[0] + 1) % g.R_MOD] + public[1:])
    out["f4_verify"] = {"accepts_proof": bool(good), "rejects_wrong_input": not bad, "public_inputs": len(public), "s": round(time.time() - t0, 2)}
    ver.close()
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    print(json.dumps(out), flush=True)
    ok = out["cli_proof_equals_python_proof"] and out["cli_proof_file_equals"] and good and not bad
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
