"""The compiled (C++) host side -- crescent_credentials_b200/host/cpp/ark_groth16_b200.hpp driven through g16_cli --
against the Python mirror, the oracle and the golden fixtures.  CPU part: marshalling, file readers, RNG mirror, error
behaviour without a device.  GPU part (-m gpu): reference-shaped prove calls whose proof bytes must equal the goldens."""
import os
import random
import subprocess

import pytest

import pyref as o
from conftest import GOLDEN, GOLDEN_NAMES, ROOT, load_golden
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200 import rng as rust_rng
from crescent_credentials_b200.r1cs import load_matrices
from test_oracle_golden import R1CS_SAMPLE_HEX

CPP_DIR = os.path.join(ROOT, "crescent_credentials_b200", "host", "cpp")
CLI = os.path.join(CPP_DIR, "g16_cli")


@pytest.fixture(scope="module")
def cli():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "crescent_credentials_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", CPP_DIR])

    def run(*args, check=True):
        p = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=600)
        if check and p.returncode != 0:
            raise AssertionError(f"g16_cli {' '.join(map(str, args))} -> {p.returncode}\n{p.stderr}")
        return p
    return run


@pytest.mark.parametrize("spec,make", [("test", rust_rng.test_rng), ("seed:0", lambda: rust_rng.StdRng.seed_from_u64(0)),
                                       ("seed:42", lambda: rust_rng.StdRng.seed_from_u64(42)),
                                       ("seed:18446744073709551615", lambda: rust_rng.StdRng.seed_from_u64((1 << 64) - 1))])
def test_stdrng_mirror_matches_python_mirror(cli, spec, make):
    """200 next_u64 draws cross the 64-word buffer boundary three times."""
    rng = make()
    want = [f"{rng.next_u64():016x}" for _ in range(200)]
    assert cli("rng", spec, 200).stdout.split() == want
    rng = make()
    want = [f"{g.sample_fr(rng):064x}" for _ in range(40)]
    assert cli("rand-fr", spec, 40).stdout.split() == want


def test_host_montgomery_marshalling(cli):
    rnd = random.Random(9)
    for name, p in (("fr", o.R_MOD), ("fq", o.Q_MOD)):
        rinv = pow(1 << 256, -1, p)
        for v in [0, 1, 2, p - 1, (p - 1) // 2, (p + 1) // 2, (1 << 253) % p] + [rnd.randrange(p) for _ in range(20)]:
            assert int(cli("fp", name, "from", hex(v)).stdout, 16) == (v << 256) % p
            assert int(cli("fp", name, "into", hex(v)).stdout, 16) == v * rinv % p
        assert cli("fp", name, "from", hex(p), check=False).returncode != 0  # from_bigint rejects unreduced input


def test_reference_byte_goldens_through_the_cpp_host(cli):
    """Montgomery little-endian bytes of Fq::one (forks/circom-compat/src/zkey.rs:397-402): the in-memory form of 1."""
    one = int(cli("fp", "fq", "from", "0x1").stdout, 16)
    assert one.to_bytes(32, "little") == bytes([157, 13, 143, 197, 141, 67, 93, 211, 61, 11, 199, 245, 40, 235, 120, 10, 44, 70, 121, 120, 111,
                                                163, 110, 102, 47, 223, 7, 154, 193, 119, 10, 14])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_r1cs_to_matrices_matches_python_loader(cli, name):
    path = os.path.join(GOLDEN, name + ".r1cs")
    mats = load_matrices(open(path, "rb").read())
    out = cli("r1cs", path).stdout.splitlines()
    head = out[0].split()
    assert [int(head[i]) for i in (1, 3, 5, 7, 8, 9)] == [mats.num_instance_variables, mats.num_witness_variables, mats.num_constraints,
                                                          mats.a_num_non_zero, mats.b_num_non_zero, mats.c_num_non_zero]
    want = []
    for k in range(3):
        vals = g.limbs_to_ints(mats.val[k])
        for i in range(mats.num_constraints):
            for e in range(int(mats.row_ptr[k][i]), int(mats.row_ptr[k][i + 1])):
                want.append(f"{k} {i} {int(mats.col[k][e])} {vals[e]:064x}")
    assert out[1:] == want


def test_r1cs_reference_sample_and_duplicate_wires(cli, tmp_path):
    """The worked example of r1cs_reader.rs:266-344 (coefficients asserted at :320-344), then a row with a repeated wire
    and a cancelling pair: summed / dropped as ark-relations' to_matrices does after inlining."""
    p = tmp_path / "sample.r1cs"
    p.write_bytes(bytes.fromhex("".join(R1CS_SAMPLE_HEX.split())))
    out = cli("r1cs", p).stdout.splitlines()
    assert out[0].startswith("num_instance 4 num_witness 3 num_constraints 3")
    assert f"0 0 5 {3:064x}" in out and f"1 2 0 {6:064x}" in out
    assert not [ln for ln in out[1:] if ln.startswith("2 1 ")]  # constraint 1 has an empty C
    cons = [([(1, 5), (2, 7), (1, o.R_MOD - 2)], [(0, 1), (3, 4), (3, o.R_MOD - 4)], [(2, 9)])]
    p.write_bytes(o.write_r1cs(4, 1, 0, 2, cons))
    out = cli("r1cs", p).stdout.splitlines()
    assert out[1:] == [f"0 0 1 {3:064x}", f"0 0 2 {7:064x}", f"1 0 0 {1:064x}", f"2 0 2 {9:064x}"]


def test_r1cs_reader_error_conventions(cli, tmp_path):
    data = bytes.fromhex("".join(R1CS_SAMPLE_HEX.split()))
    for pos, val, msg in ((0, 0x00, "Invalid magic number"), (4, 0x02, "Unsupported version"), (28, 0x02, "only supports bn256")):
        d = bytearray(data)
        d[pos] = val
        p = tmp_path / "bad.r1cs"
        p.write_bytes(bytes(d))
        r = cli("r1cs", p, check=False)
        assert r.returncode == 2 and msg in r.stderr
    p = tmp_path / "short.r1cs"
    p.write_bytes(data[:-9])
    assert cli("r1cs", p, check=False).returncode == 2


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_pk_bytes_round_trip(cli, tmp_path, name):
    """deserialize_uncompressed_unchecked . serialize_uncompressed == identity on arkworks-layout ProvingKey bytes, and the
    vector lengths are the ones the Python reader sees."""
    src = os.path.join(GOLDEN, name + ".pk.bin")
    dst = tmp_path / "rt.bin"
    out = cli("pk-roundtrip", src, dst).stdout.split()
    assert dst.read_bytes() == open(src, "rb").read()
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(open(src, "rb").read())
    got = dict(zip(out[0::2], map(int, out[1::2])))
    assert got == {"a": len(pk.arrays["a_query"]), "b_g1": len(pk.arrays["b_g1_query"]), "b_g2": len(pk.arrays["b_g2_query"]),
                   "h": len(pk.arrays["h_query"]), "l": len(pk.arrays["l_query"]), "gamma_abc": len(pk.raw_vk["gamma_abc_g1"])}
    bad = tmp_path / "trunc.bin"
    bad.write_bytes(open(src, "rb").read()[:-1])
    assert cli("pk-roundtrip", bad, dst, check=False).returncode == 2


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_proof_serialisation_matches_goldens(cli, name):
    """Proof{a, b, c} -> ark-serialize bytes (flag bits included) from the canonical coordinates of the golden proof."""
    meta, _, _ = load_golden(name)
    raw = bytes.fromhex(meta["proof_uncompressed"])
    le = lambda b: int.from_bytes(b, "little")
    words = [le(raw[i:i + 32]) for i in range(0, 256, 32)]
    # uncompressed layout: flags live in the top bits of the last coordinate of each point (indices 1, 5, 7)
    for i in (1, 5, 7):
        words[i] &= (1 << 254) - 1
    out = cli("proof-ser", *[hex(w) for w in words]).stdout.split()
    assert out == [meta["proof_compressed"], meta["proof_uncompressed"]]


def test_proof_serialisation_infinity_and_sign_flags(cli):
    y_small, y_big = 5, o.Q_MOD - 5
    for y, flag in ((y_small, 0), (y_big, 0x80)):
        out = cli("proof-ser", "0x7", hex(y), "inf", "-", "-", "-", "0x7", hex(y)).stdout.split()
        comp = bytes.fromhex(out[0])
        assert comp[31] & 0xC0 == flag and comp[32:96] == bytes(63) + b"\x40" and comp[127] & 0xC0 == flag
        assert g.Proof((7, y), None, (7, y)).serialize_compressed() == comp
        assert g.Proof((7, y), None, (7, y)).serialize_uncompressed().hex() == out[1]
    # Fq2 ordering: c1 decides, c0 breaks ties
    for y0, y1 in ((1, 0), (o.Q_MOD - 1, 0), (3, o.Q_MOD - 2), (o.Q_MOD - 3, 2)):
        out = cli("proof-ser", "inf", "-", "0x1", "0x2", hex(y0), hex(y1), "inf", "-").stdout.split()
        assert out[0] == g.Proof(None, ((1, 2), (y0, y1)), None).serialize_compressed().hex()


def test_no_cpu_fallback_in_the_cpp_host(cli, tmp_path):
    from crescent_credentials_b200 import ffi
    if ffi.load_library().g16_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    meta, _, _ = load_golden("silly")
    w = tmp_path / "w.bin"
    w.write_bytes(b"".join(int(v, 16).to_bytes(32, "little") for v in meta["z"]))
    r = cli("prove", "--r1cs", os.path.join(GOLDEN, "silly.r1cs"), "--pk", os.path.join(GOLDEN, "silly.pk.bin"), "--witness", w,
            "--r", meta["r"], "--s", meta["s"], "--out", tmp_path / "p.bin", check=False)
    assert r.returncode == 100 + ffi.ERR_NO_DEVICE and "no CPU fallback" in r.stderr
    assert not (tmp_path / "p.bin").exists()  # never a proof on error


# ---- GPU: reference-shaped calls through the compiled host ------------------------------------------------------------------------
def _witness_file(tmp_path, meta):
    w = tmp_path / "witness.bin"
    w.write_bytes(b"".join(int(v, 16).to_bytes(32, "little") for v in meta["z"]))
    return w


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("precompute", [True, False])
def test_cpp_prove_matches_golden(cli, tmp_path, name, precompute):
    """create_proof_with_reduction(CircomCircuit{r1cs, witness}, pk, r, s): .r1cs + arkworks pk bytes + witness in, proof bytes
    out -- bit-identical to the golden fixture (and so to the oracle and the Python host)."""
    meta, _, _ = load_golden(name)
    args = ["prove", "--r1cs", os.path.join(GOLDEN, name + ".r1cs"), "--pk", os.path.join(GOLDEN, name + ".pk.bin"),
            "--witness", _witness_file(tmp_path, meta), "--out", tmp_path / "proof.bin", "--repeat", 2]
    if int(meta["r"], 16) == 0 and int(meta["s"], 16) == 0:
        args.append("--no-zk")  # create_proof_with_reduction_no_zk
    else:
        args += ["--r", meta["r"], "--s", meta["s"]]
    if meta["reduction"] == "circom":
        args += ["--reduction", "circom"]
    if not precompute:
        args += ["--no-precompute", "--pageable"]  # the witness straight from the CircomCircuit's std::vector (default: page-locked copy)
    out = cli(*args)
    assert out.stdout.strip() == meta["proof_compressed"]
    assert (tmp_path / "proof.bin").read_bytes().hex() == meta["proof_uncompressed"]
    assert "kernel launches" in out.stderr and not out.stderr.startswith("g16_cli: 0 kernel")


@pytest.mark.gpu
@pytest.mark.parametrize("name,shards", [("rand300", 2), ("rand100", 3), ("dummy924_nozk", 4)])
def test_cpp_sharded_prove_matches_golden(cli, tmp_path, name, shards):
    meta, _, _ = load_golden(name)
    args = ["prove", "--r1cs", os.path.join(GOLDEN, name + ".r1cs"), "--pk", os.path.join(GOLDEN, name + ".pk.bin"),
            "--witness", _witness_file(tmp_path, meta), "--out", tmp_path / "proof.bin", "--shards", shards, "--repeat", 2]
    args += ["--no-zk"] if int(meta["r"], 16) == 0 and int(meta["s"], 16) == 0 else ["--r", meta["r"], "--s", meta["s"]]
    assert cli(*args).stdout.strip() == meta["proof_compressed"]
    assert (tmp_path / "proof.bin").read_bytes().hex() == meta["proof_uncompressed"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rand100", "rand100_circom", "rand300"])
def test_cpp_witness_map_matches_golden(cli, tmp_path, name):
    meta, _, _ = load_golden(name)
    args = ["witness-map", "--r1cs", os.path.join(GOLDEN, name + ".r1cs"), "--witness", _witness_file(tmp_path, meta),
            "--out", tmp_path / "h.bin"]
    if meta["reduction"] == "circom":
        args += ["--reduction", "circom"]
    assert int(cli(*args).stdout) == meta["domain_size"]
    raw = (tmp_path / "h.bin").read_bytes()
    assert [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)] == [int(v, 16) for v in meta["h"]]


@pytest.mark.gpu
def test_cpp_seeded_rng_prove_equals_python_host_with_same_rng(cli, tmp_path):
    """create_random_proof_with_reduction with ark_std::test_rng(): the C++ host and the Python host draw the same (r, s)
    and return the same proof bytes."""
    name = "rand300"
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    out = cli("prove", "--r1cs", os.path.join(GOLDEN, name + ".r1cs"), "--pk", os.path.join(GOLDEN, name + ".pk.bin"),
              "--witness", _witness_file(tmp_path, meta), "--out", tmp_path / "proof.bin", "--test-rng")
    mats = load_matrices(r1cs_bytes)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    prover = g.Groth16(0)
    try:
        proof = prover.create_random_proof_with_reduction(pk, mats, mats.num_instance_variables, mats.num_constraints,
                                                          [int(v, 16) for v in meta["z"]], rust_rng.test_rng())
    finally:
        prover.close()
    assert out.stdout.strip() == proof.serialize_compressed().hex()
    assert out.stdout.strip() != meta["proof_compressed"]  # different (r, s) than the fixture's


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cpp_verify_matches_oracle_verdicts(cli, tmp_path, name):
    """Groth16Verifier (verifier.rs mirror): arkworks pk bytes + uncompressed proof bytes + public inputs in, one verdict per
    proof out; PreparedVerifyingKey.alpha_g1_beta_g2 and prepare_inputs byte-identical to tests/golden/verify.json."""
    import json
    with open(os.path.join(GOLDEN, "verify.json")) as f:
        fx = json.load(f)[name]
    meta, _, _ = load_golden(name)
    ni = int(meta["num_instance"])
    good = bytes.fromhex(meta["proof_uncompressed"])
    swapped = good[192:] + good[64:192] + good[:64]  # A and C exchanged (flags travel with the points)
    (tmp_path / "good.bin").write_bytes(good)
    (tmp_path / "swapped.bin").write_bytes(swapped)
    inputs = ",".join(meta["z"][1:ni])
    wrong = ",".join([hex((int(meta["z"][1], 16) + 1) % o.R_MOD)] + meta["z"][2:ni]) if ni > 1 else ""
    rows = ";".join([inputs, inputs, wrong]) if ni > 1 else ""
    proofs = [tmp_path / "good.bin", tmp_path / "swapped.bin"] + ([tmp_path / "good.bin"] if ni > 1 else [])
    out = cli("verify", "--pk", os.path.join(GOLDEN, name + ".pk.bin"), "--proof", ",".join(map(str, proofs)), "--inputs", rows,
              "--gt-out", tmp_path / "gt.bin", "--prepared-out", tmp_path / "pi.bin")
    assert out.stdout.split() == ["1", "0"] + (["0"] if ni > 1 else [])
    gt = (tmp_path / "gt.bin").read_bytes()
    assert [hex(int.from_bytes(gt[i:i + 32], "little")) for i in range(0, 384, 32)] == fx["alpha_g1_beta_g2"]
    assert (tmp_path / "pi.bin").read_bytes().hex() == fx["prepared_inputs"]
    # a public-input vector of the wrong length is the reference's MalformedVerifyingKey error, not a verdict
    bad = cli("verify", "--pk", os.path.join(GOLDEN, name + ".pk.bin"), "--proof", tmp_path / "good.bin",
              "--inputs", (inputs + ",0x1") if inputs else "0x1", check=False)
    assert bad.returncode != 0 and "gamma_abc_g1" in bad.stderr


@pytest.mark.parametrize("name", ["rand300", "dummy924_nozk", "silly"])
def test_cli_file_writers_round_trip_the_goldens(cli, tmp_path, name):
    """r1cs-write / pk-write (used to mint full-size fixtures in the reference's file formats): the arkworks key bytes are
    reproduced exactly (flags included) from canonical point dumps, and a .r1cs written from CSR dumps parses back to the
    same matrices through both hosts."""
    import numpy as np
    meta, r1cs_bytes, pk_bytes = load_golden(name)
    pk = g.ProvingKey.deserialize_uncompressed_unchecked(pk_bytes)
    dumps = dict(alpha_g1=pk.arrays["alpha_g1"], beta_g2=pk.arrays["beta_g2"], gamma_g2=pk.raw_vk["gamma_g2"],
                 vk_delta_g1=pk.raw_vk["delta_g1"], delta_g2=pk.arrays["delta_g2"], gamma_abc_g1=pk.raw_vk["gamma_abc_g1"],
                 beta_g1=pk.arrays["beta_g1"], delta_g1=pk.arrays["delta_g1"], a_query=pk.arrays["a_query"],
                 b_g1_query=pk.arrays["b_g1_query"], b_g2_query=pk.arrays["b_g2_query"], h_query=pk.arrays["h_query"],
                 l_query=pk.arrays["l_query"])
    for k, v in dumps.items():
        np.ascontiguousarray(v, dtype="<u8").tofile(tmp_path / f"pk.{k}.bin")
    cli("pk-write", "--prefix", tmp_path / "pk", "--out", tmp_path / "pk.bin")
    assert (tmp_path / "pk.bin").read_bytes() == pk_bytes
    mats = load_matrices(r1cs_bytes)
    for k in range(3):
        np.ascontiguousarray(mats.row_ptr[k], dtype="<u8").tofile(tmp_path / f"m.{k}.ptr")
        np.ascontiguousarray(mats.col[k], dtype="<u4").tofile(tmp_path / f"m.{k}.col")
        np.ascontiguousarray(mats.val[k], dtype="<u8").tofile(tmp_path / f"m.{k}.val")
    m = mats.num_instance_variables + mats.num_witness_variables
    cli("r1cs-write", "--prefix", tmp_path / "m", "--nc", mats.num_constraints, "--nwires", m, "--ninputs",
        mats.num_instance_variables, "--out", tmp_path / "m.r1cs")
    back = load_matrices((tmp_path / "m.r1cs").read_bytes())
    assert (back.num_instance_variables, back.num_witness_variables, back.num_constraints) == (
        mats.num_instance_variables, mats.num_witness_variables, mats.num_constraints)
    for k in range(3):
        assert np.array_equal(back.row_ptr[k], mats.row_ptr[k]) and np.array_equal(back.col[k], mats.col[k])
        assert np.array_equal(back.val[k], mats.val[k])
    assert cli("r1cs", tmp_path / "m.r1cs").stdout == cli("r1cs", os.path.join(GOLDEN, name + ".r1cs")).stdout
    stats = cli("r1cs-bench", tmp_path / "m.r1cs").stdout
    assert f'"constraints": {mats.num_constraints}' in stats
