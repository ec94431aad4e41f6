"""GPU parity tests for row f-4 (Groth16 verification): every call goes through the C ABI and is compared with the
big-integer oracle (oracle/pairing.py) and the committed fixture tests/golden/verify.json.  Bit-exact: GT elements, prepared
inputs and verdicts are integers / booleans."""
import json
import os

import numpy as np
import pytest

import pairing as P
import pyref as o
from conftest import GOLDEN, GOLDEN_NAMES, load_golden
from crescent_credentials_b200 import ffi, generator, synth
from crescent_credentials_b200 import groth16 as g
from crescent_credentials_b200 import verifier as v
from test_pairing_host import load_vk_and_proof

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fixture_json():
    with open(os.path.join(GOLDEN, "verify.json")) as f:
        return json.load(f)


def gt_ints(row):
    return g.fq_from_mont(np.asarray(row, dtype=np.uint64).reshape(-1, 4))


def test_pairing_matches_oracle_and_fixture(gpu_ctx, fixture_json):
    ks = [(1, 1), (o.stream_fr(0x9A1, 1), o.stream_fr(0x9A1, 2)), (o.R_MOD - 1, 2), (5, o.stream_fr(0x9A1, 3))]
    ps = [o.G1.mul(o.G1_GEN, a) for a, _ in ks] + [None, o.G1_GEN]
    qs = [o.G2.mul(o.G2_GEN, b) for _, b in ks] + [o.G2_GEN, None]
    got = gpu_ctx.pairing(g.g1_points_to_mont(ps), g.g2_points_to_mont(qs))
    for row, p, q in zip(got, ps, qs):
        assert gt_ints(row) == P.to_tower(P.pairing(p, q))
    assert [hex(x) for x in gt_ints(got[0])] == fixture_json["pairing_generators"]
    one = P.to_tower(P.F12_ONE)
    assert gt_ints(got[4]) == one and gt_ints(got[5]) == one  # a pair holding infinity is filtered out
    # bilinearity through the device alone: e(aG, bH) == e(abG, H)
    a, b = ks[1]
    both = gpu_ctx.pairing(g.g1_points_to_mont([o.G1.mul(o.G1_GEN, a * b % o.R_MOD)]), g.g2_points_to_mont([o.G2_GEN]))
    assert (both[0] == got[1]).all()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_prepared_key_and_inputs_match_fixture(name, fixture_json):
    vk, proof, inputs = load_vk_and_proof(name)
    ver = v.Verifier(0)
    try:
        pvk = ver.prepare_verifying_key(g.VerifyingKey(vk.alpha_g1, vk.beta_g2, vk.gamma_g2, vk.delta_g1, vk.delta_g2, vk.gamma_abc_g1))
        assert [hex(x) for x in pvk.alpha_g1_beta_g2] == fixture_json[name]["alpha_g1_beta_g2"]
        pi = ver.prepare_inputs(pvk, inputs)
        assert g._ser_g1(pi, False).hex() == fixture_json[name]["prepared_inputs"]
        with pytest.raises(v.MalformedVerifyingKey):
            ver.prepare_inputs(pvk, list(inputs) + [1])
        # the same key read from arkworks' uncompressed bytes (canonical words, converted on the device)
        _, _, pkb = load_golden(name)
        pvk2 = ver.prepare_verifying_key(g.ProvingKey.deserialize_uncompressed_unchecked(pkb))
        assert pvk2.alpha_g1_beta_g2 == pvk.alpha_g1_beta_g2
        assert ver.prepare_inputs(pvk2, inputs) == pi
    finally:
        ver.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_verify_golden_proofs_and_tampered_ones(name):
    """verifier.rs:44-76 on the committed proofs: the verdict of every (proof, inputs) pair equals the oracle's."""
    vk, proof, inputs = load_vk_and_proof(name)
    opvk = P.prepare_verifying_key(vk)
    A, B, C = proof
    cases = [((A, B, C), inputs), ((C, B, A), inputs), ((A, B, o.G1.add(C, o.G1_GEN)), inputs),
             ((A, o.G2.add(B, o.G2_GEN), C), inputs), ((o.G1.neg(A), o.G2.neg(B), C), inputs), ((None, B, C), inputs),
             ((A, None, C), inputs), ((A, B, None), inputs)]
    if inputs:
        cases += [((A, B, C), [(inputs[0] + 1) % o.R_MOD] + list(inputs[1:])), ((A, B, C), [0] * len(inputs))]
    want = [P.verify_proof(opvk, pr, x) for pr, x in cases]
    assert want[0] is True and want[4] is True and not any(want[1:4])
    ver = v.Verifier(0)
    try:
        pvk = ver.prepare_verifying_key(g.VerifyingKey(vk.alpha_g1, vk.beta_g2, vk.gamma_g2, vk.delta_g1, vk.delta_g2, vk.gamma_abc_g1))
        got = ver.verify_proofs(pvk, [g.Proof(*pr) for pr, _ in cases], [x for _, x in cases])
        assert got == want
        assert ver.verify_proof(pvk, g.Proof(A, B, C), inputs) is True
        pi = ver.prepare_inputs(pvk, inputs)
        assert ver.verify_proof_with_prepared_inputs(pvk, g.Proof(A, B, C), pi) is True
        assert ver.verify_proof_with_prepared_inputs(pvk, g.Proof(A, B, C), o.G1.add(pi, o.G1_GEN)) is False
        assert ver.verify_proofs(pvk, [], []) == []
        assert ver.verify_with_processed_vk(pvk, inputs, g.Proof(A, B, C)) is True  # SNARK trait spelling (lib.rs:89-96)
        with pytest.raises(v.MalformedVerifyingKey):
            ver.verify_proof(pvk, g.Proof(A, B, C), list(inputs) + [7])
    finally:
        ver.close()


def test_prove_then_verify_round_trip_on_the_device(gpu_ctx):
    """The reference's own test shape (forks/groth16/src/test.rs:45-73): mint a key, prove, verify, reject a wrong input --
    here with generator, prover and verifier all on the GPU, on the S-2^12 instance, for a batch of proofs."""
    inst = synth.make_instance(gpu_ctx, "S-2^12", seed=0x7E57)
    td = generator.Trapdoor(alpha=o.stream_fr(0xC0DE, 1), beta=o.stream_fr(0xC0DE, 2), gamma=1, delta=o.stream_fr(0xC0DE, 3),
                            t=o.stream_fr(0xC0DE, 4))
    pk, _ = generator.generate_parameters_with_qap(gpu_ctx, inst.matrices, td)
    prover = g.Groth16(0)
    ver = v.Verifier(0)
    try:
        proofs = [prover.create_proof_with_reduction_and_matrices(pk, o.stream_fr(0xABC, 2 * i + 1), o.stream_fr(0xABC, 2 * i + 2),
                                                                  inst.matrices, inst.ni, inst.nc, inst.z_mont) for i in range(3)]
        proofs.append(prover.create_proof_with_reduction_no_zk(pk, inst.matrices, inst.ni, inst.nc, inst.z_mont))
        public = g.fr_from_mont(inst.z_mont[1:inst.ni])
        pvk = ver.prepare_verifying_key(pk)
        assert pvk.num_public_inputs == inst.ni - 1
        batch = proofs + [g.Proof(proofs[0].a, proofs[1].b, proofs[0].c), proofs[2]]
        xs = [public] * 5 + [[(public[0] + 1) % o.R_MOD] + public[1:]]
        assert ver.verify_proofs(pvk, batch, xs) == [True, True, True, True, False, False]
        # a larger batch than one block per SM row, alternating verdicts
        many = [proofs[i % 4] if i % 3 else g.Proof(proofs[0].c, proofs[0].b, proofs[0].a) for i in range(200)]
        assert ver.verify_proofs(pvk, many, [public] * 200) == [bool(i % 3) for i in range(200)]
    finally:
        prover.close()
        ver.close()


def test_verify_without_a_key_fails_loudly(gpu_ctx):
    ctx = ffi.Context(0)
    try:
        with pytest.raises(ffi.G16Error):
            ctx.vk_alpha_beta()
        with pytest.raises(ffi.G16Error):
            ctx.verify_batch(np.zeros(272, dtype=np.uint8), np.zeros((1, 0, 4), dtype=np.uint64), 1)
    finally:
        ctx.close()
