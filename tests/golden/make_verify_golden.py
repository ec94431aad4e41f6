"""Writes tests/golden/verify.json: GT / G1 values the verifier (row f-4) must reproduce, computed by oracle/pairing.py.
The reference holds no golden for the verifier either (its tests are prove -> verify round trips); these pin the oracle's
outputs so that the CUDA path, the host mirrors and later rounds are compared with fixed bytes.
Run:  python tests/golden/make_verify_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import pairing as P  # noqa: E402
import pyref as o  # noqa: E402
from conftest import GOLDEN_NAMES  # noqa: E402
from test_pairing_host import load_vk_and_proof  # noqa: E402


def main():
    fx = {"pairing_generators": [hex(v) for v in P.to_tower(P.pairing(o.G1_GEN, o.G2_GEN))]}
    for name in GOLDEN_NAMES:
        vk, proof, inputs = load_vk_and_proof(name)
        pvk = P.prepare_verifying_key(vk)
        assert P.verify_proof(pvk, proof, inputs)
        f = P.multi_miller_loop([proof[0], o.G1.to_affine(P.prepare_inputs(pvk, inputs)), proof[2]],
                                [proof[1], pvk.gamma_g2_neg, pvk.delta_g2_neg])
        fx[name] = {
            "alpha_g1_beta_g2": [hex(v) for v in P.to_tower(pvk.alpha_g1_beta_g2)],
            "prepared_inputs": o.ser_g1(o.G1.to_affine(P.prepare_inputs(pvk, inputs)), False).hex(),
            "miller_loop": [hex(v) for v in P.to_tower(f)],
            "accepts": True,
        }
    with open(os.path.join(HERE, "verify.json"), "w") as f:
        json.dump(fx, f)
    print("wrote verify.json")


if __name__ == "__main__":
    main()
