"""Generates the golden fixtures under tests/golden/ from the big-integer oracle (oracle/pyref.py).

The reference cannot be built here and holds no golden proof, so these vectors are the pinned oracle outputs for the
reference's own test circuits (MySillyCircuit, forks/groth16/src/test.rs:14-43; DummyCircuit,
creds/src/rangeproof.rs:446-486) and for seeded random satisfiable R1CS instances.  Every fixture is self-checked:
the proof's discrete logs satisfy the Groth16 equation (verify_in_exponent) and match the closed form.
Run:  python tests/golden/make_golden.py   (about a minute of pure-Python big-integer arithmetic)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as o  # noqa: E402


def matrices_to_r1cs_bytes(m):
    cons = []
    for i in range(m.num_constraints):
        cons.append(tuple([(w, v) for v, w in m.__dict__[k][i]] for k in ("a", "b", "c")))
    # circom header: 1 + n_pub_out + n_pub_in = num_instance
    return o.write_r1cs(m.num_instance_variables + m.num_witness_variables, m.num_instance_variables - 1, 0,
                        m.num_witness_variables, cons)


def fixture(name, m, z, td, r, s, reduction="libsnark"):
    ni, nc = m.num_instance_variables, m.num_constraints
    pk, qap = o.generate_parameters(m, td, reduction)
    proof, h, msms = o.create_proof_with_reduction_and_matrices(pk, r, s, m, ni, nc, z, reduction)
    A, B, C = o.proof_scalars_closed_form(td, qap, m, z, h, r, s)
    assert o.G1.mul(o.G1_GEN, A) == proof[0] and o.G2.mul(o.G2_GEN, B) == proof[1] and o.G1.mul(o.G1_GEN, C) == proof[2]
    assert o.verify_in_exponent(td, qap, m, z, A, B, C)
    # round trip of the .r1cs writer/reader and of the matrix flattening
    rb = matrices_to_r1cs_bytes(m)
    m2 = o.r1cs_to_matrices(o.read_r1cs(rb))
    assert (m2.a, m2.b, m2.c) == ([sorted(r_, key=lambda t: t[1]) for r_ in m.a], [sorted(r_, key=lambda t: t[1]) for r_ in m.b],
                                  [sorted(r_, key=lambda t: t[1]) for r_ in m.c])
    meta = dict(
        name=name, reduction=reduction, num_instance=ni, num_witness=m.num_witness_variables, num_constraints=nc,
        domain_size=qap[4], r=hex(r), s=hex(s),
        trapdoor={k: hex(getattr(td, k)) for k in ("alpha", "beta", "gamma", "delta", "t")},
        z=[hex(v) for v in z], h=[hex(v) for v in h],
        msm={k: (o.ser_g2(v, False) if k == "b_g2" else o.ser_g1(v, False)).hex() for k, v in msms.items()},
        proof_compressed=o.ser_proof(proof, True).hex(), proof_uncompressed=o.ser_proof(proof, False).hex(),
        proof_dlog=[hex(A), hex(B), hex(C)],
    )
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    with open(os.path.join(HERE, name + ".r1cs"), "wb") as f:
        f.write(rb)
    with open(os.path.join(HERE, name + ".pk.bin"), "wb") as f:
        f.write(o.ser_pk(pk, False))
    print(name, "ok: n =", qap[4], "pk bytes", len(o.ser_pk(pk, False)))


def main():
    td = o.Trapdoor(alpha=o.stream_fr(0xA11CE, 1), beta=o.stream_fr(0xA11CE, 2), gamma=1, delta=o.stream_fr(0xA11CE, 3),
                    t=o.stream_fr(0xA11CE, 4))  # gamma = 1 as in the fork's generator (generator.rs:28)
    r, s = o.stream_fr(0xBEEF, 1), o.stream_fr(0xBEEF, 2)
    m, z = o.my_silly_circuit(o.stream_fr(7, 1), o.stream_fr(7, 2))
    fixture("silly", m, z, td, r, s)
    fixture("silly_nozk", m, z, td, 0, 0)  # create_proof_with_reduction_no_zk: r = s = 0, B-in-G1 skipped
    m, z = o.random_satisfiable_r1cs(11, 100, 4, 60)
    fixture("rand100", m, z, td, r, s)
    fixture("rand100_circom", m, z, td, r, s, "circom")
    m, z = o.random_satisfiable_r1cs(12, 300, 9, 250, max_row=6)
    fixture("rand300", m, z, td, r, s)
    m, z = o.dummy_circuit(7, o.stream_fr(0xD0, 1))
    fixture("dummy924_nozk", m, z, td, 0, 0)  # the shape creds/src/rangeproof.rs:490-511 proves with r = s = 0


if __name__ == "__main__":
    main()
