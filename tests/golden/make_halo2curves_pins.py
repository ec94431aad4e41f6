"""Writes tests/golden/halo2curves_bn256_pins.json: BN254 pairing constants that the REFERENCE TREE itself holds (its
forks/halo2curves copy -- a different library from the arkworks one Crescent's verifier uses, same curve and same tower),
extracted as data so that the oracle and the device constants can be pinned to in-tree bytes on machines without
/root/reference:
    forks/halo2curves/src/bn256/mod.rs:17,20-24      BN_X, SIX_U_PLUS_2_NAF (the signed digits of 6x+2)
    forks/halo2curves/src/bn256/fq12.rs:40-...       FROBENIUS_COEFF_FQ12_C1[i] = xi^((q^i - 1)/6), Montgomery limbs
    forks/halo2curves/src/bn256/fq6.rs:46-...        FROBENIUS_COEFF_FQ6_C1[i] = xi^((q^i - 1)/3), _C2[i] = xi^(2(q^i - 1)/3)
    forks/halo2curves/src/bn256/engine.rs:164-177    XI_TO_Q_MINUS_1_OVER_2
Only numbers are copied (no code).  Run here (needs /root/reference):  python tests/golden/make_halo2curves_pins.py"""
import json
import os
import re

REF = "/root/reference/forks/halo2curves/src/bn256"
HERE = os.path.dirname(os.path.abspath(__file__))


def fq2_array(text, name):
    body = text[text.index("const " + name):]  # the definition, not a use
    body = body[body.index("=") + 1:]
    depth, end = 0, 0
    for i, ch in enumerate(body):
        if ch == "[" and depth == 0 and body[:i].strip() == "":
            depth = 1
            continue
        if depth:
            if ch == "[":
                depth += 1
            elif ch == "]":
                depth -= 1
                if depth == 0:
                    end = i
                    break
    body = body[:end]
    limbs = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", re.sub(r"//[^\n]*", "", body))]
    assert len(limbs) % 8 == 0, (name, len(limbs))
    return [[hex(v) for v in limbs[i:i + 8]] for i in range(0, len(limbs), 8)]  # c0 limbs (4) then c1 limbs (4), LE u64


def main():
    mod = open(os.path.join(REF, "mod.rs")).read()
    bn_x = int(re.search(r"BN_X: u64 = (\d+)", mod).group(1))
    naf = re.search(r"SIX_U_PLUS_2_NAF: \[i8; 65\] = \[(.*?)\];", mod, re.S).group(1)
    naf = [int(v) for v in re.findall(r"-?\d+", naf)]
    assert len(naf) == 65
    fq12 = open(os.path.join(REF, "fq12.rs")).read()
    fq6 = open(os.path.join(REF, "fq6.rs")).read()
    eng = open(os.path.join(REF, "engine.rs")).read()
    m = re.search(r"XI_TO_Q_MINUS_1_OVER_2: Fq2 = Fq2 \{(.*?)\};", eng, re.S).group(1)
    over2 = [hex(int(x, 16)) for x in re.findall(r"0x[0-9a-fA-F]+", m)]
    out = {
        "source": "forks/halo2curves/src/bn256/{mod,fq12,fq6,engine}.rs of the reference tree; Montgomery limbs (R = 2^256), little-endian u64",
        "BN_X": bn_x,
        "SIX_U_PLUS_2_NAF": naf,
        "FROBENIUS_COEFF_FQ12_C1": fq2_array(fq12, "FROBENIUS_COEFF_FQ12_C1"),
        "FROBENIUS_COEFF_FQ6_C1": fq2_array(fq6, "FROBENIUS_COEFF_FQ6_C1"),
        "FROBENIUS_COEFF_FQ6_C2": fq2_array(fq6, "FROBENIUS_COEFF_FQ6_C2"),
        "XI_TO_Q_MINUS_1_OVER_2": over2,
    }
    assert len(out["FROBENIUS_COEFF_FQ12_C1"]) == 12 and len(out["FROBENIUS_COEFF_FQ6_C1"]) == 6 and len(over2) == 8
    with open(os.path.join(HERE, "halo2curves_bn256_pins.json"), "w") as f:
        json.dump(out, f)
    print("wrote halo2curves_bn256_pins.json")


if __name__ == "__main__":
    main()
