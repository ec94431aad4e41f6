"""
oracle/pyref.py -- big-integer CPU restatement of the Groth16 `prove` hot path of
microsoft/crescent-credentials (fork of ark-groth16 0.4) over BN254.

THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product path (libg16b200.so and the
crescent_credentials_b200 package) never does.

PARITY STATUS: the reference prover cannot be built in this container (no cargo/rustc, arkworks
crates not vendored), and the reference tree holds no golden proof / H vector / MSM output.
What *is* pinned here, byte for byte, against the reference tree:
  * Montgomery LE encodings of Fq::one, the G1 generator and the G2 generator
    (forks/circom-compat/src/zkey.rs:397-431, decimal coordinates :442-462),
  * the iden3 .r1cs layout with its worked example (forks/circom-compat/src/circom/r1cs_reader.rs:266-344),
  * the BN254 Fr modulus bytes (r1cs_reader.rs:182-190).
Everything about arkworks' *internal* conventions (domain generator, coset shift, serialisation
flag bits) is restated from the published ark-ff / ark-poly / ark-ec / ark-serialize 0.4 sources
and marked ASSUMPTION below: "parity unpinned" for those until an arkworks binary is available.
Mathematical determinism carries the rest: H coefficients, MSM results and proof points are
unique exact values, independent of evaluation order.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple, Optional

# --------------------------------------------------------------------------------------------
# BN254 constants (ark-bn254 0.4.0; cross-checked in-tree against forks/halo2curves/src/bn256/fr.rs:7-15,
# fq.rs:9-17 and forks/circom-compat/src/circom/r1cs_reader.rs:183)
# --------------------------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus r
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus q
FR_GENERATOR = 5            # ASSUMPTION: ark-bn254 Fr::GENERATOR = 5 (halo2curves uses 7; do not mix)
FR_TWO_ADICITY = 28
MONT_R = 1 << 256           # Montgomery radix for both 4x64-bit fields
G1_B = 3                    # y^2 = x^3 + 3
G1_GEN = (1, 2)
# G2 generator, decimal in forks/circom-compat/src/zkey.rs:442-462
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
# 2^28-th root of unity rho = 5^((r-1)/2^28)
FR_ROOT_2_28 = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)


# --------------------------------------------------------------------------------------------
# Fr / Fq helpers (plain ints mod p) and Fq2 = Fq[u]/(u^2+1) as (c0, c1)
# --------------------------------------------------------------------------------------------
def inv_mod(a: int, p: int) -> int:
    if a % p == 0:
        raise ZeroDivisionError("inverse of zero")
    return pow(a, -1, p)


def to_mont(a: int, p: int) -> int:
    return (a << 256) % p


def from_mont(a: int, p: int) -> int:
    return (a * inv_mod(MONT_R % p, p)) % p


def le_bytes(a: int, n: int = 32) -> bytes:
    return int(a).to_bytes(n, "little")


def mont_le_bytes(a: int, p: int) -> bytes:
    """arkworks in-memory form: 4xu64 LE limbs of a*R mod p == 32 LE bytes."""
    return le_bytes(to_mont(a, p))


class Fq2:
    """Fq2 = Fq[u]/(u^2 + 1); ark-bn254 Fq2Config NONRESIDUE = -1 (halo2curves fq2.rs:50 agrees)."""
    __slots__ = ()

    @staticmethod
    def add(a, b):
        return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)

    @staticmethod
    def sub(a, b):
        return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)

    @staticmethod
    def neg(a):
        return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)

    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)

    @staticmethod
    def sqr(a):
        return Fq2.mul(a, a)

    @staticmethod
    def inv(a):
        n = inv_mod((a[0] * a[0] + a[1] * a[1]) % Q_MOD, Q_MOD)
        return ((a[0] * n) % Q_MOD, (-a[1] * n) % Q_MOD)

    @staticmethod
    def scal(a, k: int):
        return ((a[0] * k) % Q_MOD, (a[1] * k) % Q_MOD)

    ZERO = (0, 0)
    ONE = (1, 0)


# G2 curve coefficient b' = 3/(9+u)  (forks/halo2curves/src/bn256/curve.rs:83-96)
G2_B = Fq2.mul((3, 0), Fq2.inv((9, 1)))


# --------------------------------------------------------------------------------------------
# Generic short-Weierstrass a=0 group law, parameterised by a tiny field vtable.
# Points: None == infinity, else affine (x, y).  Jacobian (X, Y, Z) internally.
# --------------------------------------------------------------------------------------------
class _FqOps:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % Q_MOD)
    sub = staticmethod(lambda a, b: (a - b) % Q_MOD)
    mul = staticmethod(lambda a, b: (a * b) % Q_MOD)
    neg = staticmethod(lambda a: (-a) % Q_MOD)
    inv = staticmethod(lambda a: inv_mod(a, Q_MOD))
    is_zero = staticmethod(lambda a: a % Q_MOD == 0)


class _Fq2Ops:
    zero, one = Fq2.ZERO, Fq2.ONE
    add, sub, mul, neg, inv = Fq2.add, Fq2.sub, Fq2.mul, Fq2.neg, Fq2.inv
    is_zero = staticmethod(lambda a: a[0] % Q_MOD == 0 and a[1] % Q_MOD == 0)


class Curve:
    def __init__(self, F, b, gen):
        self.F, self.b, self.gen = F, b, gen

    def is_on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)) == F.zero

    # Jacobian ----------------------------------------------------------------
    def to_jac(self, P):
        return None if P is None else (P[0], P[1], self.F.one)

    def to_affine(self, J):
        if J is None:
            return None
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.mul(zi, zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jdouble(self, J):
        if J is None:
            return None
        F = self.F
        X, Y, Z = J
        if F.is_zero(Y):
            return None
        A = F.mul(X, X)
        B = F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.mul(E, E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jadd(self, J1, J2):
        if J1 is None:
            return J2
        if J2 is None:
            return J1
        F = self.F
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        Z1Z1 = F.mul(Z1, Z1)
        Z2Z2 = F.mul(Z2, Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jdouble(J1)
            return None
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jneg(self, J):
        return None if J is None else (J[0], self.F.neg(J[1]), J[2])

    # affine convenience -------------------------------------------------------
    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def jmul(self, J, k: int):
        """k*J, k any non-negative integer (mul_bigint semantics: no reduction mod r)."""
        acc = None
        for bit in bin(k)[2:] if k else "":
            acc = self.jdouble(acc)
            if bit == "1":
                acc = self.jadd(acc, J)
        return acc

    def mul(self, P, k: int):
        return self.to_affine(self.jmul(self.to_jac(P), k % R_MOD))

    def msm(self, points: Sequence, scalars: Sequence[int]):
        """Exact Sigma s_i * P_i over min(len) pairs -- VariableBaseMSM::msm_bigint call-site semantics
        (forks/groth16/src/prover.rs:66,74,266).  Returns Jacobian.  Simple bucket method, 8-bit windows."""
        n = min(len(points), len(scalars))
        c = 8
        acc = None
        for w in reversed(range((256 + c - 1) // c)):
            for _ in range(c):
                acc = self.jdouble(acc)
            buckets = [None] * ((1 << c) - 1)
            any_ = False
            for i in range(n):
                d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
                if d and points[i] is not None:
                    buckets[d - 1] = self.jadd(buckets[d - 1], self.to_jac(points[i]))
                    any_ = True
            if any_:
                run = None
                tot = None
                for bkt in reversed(buckets):
                    run = self.jadd(run, bkt)
                    tot = self.jadd(tot, run)
                acc = self.jadd(acc, tot)
        return acc

    def fixed_base_table(self, window: int = 8):
        """table[w][d] = d * 2^(window*w) * gen, affine, for fixed-base multiplication (generator.rs:133-194
        uses FixedBase::msm; only the resulting points matter)."""
        tbl = []
        base = self.to_jac(self.gen)
        for _ in range((254 + window - 1) // window):
            row = [None]
            cur = None
            for _d in range(1, 1 << window):
                cur = self.jadd(cur, base)
                row.append(cur)
            tbl.append(row)
            for _ in range(window):
                base = self.jdouble(base)
        return tbl

    def fixed_mul_j(self, tbl, k: int, window: int = 8):
        acc = None
        k %= R_MOD
        w = 0
        while k:
            d = k & ((1 << window) - 1)
            if d:
                acc = self.jadd(acc, tbl[w][d])
            k >>= window
            w += 1
        return acc


G1 = Curve(_FqOps, G1_B, G1_GEN)
G2 = Curve(_Fq2Ops, G2_B, G2_GEN)


# --------------------------------------------------------------------------------------------
# Radix-2 evaluation domain (ark-poly 0.4 Radix2EvaluationDomain, ASSUMPTION-level conventions):
#   size n = 2^ceil(log2 k); generator omega = rho^(2^(28-log n)); ifft scales by 1/n;
#   coset FFT = multiply coeff i by g^i then FFT; coset iFFT = iFFT then multiply by g^-i;
#   Z(x) = x^n - 1.
# --------------------------------------------------------------------------------------------
class Domain:
    def __init__(self, min_size: int):
        n = 1
        log_n = 0
        while n < min_size:
            n <<= 1
            log_n += 1
        if log_n > FR_TWO_ADICITY:
            raise ValueError("PolynomialDegreeTooLarge")  # r1cs_to_qap.rs:156-157
        self.n, self.log_n = n, log_n
        self.omega = pow(FR_ROOT_2_28, 1 << (FR_TWO_ADICITY - log_n), R_MOD)
        self.omega_inv = inv_mod(self.omega, R_MOD)
        self.n_inv = inv_mod(n, R_MOD)

    def element(self, i: int) -> int:
        return pow(self.omega, i, R_MOD)

    def _ntt(self, a: List[int], w: int) -> List[int]:
        n = self.n
        a = list(a) + [0] * (n - len(a))
        # bit-reverse, then iterative DIT
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, R_MOD)
            half = length >> 1
            for s in range(0, n, length):
                t = 1
                for k in range(s, s + half):
                    u = a[k]
                    v = a[k + half] * t % R_MOD
                    a[k] = (u + v) % R_MOD
                    a[k + half] = (u - v) % R_MOD
                    t = t * wl % R_MOD
            length <<= 1
        return a

    def fft(self, a):
        return self._ntt(a, self.omega)

    def ifft(self, a):
        return [x * self.n_inv % R_MOD for x in self._ntt(a, self.omega_inv)]

    def coset_fft(self, a, g=FR_GENERATOR):
        a = list(a) + [0] * (self.n - len(a))
        p = 1
        out = []
        for x in a:
            out.append(x * p % R_MOD)
            p = p * g % R_MOD
        return self.fft(out)

    def coset_ifft(self, a, g=FR_GENERATOR):
        a = self.ifft(a)
        gi = inv_mod(g, R_MOD)
        p = 1
        out = []
        for x in a:
            out.append(x * p % R_MOD)
            p = p * gi % R_MOD
        return out

    def vanishing(self, x: int) -> int:
        return (pow(x, self.n, R_MOD) - 1) % R_MOD

    def lagrange_at(self, t: int) -> List[int]:
        """evaluate_all_lagrange_coefficients(t) for t outside the domain:
        L_i(t) = Z(t) * omega^i / (n * (t - omega^i))."""
        zt = self.vanishing(t)
        out = []
        wi = 1
        for _ in range(self.n):
            out.append(zt * wi % R_MOD * inv_mod(self.n * (t - wi) % R_MOD, R_MOD) % R_MOD)
            wi = wi * self.omega % R_MOD
        return out


# --------------------------------------------------------------------------------------------
# R1CS matrices in the shape of ark-relations ConstraintMatrices:
#   a, b, c: list of rows, each a list of (coeff, column)
# --------------------------------------------------------------------------------------------
class Matrices:
    def __init__(self, num_instance: int, num_witness: int, a, b, c):
        self.num_instance_variables = num_instance
        self.num_witness_variables = num_witness
        self.num_constraints = len(a)
        self.a, self.b, self.c = a, b, c


def evaluate_constraint(terms, z) -> int:
    """forks/groth16/src/r1cs_to_qap.rs:16-45 (the coeff==1 fast path is value-neutral)."""
    s = 0
    for coeff, idx in terms:
        s += coeff * z[idx]
    return s % R_MOD


def witness_map_libsnark(m: Matrices, num_inputs: int, num_constraints: int, z: Sequence[int]) -> List[int]:
    """LibsnarkReduction::witness_map_from_matrices, forks/groth16/src/r1cs_to_qap.rs:150-213."""
    dom = Domain(num_constraints + num_inputs)                       # :156-158
    n = dom.n
    a = [0] * n
    b = [0] * n
    for i in range(num_constraints):                                 # :164-171
        a[i] = evaluate_constraint(m.a[i], z)
        b[i] = evaluate_constraint(m.b[i], z)
    for i in range(num_inputs):                                      # :173-177
        a[num_constraints + i] = z[i] % R_MOD
    a = dom.ifft(a)                                                  # :179-180
    b = dom.ifft(b)
    a = dom.coset_fft(a)                                             # :182-185
    b = dom.coset_fft(b)
    ab = [x * y % R_MOD for x, y in zip(a, b)]                       # :187
    c = [0] * n
    for i in range(num_constraints):                                 # :191-196
        c[i] = evaluate_constraint(m.c[i], z)
    c = dom.ifft(c)                                                  # :198-199
    c = dom.coset_fft(c)
    zinv = inv_mod(dom.vanishing(FR_GENERATOR), R_MOD)               # :201-204
    ab = [(x - y) * zinv % R_MOD for x, y in zip(ab, c)]             # :205-208
    return dom.coset_ifft(ab)                                        # :210


def witness_map_circom(m: Matrices, num_inputs: int, num_constraints: int, z: Sequence[int]) -> List[int]:
    """CircomReduction::witness_map_from_matrices, forks/circom-compat/src/circom/qap.rs:25-90."""
    dom = Domain(num_constraints + num_inputs)
    n = dom.n
    a = [0] * n
    b = [0] * n
    for i in range(num_constraints):
        a[i] = evaluate_constraint(m.a[i], z)
        b[i] = evaluate_constraint(m.b[i], z)
    for i in range(num_inputs):
        a[num_constraints + i] = z[i] % R_MOD
    c = [0] * n
    for i in range(num_constraints):                                 # qap.rs:52-58
        c[i] = a[i] * b[i] % R_MOD
    a = dom.ifft(a)
    b = dom.ifft(b)
    root = Domain(2 * n).element(1)                                  # qap.rs:63-68
    a = dom.coset_fft(a, root)                                       # distribute_powers + fft, :69-73
    b = dom.coset_fft(b, root)
    ab = [x * y % R_MOD for x, y in zip(a, b)]
    c = dom.ifft(c)
    c = dom.coset_fft(c, root)
    return [(x - y) % R_MOD for x, y in zip(ab, c)]                  # :84-88 (evaluations, no division)


# --------------------------------------------------------------------------------------------
# Proving key with a known trapdoor (generator.rs:50-228 + r1cs_to_qap.rs:106-148,215-225)
# --------------------------------------------------------------------------------------------
class VerifyingKey:
    def __init__(self):
        self.alpha_g1 = self.beta_g2 = self.gamma_g2 = self.delta_g1 = self.delta_g2 = None
        self.gamma_abc_g1: List = []


class ProvingKey:
    def __init__(self):
        self.vk = VerifyingKey()
        self.beta_g1 = self.delta_g1 = None
        self.a_query: List = []
        self.b_g1_query: List = []
        self.b_g2_query: List = []
        self.h_query: List = []
        self.l_query: List = []


class Trapdoor:
    def __init__(self, alpha, beta, gamma, delta, t):
        self.alpha, self.beta, self.gamma, self.delta, self.t = alpha, beta, gamma, delta, t


def instance_map_with_evaluation(m: Matrices, t: int):
    """LibsnarkReduction::instance_map_with_evaluation, r1cs_to_qap.rs:106-148."""
    nc = m.num_constraints
    ni = m.num_instance_variables
    dom = Domain(nc + ni)
    zt = dom.vanishing(t)
    u = dom.lagrange_at(t)
    qap_num_variables = (ni - 1) + m.num_witness_variables
    a = [0] * (qap_num_variables + 1)
    b = [0] * (qap_num_variables + 1)
    c = [0] * (qap_num_variables + 1)
    for i in range(ni):                                              # :128-133
        a[i] = u[nc + i]
    for i in range(nc):                                              # :135-145
        ui = u[i]
        for coeff, idx in m.a[i]:
            a[idx] = (a[idx] + ui * coeff) % R_MOD
        for coeff, idx in m.b[i]:
            b[idx] = (b[idx] + ui * coeff) % R_MOD
        for coeff, idx in m.c[i]:
            c[idx] = (c[idx] + ui * coeff) % R_MOD
    return a, b, c, zt, qap_num_variables, dom.n


def generate_parameters(m: Matrices, td: Trapdoor, h_style: str = "libsnark"):
    """generate_parameters_with_qap, forks/groth16/src/generator.rs:50-228 with the fork's canonical
    generators (:34-35).  Returns (pk, qap) where qap = (a, b, c, zt, n) scalars for exponent checks."""
    ni = m.num_instance_variables
    a, b, c, zt, qap_nv, n = instance_map_with_evaluation(m, td.t)
    gi = inv_mod(td.gamma, R_MOD)
    di = inv_mod(td.delta, R_MOD)
    gamma_abc = [(td.beta * a[i] + td.alpha * b[i] + c[i]) * gi % R_MOD for i in range(ni)]       # :113-117
    l = [(td.beta * a[i] + td.alpha * b[i] + c[i]) * di % R_MOD for i in range(ni, qap_nv + 1)]  # :119-123
    if h_style == "libsnark":                                                                    # r1cs_to_qap.rs:215-225
        hs = [zt * di % R_MOD * pow(td.t, i, R_MOD) % R_MOD for i in range(n - 1)]
    else:                                                                                        # qap.rs:92-107
        sc = [di * pow(td.t, i, R_MOD) % R_MOD for i in range(2 * (n - 1) + 1)]
        d2 = Domain(len(sc))
        sc = d2.ifft(sc)
        hs = sc[1::2]
    t1 = G1.fixed_base_table()
    t2 = G2.fixed_base_table()
    f1 = lambda k: G1.to_affine(G1.fixed_mul_j(t1, k))
    f2 = lambda k: G2.to_affine(G2.fixed_mul_j(t2, k))
    pk = ProvingKey()
    pk.vk.alpha_g1 = f1(td.alpha)
    pk.vk.beta_g2 = f2(td.beta)
    pk.vk.gamma_g2 = f2(td.gamma)
    pk.vk.delta_g1 = f1(td.delta)          # fork-only field, data_structures.rs:39, generator.rs:204
    pk.vk.delta_g2 = f2(td.delta)
    pk.vk.gamma_abc_g1 = [f1(k) for k in gamma_abc]
    pk.beta_g1 = f1(td.beta)
    pk.delta_g1 = f1(td.delta)
    pk.a_query = [f1(k) for k in a]
    pk.b_g1_query = [f1(k) for k in b]
    pk.b_g2_query = [f2(k) for k in b]
    pk.h_query = [f1(k) for k in hs]
    pk.l_query = [f1(k) for k in l]
    return pk, (a, b, c, zt, n, hs, l)


# --------------------------------------------------------------------------------------------
# Prover (forks/groth16/src/prover.rs:26-136, 256-274)
# --------------------------------------------------------------------------------------------
def calculate_coeff(C: Curve, initial_j, query, vk_param, assignment):
    """prover.rs:256-274: initial + query[0] + MSM(query[1..], assignment) + vk_param."""
    acc = C.msm(query[1:], assignment)
    res = C.jadd(initial_j, C.to_jac(query[0]))
    res = C.jadd(res, acc)
    res = C.jadd(res, C.to_jac(vk_param))
    return res


def create_proof_with_assignment(pk: ProvingKey, r: int, s: int, h, input_assignment, aux_assignment):
    """prover.rs:54-136.  Returns (A, B, C) affine plus the five raw MSM results (affine) for parity."""
    h_acc = G1.msm(pk.h_query, h)                                              # :63-66
    l_acc = G1.msm(pk.l_query, aux_assignment)                                 # :70-74
    rs_delta = G1.jmul(G1.jmul(G1.to_jac(pk.delta_g1), r), s)                  # :76-80
    assignment = list(input_assignment) + list(aux_assignment)                 # :89
    r_g1 = G1.jmul(G1.to_jac(pk.delta_g1), r)                                  # :94
    g_a = calculate_coeff(G1, r_g1, pk.a_query, pk.vk.alpha_g1, assignment)    # :96
    s_g_a = G1.jmul(g_a, s)                                                    # :98
    if r % R_MOD != 0:                                                         # :102-112
        s_g1 = G1.jmul(G1.to_jac(pk.delta_g1), s)
        g1_b = calculate_coeff(G1, s_g1, pk.b_g1_query, pk.beta_g1, assignment)
    else:
        g1_b = None
    s_g2 = G2.jmul(G2.to_jac(pk.vk.delta_g2), s)                               # :116
    g2_b = calculate_coeff(G2, s_g2, pk.b_g2_query, pk.vk.beta_g2, assignment)  # :117
    r_g1_b = G1.jmul(g1_b, r)                                                  # :118
    g_c = s_g_a                                                                # :124-128
    g_c = G1.jadd(g_c, r_g1_b)
    g_c = G1.jadd(g_c, G1.jneg(rs_delta))
    g_c = G1.jadd(g_c, l_acc)
    g_c = G1.jadd(g_c, h_acc)
    msms = {
        "h": G1.to_affine(h_acc),
        "l": G1.to_affine(l_acc),
        "a": G1.to_affine(G1.msm(pk.a_query[1:], assignment)),
        "b_g1": G1.to_affine(G1.msm(pk.b_g1_query[1:], assignment)),
        "b_g2": G2.to_affine(G2.msm(pk.b_g2_query[1:], assignment)),
    }
    return (G1.to_affine(g_a), G2.to_affine(g2_b), G1.to_affine(g_c)), msms


def create_proof_with_reduction_and_matrices(pk, r, s, m: Matrices, num_inputs, num_constraints, z,
                                             reduction: str = "libsnark"):
    """prover.rs:26-51."""
    wm = witness_map_libsnark if reduction == "libsnark" else witness_map_circom
    h = wm(m, num_inputs, num_constraints, z)
    proof, msms = create_proof_with_assignment(pk, r, s, h, z[1:num_inputs], z[num_inputs:])
    return proof, h, msms


def proof_scalars_closed_form(td: Trapdoor, qap, m: Matrices, z, h, r, s):
    """Discrete logs of (A, B, C) computed in Fr from the trapdoor -- the "in the exponent" check that
    replaces the pairing: A = alpha + sum z_i u_i(t) + r delta, B = beta + sum z_i v_i(t) + s delta,
    C = sum_{aux} z_i l_i + sum h_i hs_i + s A + r B - r s delta."""
    a, b, c, zt, n, hs, l = qap
    ni = m.num_instance_variables
    A = (td.alpha + sum(zi * ai for zi, ai in zip(z, a)) + r * td.delta) % R_MOD
    B = (td.beta + sum(zi * bi for zi, bi in zip(z, b)) + s * td.delta) % R_MOD
    C = (sum(zi * li for zi, li in zip(z[ni:], l)) + sum(hi * si for hi, si in zip(h, hs))
         + s * A + r * B - r * s % R_MOD * td.delta) % R_MOD
    return A, B, C


def verify_in_exponent(td: Trapdoor, qap, m: Matrices, z, A: int, B: int, C: int) -> bool:
    """Groth16 acceptance equation (verifier.rs:44-65) taken to discrete logs:
    A*B = alpha*beta + (sum_{i<l} z_i*abc_i)*gamma + C*delta."""
    a, b, c, zt, n, hs, l = qap
    ni = m.num_instance_variables
    gi = inv_mod(td.gamma, R_MOD)
    ic = sum(z[i] * ((td.beta * a[i] + td.alpha * b[i] + c[i]) * gi % R_MOD) for i in range(ni)) % R_MOD
    return (A * B - td.alpha * td.beta - ic * td.gamma - C * td.delta) % R_MOD == 0


# --------------------------------------------------------------------------------------------
# ark-serialize 0.4 canonical encoding (ASSUMPTION-level):
#   Fp: 32 B LE canonical.  Fq2: c0 || c1.  SW flags in the top two bits of the last byte:
#   bit7 = y is the lexicographically larger of {y,-y} ("negative"), bit6 = infinity.
#   Compressed = x with flags; uncompressed = x || y with the flags on y's last byte.
#   Fq2 ordering: by c1, then c0.
# --------------------------------------------------------------------------------------------
def _fq_is_neg(y: int) -> bool:
    return y > (Q_MOD - y) % Q_MOD


def _fq2_is_neg(y) -> bool:
    ny = Fq2.neg(y)
    return (y[1], y[0]) > (ny[1], ny[0])


def ser_g1(P, compressed: bool) -> bytes:
    if P is None:
        buf = bytearray(32 if compressed else 64)
        buf[-1] |= 0x40
        return bytes(buf)
    x, y = P
    flag = 0x80 if _fq_is_neg(y) else 0
    if compressed:
        buf = bytearray(le_bytes(x))
    else:
        buf = bytearray(le_bytes(x) + le_bytes(y))
    buf[-1] |= flag
    return bytes(buf)


def ser_g2(P, compressed: bool) -> bytes:
    if P is None:
        buf = bytearray(64 if compressed else 128)
        buf[-1] |= 0x40
        return bytes(buf)
    x, y = P
    flag = 0x80 if _fq2_is_neg(y) else 0
    if compressed:
        buf = bytearray(le_bytes(x[0]) + le_bytes(x[1]))
    else:
        buf = bytearray(le_bytes(x[0]) + le_bytes(x[1]) + le_bytes(y[0]) + le_bytes(y[1]))
    buf[-1] |= flag
    return bytes(buf)


def ser_proof(proof, compressed: bool) -> bytes:
    """Proof = a || b || c (data_structures.rs:7-14): 128 B compressed, 256 B uncompressed."""
    A, B, C = proof
    return ser_g1(A, compressed) + ser_g2(B, compressed) + ser_g1(C, compressed)


def deser_g1_uncompressed(buf: bytes):
    flags = buf[63] & 0xC0
    if flags & 0x40:
        return None
    x = int.from_bytes(buf[:32], "little")
    yb = bytearray(buf[32:64])
    yb[-1] &= 0x3F
    return (x, int.from_bytes(yb, "little"))


def deser_g2_uncompressed(buf: bytes):
    if buf[127] & 0x40:
        return None
    v = [int.from_bytes(buf[i * 32:(i + 1) * 32], "little") for i in range(3)]
    yb = bytearray(buf[96:128])
    yb[-1] &= 0x3F
    return ((v[0], v[1]), (v[2], int.from_bytes(yb, "little")))


def _ser_vec(items, f) -> bytes:
    return struct.pack("<Q", len(items)) + b"".join(f(x) for x in items)


def ser_vk(vk: VerifyingKey, compressed=False) -> bytes:
    """VerifyingKey field order incl. the fork-only delta_g1 (data_structures.rs:31-44)."""
    g1 = lambda P: ser_g1(P, compressed)
    g2 = lambda P: ser_g2(P, compressed)
    return (g1(vk.alpha_g1) + g2(vk.beta_g2) + g2(vk.gamma_g2) + g1(vk.delta_g1) + g2(vk.delta_g2)
            + _ser_vec(vk.gamma_abc_g1, g1))


def ser_pk(pk: ProvingKey, compressed=False) -> bytes:
    """ProvingKey field order (data_structures.rs:101-118); Crescent persists uncompressed
    (creds/src/utils.rs:140-197)."""
    g1 = lambda P: ser_g1(P, compressed)
    g2 = lambda P: ser_g2(P, compressed)
    return (ser_vk(pk.vk, compressed) + g1(pk.beta_g1) + g1(pk.delta_g1) + _ser_vec(pk.a_query, g1)
            + _ser_vec(pk.b_g1_query, g1) + _ser_vec(pk.b_g2_query, g2) + _ser_vec(pk.h_query, g1)
            + _ser_vec(pk.l_query, g1))


# --------------------------------------------------------------------------------------------
# iden3 .r1cs v1 (forks/circom-compat/src/circom/r1cs_reader.rs:54-256)
# --------------------------------------------------------------------------------------------
R1CS_PRIME_BYTES = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")


def read_r1cs(data: bytes):
    """Returns dict(header..., constraints=[(A,B,C)] with each a list of (wire, coeff), wire_mapping)."""
    if data[:4] != b"r1cs":
        raise ValueError("Invalid magic number")
    version, nsec = struct.unpack_from("<II", data, 4)
    if version != 1:
        raise ValueError("Unsupported version")
    off = 12
    secs = {}
    for _ in range(nsec):
        ty, sz = struct.unpack_from("<IQ", data, off)
        off += 12
        secs[ty] = (off, sz)
        off += sz
    ho, hs = secs[1]
    field_size = struct.unpack_from("<I", data, ho)[0]
    if field_size != 32:
        raise ValueError("This parser only supports 32-byte fields")
    if hs != 32 + field_size:
        raise ValueError("Invalid header section size")
    prime = data[ho + 4:ho + 36]
    if prime != R1CS_PRIME_BYTES:
        raise ValueError("This parser only supports bn256")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", data, ho + 36)
    co, _ = secs[2]
    p = co
    cons = []
    for _ in range(n_constraints):
        row = []
        for _k in range(3):
            nv = struct.unpack_from("<I", data, p)[0]
            p += 4
            vec = []
            for _j in range(nv):
                w = struct.unpack_from("<I", data, p)[0]
                v = int.from_bytes(data[p + 4:p + 36], "little")
                p += 36
                vec.append((w, v))
            row.append(vec)
        cons.append(tuple(row))
    mo, ms = secs[3]
    if ms != n_wires * 8:
        raise ValueError("Invalid map section size")
    wmap = list(struct.unpack_from("<%dQ" % n_wires, data, mo))
    if wmap[0] != 0:
        raise ValueError("Wire 0 should always be mapped to 0")
    return dict(version=version, field_size=field_size, prime=prime, n_wires=n_wires, n_pub_out=n_pub_out,
                n_pub_in=n_pub_in, n_prv_in=n_prv_in, n_labels=n_labels, n_constraints=n_constraints,
                constraints=cons, wire_mapping=wmap)


def write_r1cs(n_wires, n_pub_out, n_pub_in, n_prv_in, constraints, n_labels=None, wire_mapping=None) -> bytes:
    n_labels = n_wires if n_labels is None else n_labels
    wire_mapping = list(range(n_wires)) if wire_mapping is None else wire_mapping
    hdr = struct.pack("<I", 32) + R1CS_PRIME_BYTES + struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in,
                                                                 n_prv_in, n_labels, len(constraints))
    body = bytearray()
    for row in constraints:
        for vec in row:
            body += struct.pack("<I", len(vec))
            for w, v in vec:
                body += struct.pack("<I", w) + le_bytes(v % R_MOD)
    mp = struct.pack("<%dQ" % n_wires, *wire_mapping)
    out = b"r1cs" + struct.pack("<II", 1, 3)
    for ty, sec in ((1, hdr), (2, bytes(body)), (3, mp)):
        out += struct.pack("<IQ", ty, len(sec)) + sec
    return out


def r1cs_to_matrices(r) -> Matrices:
    """Variable order of CircomCircuit::generate_constraints (forks/circom-compat/src/circom/circuit.rs:28-87)
    with wire_mapping forced to None (builder.rs:64): column == circom wire index.  ark-relations'
    to_matrices() sums duplicate-variable terms of one LC and drops zero coefficients; do the same."""
    ni = 1 + r["n_pub_in"] + r["n_pub_out"]                      # r1cs_reader.rs:27
    rows = {0: [], 1: [], 2: []}
    for con in r["constraints"]:
        for k in range(3):
            acc = {}
            for w, v in con[k]:
                acc[w] = (acc.get(w, 0) + v) % R_MOD
            rows[k].append([(v, w) for w, v in sorted(acc.items()) if v != 0])
    return Matrices(ni, r["n_wires"] - ni, rows[0], rows[1], rows[2])


# --------------------------------------------------------------------------------------------
# The reference's own test circuits, restated as matrices
# --------------------------------------------------------------------------------------------
def my_silly_circuit(a: int, b: int):
    """forks/groth16/src/test.rs:14-43: witness a, b; input c = a*b; six copies of a*b=c.
    Variable order: instance [1, c], witness [a, b] -> columns: 0=one, 1=c, 2=a, 3=b."""
    rows_a = [[(1, 2)] for _ in range(6)]
    rows_b = [[(1, 3)] for _ in range(6)]
    rows_c = [[(1, 1)] for _ in range(6)]
    z = [1, a * b % R_MOD, a % R_MOD, b % R_MOD]
    return Matrices(2, 2, rows_a, rows_b, rows_c), z


def dummy_circuit(a: int, b: int, num_variables=(1 << 10) - 100, num_constraints=(1 << 10) - 100, num_inputs=5):
    """creds/src/rangeproof.rs:446-486: witness a, b first; input c=a*b; (num_inputs-1) more inputs = a;
    (num_variables-num_inputs-2) more witnesses = a; (num_constraints-1) x [a*b=c]; one empty constraint.
    Instance = [1, c, a, a, a, a] (ni = num_inputs+1), witness = [a, b, a, a, ...]."""
    ni = num_inputs + 1
    nw = 2 + (num_variables - num_inputs - 2)
    col_a, col_b, col_c = ni + 0, ni + 1, 1
    rows_a = [[(1, col_a)] for _ in range(num_constraints - 1)] + [[]]
    rows_b = [[(1, col_b)] for _ in range(num_constraints - 1)] + [[]]
    rows_c = [[(1, col_c)] for _ in range(num_constraints - 1)] + [[]]
    z = [1, a * b % R_MOD] + [a % R_MOD] * (num_inputs - 1) + [a % R_MOD, b % R_MOD] + [a % R_MOD] * (nw - 2)
    return Matrices(ni, nw, rows_a, rows_b, rows_c), z


# --------------------------------------------------------------------------------------------
# Deterministic data streams shared with the C oracle, the tests and the bench:
# SplitMix64 from seed ^ index.
# --------------------------------------------------------------------------------------------
def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def stream_fr(seed: int, index: int) -> int:
    """Uniform-ish Fr element: 4 SplitMix64 words, top limb masked to 254 bits, reduced mod r."""
    v = 0
    for k in range(4):
        v |= splitmix64(seed ^ (index * 4 + k)) << (64 * k)
    v &= (1 << 254) - 1
    return v % R_MOD


def random_satisfiable_r1cs(seed: int, nc: int, ni: int, nw: int, max_row: int = 4):
    """Small random R1CS + satisfying witness: A, B rows random; C row = random terms plus a term on the
    constant wire 0 solved so that <C,z> = <A,z><B,z> (the construction the bench's S-rs256 twin uses)."""
    m_wires = ni + nw
    z = [1] + [stream_fr(seed ^ 0x5A5A, i) if (i % 3) else (stream_fr(seed ^ 0x5A5A, i) & 0xFF) for i in range(1, m_wires)]
    ctr = [0]

    def rnd():
        ctr[0] += 1
        return splitmix64(seed ^ (0xABCD << 20) ^ ctr[0])

    def row(allow_empty=True):
        k = rnd() % (max_row + 1)
        if k == 0 and not allow_empty:
            k = 1
        cols = sorted({rnd() % m_wires for _ in range(k)})
        out = []
        for c in cols:
            sel = rnd() % 10
            if sel < 6:
                v = 1
            elif sel < 7:
                v = R_MOD - 1
            elif sel < 9:
                v = pow(2, rnd() % 121, R_MOD)
            else:
                v = stream_fr(seed ^ 0x77, rnd() % (1 << 30))
            out.append((v, c))
        return out

    A, B, C = [], [], []
    for _ in range(nc):
        ra, rb = row(False), row(False)
        rc = [t for t in row() if t[1] != 0]
        tgt = evaluate_constraint(ra, z) * evaluate_constraint(rb, z) % R_MOD
        k0 = (tgt - evaluate_constraint(rc, z)) % R_MOD
        if k0:
            rc = [(k0, 0)] + rc
        A.append(ra)
        B.append(rb)
        C.append(rc)
    return Matrices(ni, nw, A, B, C), z
