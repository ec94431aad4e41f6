// pairing.cuh -- the BN254 optimal-ate pairing for the Groth16 verifier kernels (row f-4 of the scope table).
//
// Replaces what the reference's verifier gets from ark-ec 0.4 `models::bn` + ark-bn254 0.4 (third-party to the reference
// tree) at forks/groth16/src/verifier.rs:13-20 (E::pairing in prepare_verifying_key) and :44-65 (E::multi_miller_loop +
// E::final_exponentiation in verify_proof_with_prepared_inputs).  GT elements come out bit-identical to that library's:
// same tower (Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 9 + u), same Miller function (signed digits of 6x+2,
// homogeneous projective line steps, the two Frobenius additions), same final power 2x(6x^2+3x+1)(q^4-q^2+1)/r of the
// hard part.  Field elements are unique, so the formulas below are free to differ from arkworks' (Karatsuba over Fq6, complex
// squaring, Granger-Scott squarings in the cyclotomic subgroup) as long as the VALUES agree -- oracle/pairing.py checks them.
//
// Everything is __host__ __device__: tests/host_pairing_shim.cpp compiles this file with gcc and the CPU suite runs the exact
// control flow of the kernels against the big-integer oracle.
#pragma once
#include <stddef.h>

#include "ec.cuh"

#define G16_PAIRING_SCALARS
#include "pairing_consts.inc"
#undef G16_PAIRING_SCALARS

namespace g16 {

// ---- constants ---------------------------------------------------------------------------------------------------------------
enum {
    PC_FROB = 0,      // FROB[n][k] = xi^(k (q^n - 1) / 6) at PC_FROB + ((n - 1) * 5 + (k - 1)) * 2, n = 1..3, k = 1..5
    PC_TWIST_X = 30,  // xi^((q - 1) / 3)
    PC_TWIST_Y = 32,  // xi^((q - 1) / 2)
    PC_G2_B = 34,     // 3 / xi
    PC_TWO_INV = 36,  // 1 / 2
};
static_assert(G16_PC_COUNT == 37, "pairing_consts.inc out of date");
static const uint32_t kPairingConstHost[G16_PC_COUNT][8] =
#include "pairing_consts.inc"
    ;
#ifdef __CUDACC__
static __device__ const uint32_t kPairingConstDev[G16_PC_COUNT][8] =
#include "pairing_consts.inc"
    ;
#endif
constexpr int kEllCoeffs = 64 + 25 + 2;  // line coefficients of one G2 point: 64 doublings, 25 non-zero digits, Q1, Q2

G16_HD Fq pc_fq(int i) {
#ifdef __CUDA_ARCH__
    const uint32_t* t = kPairingConstDev[i];
#else
    const uint32_t* t = kPairingConstHost[i];
#endif
    Fq r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = t[j];
    return r;
}
G16_HD Fq2 pc_fq2(int i) { return Fq2{pc_fq(i), pc_fq(i + 1)}; }

// Code-size switch for the verifier kernels (ncu: stall_no_instruction 1.8 per issue at saturation): with G16_PAIRING_COMPACT the
// small helpers that are inlined at many call sites (the xi multiplication: 10 field additions; the line evaluation: 4 products)
// become single functions: 147 k -> 101 k SASS instructions in verify.o, but MEASURED SLOWER (741 k vs 809 k proofs/s at 65 536
// proofs; single proof 22.6 vs 22.5 ms) -- ptxas schedules better across the inlined helpers than the fetch stalls cost.  Off.
#ifdef G16_PAIRING_COMPACT
#define G16_PAIRING_HELPER G16_HD_NOINLINE
#else
#define G16_PAIRING_HELPER G16_HD
#endif

// ---- Fq2 helpers -------------------------------------------------------------------------------------------------------------
G16_PAIRING_HELPER Fq2 fq2_mul_xi(const Fq2& a) {  // (9 + u)(a0 + a1 u) = (9 a0 - a1) + (9 a1 + a0) u
    Fq t0 = a.c0.dbl().dbl().dbl() + a.c0;
    Fq t1 = a.c1.dbl().dbl().dbl() + a.c1;
    return Fq2{t0 - a.c1, t1 + a.c0};
}
G16_HD Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, a.c1.neg()}; }
G16_HD Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) { return Fq2{a.c0 * k, a.c1 * k}; }

// ---- Fq6 = Fq2[v]/(v^3 - xi) ------------------------------------------------------------------------------------------------
struct Fq6 {
    Fq2 c0, c1, c2;
    static G16_HD Fq6 zero() { return Fq6{Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static G16_HD Fq6 one() { return Fq6{Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    G16_HD bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
    G16_HD bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    friend G16_HD Fq6 operator+(const Fq6& a, const Fq6& b) { return Fq6{a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
    friend G16_HD Fq6 operator-(const Fq6& a, const Fq6& b) { return Fq6{a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
    G16_HD Fq6 neg() const { return Fq6{c0.neg(), c1.neg(), c2.neg()}; }
    G16_HD Fq6 dbl() const { return Fq6{c0.dbl(), c1.dbl(), c2.dbl()}; }
    G16_HD Fq6 mul_by_v() const { return Fq6{fq2_mul_xi(c2), c0, c1}; }
    G16_HD Fq6 mul_by_fq2(const Fq2& k) const { return Fq6{c0 * k, c1 * k, c2 * k}; }
    // Karatsuba: 6 Fq2 products
    friend G16_HD_NOINLINE Fq6 operator*(const Fq6& a, const Fq6& b) {
        Fq2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
        Fq2 t0 = (a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2;
        Fq2 t1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1;
        Fq2 t2 = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2;
        return Fq6{v0 + fq2_mul_xi(t0), t1 + fq2_mul_xi(v2), t2 + v1};
    }
    // product with the sparse element d0 + d1 v: 5 Fq2 products
    G16_HD_NOINLINE Fq6 mul_by_01(const Fq2& d0, const Fq2& d1) const {
        Fq2 v0 = c0 * d0, v1 = c1 * d1;
        Fq2 t0 = (c1 + c2) * d1 - v1;
        Fq2 t1 = (c0 + c1) * (d0 + d1) - v0 - v1;
        Fq2 t2 = (c0 + c2) * d0 - v0 + v1;
        return Fq6{v0 + fq2_mul_xi(t0), t1, t2};
    }
    G16_HD_NOINLINE Fq6 inverse() const {
        Fq2 t0 = c0.sqr() - fq2_mul_xi(c1 * c2);
        Fq2 t1 = fq2_mul_xi(c2.sqr()) - c0 * c1;
        Fq2 t2 = c1.sqr() - c0 * c2;
        Fq2 d = c0 * t0 + fq2_mul_xi(c2 * t1 + c1 * t2);
        Fq2 di = d.inverse_fast();
        return Fq6{t0 * di, t1 * di, t2 * di};
    }
};

// ---- Fq12 = Fq6[w]/(w^2 - v) -------------------------------------------------------------------------------------------------
// ark-serialize order = memory order here: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (each an Fq2 = c0, c1): 48 x u64.
struct Fq12 {
    Fq6 c0, c1;
    static G16_HD Fq12 one() { return Fq12{Fq6::one(), Fq6::zero()}; }
    G16_HD bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
    G16_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    G16_HD Fq12 conj() const { return Fq12{c0, c1.neg()}; }  // x^(q^6); the inverse on the cyclotomic subgroup
    // Karatsuba: 3 Fq6 products
    friend G16_HD_NOINLINE Fq12 operator*(const Fq12& a, const Fq12& b) {
        Fq6 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;
        Fq6 t = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1;
        return Fq12{v0 + v1.mul_by_v(), t};
    }
    // complex squaring: 2 Fq6 products
    G16_HD_NOINLINE Fq12 sqr() const {
        Fq6 ab = c0 * c1;
        Fq6 t = (c0 + c1) * (c0 + c1.mul_by_v()) - ab - ab.mul_by_v();
        return Fq12{t, ab.dbl()};
    }
    G16_HD_NOINLINE Fq12 inverse() const {
        Fq6 n = (c0 * c0 - (c1 * c1).mul_by_v()).inverse();
        return Fq12{c0 * n, (c1 * n).neg()};
    }
    // x^(q^n), n = 1..3: as an element of Fq2[w]/(w^6 - xi) coefficient k (of w^k; tower slot (i, j) holds k = 2i + j) maps to
    // conj^n(coefficient) * xi^(k (q^n - 1)/6)
    G16_HD_NOINLINE Fq12 frobenius(int n) const {
        const int base = PC_FROB + (n - 1) * 10;
        const bool cj = n & 1;
        auto m = [&](const Fq2& a, int k) {
            Fq2 c = cj ? fq2_conj(a) : a;
            return k == 0 ? c : c * pc_fq2(base + (k - 1) * 2);
        };
        return Fq12{Fq6{m(c0.c0, 0), m(c0.c1, 2), m(c0.c2, 4)}, Fq6{m(c1.c0, 1), m(c1.c1, 3), m(c1.c2, 5)}};
    }
    // this * (e0 + (d0 + d1 v) w): the line of a D-type twist ("mul_by_034"): 13 Fq2 products
    G16_HD_NOINLINE void mul_by_034(const Fq2& e0, const Fq2& d0, const Fq2& d1) {
        Fq6 a = c0.mul_by_fq2(e0);
        Fq6 b = c1.mul_by_01(d0, d1);
        Fq6 e = (c0 + c1).mul_by_01(e0 + d0, d1);
        c1 = e - (a + b);
        c0 = a + b.mul_by_v();
    }
    // squaring of an element of the cyclotomic subgroup (Granger & Scott, "Faster squaring in the cyclotomic subgroup of sixth
    // degree extensions", 3.2): three Fq4 squarings = 6 Fq2 products instead of 12
    G16_HD_NOINLINE Fq12 cyclotomic_sqr() const {
        auto fp4_sqr = [](const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {  // (a + b y)^2 with y^2 = xi
            Fq2 ab = a * b;
            r0 = (a + b) * (fq2_mul_xi(b) + a) - ab - fq2_mul_xi(ab);
            r1 = ab.dbl();
        };
        Fq2 t0, t1, t2, t3, t4, t5;
        fp4_sqr(c0.c0, c1.c1, t0, t1);
        fp4_sqr(c1.c0, c0.c2, t2, t3);
        fp4_sqr(c0.c1, c1.c2, t4, t5);
        auto m3s2 = [](const Fq2& t, const Fq2& z) {  // 3t - 2z
            Fq2 d = t - z;
            return d.dbl() + t;
        };
        auto m3a2 = [](const Fq2& t, const Fq2& z) {  // 3t + 2z
            Fq2 d = t + z;
            return d.dbl() + t;
        };
        Fq12 r;
        r.c0.c0 = m3s2(t0, c0.c0);
        r.c1.c1 = m3a2(t1, c1.c1);
        r.c1.c0 = m3a2(fq2_mul_xi(t5), c1.c0);
        r.c0.c2 = m3s2(t4, c0.c2);
        r.c0.c1 = m3s2(t2, c0.c1);
        r.c1.c2 = m3a2(t3, c1.c2);
        return r;
    }
};

// ---- G2 line steps (ark-ec models/bn/g2.rs: G2HomProjective::double_in_place / add_in_place, TwistType::D) -------------------
struct EllCoeff {
    Fq2 c0, c1, c2;
};
struct G2Hom {
    Fq2 x, y, z;
};

G16_HD_NOINLINE EllCoeff g2_double_step(G2Hom& r) {
    const Fq two_inv = pc_fq(PC_TWO_INV);
    Fq2 a = fq2_mul_fq(r.x * r.y, two_inv);
    Fq2 b = r.y.sqr();
    Fq2 c = r.z.sqr();
    Fq2 e = pc_fq2(PC_G2_B) * (c.dbl() + c);
    Fq2 f = e.dbl() + e;
    Fq2 g = fq2_mul_fq(b + f, two_inv);
    Fq2 h = (r.y + r.z).sqr() - (b + c);
    Fq2 i = e - b;
    Fq2 j = r.x.sqr();
    Fq2 e2 = e.sqr();
    r.x = a * (b - f);
    r.y = g.sqr() - (e2.dbl() + e2);
    r.z = b * h;
    return EllCoeff{h.neg(), j.dbl() + j, i};
}

G16_HD_NOINLINE EllCoeff g2_add_step(G2Hom& r, const G2Affine& q) {
    Fq2 theta = r.y - q.y * r.z;
    Fq2 lambda = r.x - q.x * r.z;
    Fq2 c = theta.sqr();
    Fq2 d = lambda.sqr();
    Fq2 e = lambda * d;
    Fq2 f = r.z * c;
    Fq2 g = r.x * d;
    Fq2 h = e + f - g.dbl();
    r.x = lambda * h;
    r.y = theta * (g - h) - e * r.y;
    r.z = r.z * e;
    Fq2 j = theta * q.x - lambda * q.y;
    return EllCoeff{lambda, theta.neg(), j};
}

G16_HD G2Affine g2_mul_by_char(const G2Affine& q) {
    return G2Affine{fq2_conj(q.x) * pc_fq2(PC_TWIST_X), fq2_conj(q.y) * pc_fq2(PC_TWIST_Y)};
}

// f *= line(P)   (Bn::ell, TwistType::D)
G16_PAIRING_HELPER void ell(Fq12& f, const EllCoeff& c, const G1Affine& p) { f.mul_by_034(fq2_mul_fq(c.c0, p.y), fq2_mul_fq(c.c1, p.x), c.c2); }

G16_HD int ate_digit(int i) { return (int)((G16_ATE_POS >> i) & 1ull) - (int)((G16_ATE_NEG >> i) & 1ull); }

// G2Prepared::from: the kEllCoeffs line coefficients of a fixed G2 point (not the point at infinity)
G16_HD_NOINLINE void g2_prepare(const G2Affine& q, EllCoeff* out) {
    G2Hom r{q.x, q.y, Fq2::one()};
    const G2Affine nq = q.neg();
    int idx = 0;
    for (int i = 63; i >= 0; i--) {
        out[idx++] = g2_double_step(r);
        int d = ate_digit(i);
        if (d != 0) out[idx++] = g2_add_step(r, d > 0 ? q : nq);
    }
    G2Affine q1 = g2_mul_by_char(q);
    G2Affine q2 = g2_mul_by_char(q1).neg();
    out[idx++] = g2_add_step(r, q1);
    out[idx++] = g2_add_step(r, q2);
}

// Bn::multi_miller_loop over up to three pairs: pair 0 computes the lines of its G2 point on the fly (a proof's B differs
// per proof), pairs 1 and 2 read prepared coefficient tables (the verifying key's -gamma and -delta).  act[j] = false skips
// pair j (the reference filters out pairs holding a point at infinity).
G16_HD_NOINLINE Fq12 miller_loop3(const G1Affine* p, const bool* act, const G2Affine& q0, const EllCoeff* t1, const EllCoeff* t2) {
    Fq12 f = Fq12::one();
    G2Hom r{q0.x, q0.y, Fq2::one()};
    const G2Affine nq0 = q0.neg();
    int idx = 0;
    for (int i = 63; i >= 0; i--) {
        if (i != 63) f = f.sqr();
        if (act[0]) ell(f, g2_double_step(r), p[0]);
        if (act[1]) ell(f, t1[idx], p[1]);
        if (act[2]) ell(f, t2[idx], p[2]);
        idx++;
        int d = ate_digit(i);
        if (d != 0) {
            if (act[0]) ell(f, g2_add_step(r, d > 0 ? q0 : nq0), p[0]);
            if (act[1]) ell(f, t1[idx], p[1]);
            if (act[2]) ell(f, t2[idx], p[2]);
            idx++;
        }
    }
    if (act[0]) {
        G2Affine q1 = g2_mul_by_char(q0);
        G2Affine q2 = g2_mul_by_char(q1).neg();
        ell(f, g2_add_step(r, q1), p[0]);
        ell(f, g2_add_step(r, q2), p[0]);
    }
    for (int k = 0; k < 2; k++) {
        if (act[1]) ell(f, t1[idx + k], p[1]);
        if (act[2]) ell(f, t2[idx + k], p[2]);
    }
    return f;
}

// f^x for f in the cyclotomic subgroup, x the BN parameter (63 bits)
G16_HD_NOINLINE Fq12 cyclotomic_exp_x(const Fq12& f) {
    Fq12 r = f;  // leading bit (bit 62)
    for (int i = 61; i >= 0; i--) {
        r = r.cyclotomic_sqr();
        if ((G16_BN_X >> i) & 1ull) r = r * f;
    }
    return r;
}
G16_HD Fq12 exp_by_neg_x(const Fq12& f) { return cyclotomic_exp_x(f).conj(); }

// Bn::final_exponentiation.  ok = false when f is zero (the reference returns None -> SynthesisError::UnexpectedIdentity).
G16_HD_NOINLINE Fq12 final_exponentiation(const Fq12& f, bool& ok) {
    ok = !f.is_zero();
    if (!ok) return Fq12::one();
    // easy part: f^((q^6 - 1)(q^2 + 1))
    Fq12 r = f.conj() * f.inverse();
    r = r.frobenius(2) * r;
    // hard part (Fuentes-Castaneda, Knapp, Rodriguez-Henriquez): r^(2x(6x^2 + 3x + 1)(q^4 - q^2 + 1)/r)
    Fq12 y0 = exp_by_neg_x(r);
    Fq12 y1 = y0.cyclotomic_sqr();
    Fq12 y2 = y1.cyclotomic_sqr();
    Fq12 y3 = y2 * y1;
    Fq12 y4 = exp_by_neg_x(y3);
    Fq12 y5 = y4.cyclotomic_sqr();
    Fq12 y6 = exp_by_neg_x(y5).conj();
    y3 = y3.conj();
    Fq12 y7 = y6 * y4;
    Fq12 y8 = y7 * y3;
    Fq12 y9 = y8 * y1;
    Fq12 y10 = y8 * y4;
    Fq12 y11 = y10 * r;
    Fq12 y13 = y9.frobenius(1) * y11;
    Fq12 y14 = y8.frobenius(2) * y13;
    Fq12 y15 = (r.conj() * y9).frobenius(3);
    return y15 * y14;
}

}  // namespace g16

// ---- the verifier's per-proof work (forks/groth16/src/verifier.rs) ---------------------------------------------------------------
namespace g16 {

constexpr int kAbcWindowBits = 8;                       // fixed-base windows over the verifying key's gamma_abc_g1 points
constexpr int kAbcWindows = 256 / kAbcWindowBits;       // 32
constexpr int kAbcDigits = (1 << kAbcWindowBits) - 1;   // 255 table entries per window: d * 2^(8w) * P, d = 1..255

// prepare_inputs (verifier.rs:25-39): gamma_abc_g1[0] + sum_i x_i * gamma_abc_g1[i + 1], affine.  The scalar
// multiplications read the window tables built once per verifying key: tbl[(i * 32 + w) * 255 + d - 1] = d * 2^(8w) * abc[i + 1].
G16_HD_NOINLINE G1Affine prepare_inputs_one(const G1Affine& abc0, const G1Affine* tbl, const Fr* inputs, size_t n_inputs) {
    G1XYZZ acc = G1XYZZ::from_affine(abc0);
    for (size_t i = 0; i < n_inputs; i++) {
        const Fr k = inputs[i].from_mont();  // into_bigint
        const G1Affine* row = tbl + i * (size_t)(kAbcWindows * kAbcDigits);
        for (int w = 0; w < kAbcWindows; w++) {
            uint32_t d = (k.v[w >> 2] >> (8 * (w & 3))) & 0xffu;
            if (d) acc.madd(row[w * kAbcDigits + (d - 1)]);
        }
    }
    return acc.to_affine();
}

// verify_proof_with_prepared_inputs (verifier.rs:44-65): e(A, B) * e(prepared, -gamma) * e(C, -delta) == e(alpha, beta).
// Returns 1 (accepted), 0 (rejected) or 2 (final_exponentiation returned None: SynthesisError::UnexpectedIdentity).
// g2_inf: bit 1 / bit 2 set when the key's gamma_g2 / delta_g2 is the point at infinity (that pair is then filtered out).
G16_HD_NOINLINE int verify_one(const G1Affine& a, const G2Affine& b, const G1Affine& c, const G1Affine& prepared,
                               const EllCoeff* neg_gamma, const EllCoeff* neg_delta, const Fq12& alpha_beta, unsigned g2_inf = 0) {
    G1Affine p[3] = {a, prepared, c};
    bool act[3] = {!a.is_inf() && !b.is_inf(), !prepared.is_inf() && !(g2_inf & 2u), !c.is_inf() && !(g2_inf & 4u)};
    Fq12 f = miller_loop3(p, act, b, neg_gamma, neg_delta);
    bool ok;
    Fq12 t = final_exponentiation(f, ok);
    if (!ok) return 2;
    return t == alpha_beta ? 1 : 0;
}

}  // namespace g16
