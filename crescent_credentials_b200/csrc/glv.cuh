// glv.cuh -- the GLV split of a BN254 G1 scalar multiplication (Gallant-Lambert-Vanstone) for the ONE place where a fresh-point
// multiplication is a lone lane's critical path: k_scale_point (s * MSM_a, r * MSM_b1 of forks/groth16/src/prover.rs:98,118).
//
// G1: y^2 = x^3 + 3 has the endomorphism phi(x, y) = (beta x, y) = lambda * (x, y) with beta^3 = 1 in Fq, lambda^3 = 1 in Fr.
// k = k1 + k2 * lambda (mod r) with |k1|, |k2| < 2^128, so k P = k1 P + k2 phi(P): two half-length multiplications that run on
// two warps at once -- half the doublings on the critical path.  The result is the same group element as k P (exact arithmetic),
// so nothing downstream can tell.  Constants derived and checked by tests/test_abi.py (lattice basis from the extended Euclidean
// algorithm on (r, lambda); 2 * 10^5 random scalars decompose with |k1|, |k2| < 2^127).
#pragma once
#include "ec.cuh"

namespace g16 {

struct GlvHalf {
    uint32_t k[5];  // magnitude, little-endian (< 2^160; in practice < 2^128)
    int neg;        // the half enters with a minus sign
};
struct GlvScalar {
    GlvHalf h[2];  // k = (+-h[0]) + (+-h[1]) * lambda  (mod r)
    int ok;        // 0: magnitudes out of range (never seen): the caller falls back to the plain multiplication
    int pad;
};
constexpr int kGlvNibbles = 33;  // 132 bits

// beta in Montgomery form (beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe)
G16_HD Fq glv_beta() {
    Fq b;
    const uint32_t w[8] = {0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu, 0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du};
#pragma unroll
    for (int i = 0; i < 8; i++) b.v[i] = w[i];
    return b;
}
// phi(P) in XYZZ coordinates: x = X / ZZ, so only X is scaled
G16_HD G1XYZZ glv_endo(const G1XYZZ& p) {
    G1XYZZ r = p;
    if (!p.is_inf()) r.x = p.x * glv_beta();
    return r;
}

// ---- host: decomposition of a canonical scalar (8 x u32 little-endian, < r) -----------------------------------------------------
// lattice basis v1 = (a1, b1), v2 = (a2, b2) of {(x, y): x + y lambda = 0 mod r} with a1 b2 - a2 b1 = r:
//   a1 = b2 = 0x89d3256894d213e3,  b1 = -0x6f4d8248eeb859fc8211bbeb7d4f1128,  a2 = 0x6f4d8248eeb859fd0be4e1541221250b
// c1 = round(b2 k / r), c2 = round(-b1 k / r) through g_i = round(2^380 * . / r);  k1 = k - c1 a1 - c2 a2,  k2 = c1 |b1| - c2 b2.
namespace glv_detail {
typedef unsigned __int128 u128;
struct U256 {
    uint64_t l[4];
};
// (a * b) >> 380 for a, b < 2^256 (the product has up to 512 bits; the result fits 132 bits)
inline U256 mul_shr380(const U256& a, const U256& b) {
    uint64_t p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.l[i] * b.l[j] + p[i + j];
            p[i + j] = (uint64_t)c;
            c >>= 64;
        }
        p[i + 4] = (uint64_t)c;
    }
    // 380 = 5 * 64 + 60
    U256 r;
    r.l[0] = (p[5] >> 60) | (p[6] << 4);
    r.l[1] = (p[6] >> 60) | (p[7] << 4);
    r.l[2] = p[7] >> 60;
    r.l[3] = 0;
    return r;
}
inline U256 mul_lo(const U256& a, const U256& b) {  // a * b mod 2^256
    uint64_t p[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; i + j < 4; j++) {
            c += (u128)a.l[i] * b.l[j] + p[i + j];
            p[i + j] = (uint64_t)c;
            c >>= 64;
        }
    }
    return U256{{p[0], p[1], p[2], p[3]}};
}
inline U256 sub(const U256& a, const U256& b) {  // mod 2^256
    U256 r;
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)bw;
        r.l[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
    return r;
}
inline GlvHalf half_of(const U256& x) {  // two's complement -> sign + magnitude
    GlvHalf h;
    U256 m = x;
    h.neg = (int)(x.l[3] >> 63);
    if (h.neg) m = sub(U256{{0, 0, 0, 0}}, x);
    h.k[0] = (uint32_t)m.l[0];
    h.k[1] = (uint32_t)(m.l[0] >> 32);
    h.k[2] = (uint32_t)m.l[1];
    h.k[3] = (uint32_t)(m.l[1] >> 32);
    h.k[4] = (uint32_t)m.l[2];
    return h;
}
inline bool fits(const U256& x) {  // |x| < 2^131
    U256 m = (x.l[3] >> 63) ? sub(U256{{0, 0, 0, 0}}, x) : x;
    return m.l[3] == 0 && (m.l[2] >> 3) == 0;
}
}  // namespace glv_detail

inline GlvScalar glv_decompose(const uint32_t canon[8]) {
    using namespace glv_detail;
    const U256 g1{{0x28fa7d32d2fafba6ull, 0x76eb9c714773a6efull, 0x2d91d232ec7e0b3dull, 0}};
    const U256 g2{{0x9869375169b9be00ull, 0xda5e38cfb5eaa26dull, 0xf7a7bd9d4391eb18ull, 0x24ccef014a773d2cull}};
    const U256 a1{{0x89d3256894d213e3ull, 0, 0, 0}};
    const U256 b1m{{0x8211bbeb7d4f1128ull, 0x6f4d8248eeb859fcull, 0, 0}};  // |b1|
    const U256 a2{{0x0be4e1541221250bull, 0x6f4d8248eeb859fdull, 0, 0}};
    const U256 b2 = a1;
    U256 k;
    for (int i = 0; i < 4; i++) k.l[i] = (uint64_t)canon[2 * i] | ((uint64_t)canon[2 * i + 1] << 32);
    const U256 c1 = mul_shr380(k, g1), c2 = mul_shr380(k, g2);
    const U256 k1 = sub(sub(k, mul_lo(c1, a1)), mul_lo(c2, a2));
    const U256 k2 = sub(mul_lo(c1, b1m), mul_lo(c2, b2));
    GlvScalar s;
    s.h[0] = half_of(k1);
    s.h[1] = half_of(k2);
    s.ok = fits(k1) && fits(k2);
    s.pad = 0;
    return s;
}

// one half of the split product: (+-k_half) * P or (+-k_half) * phi(P)
G16_HD G1XYZZ glv_half_mul(const G1XYZZ& p, const GlvHalf& h, int which) {
    G1XYZZ base = which ? glv_endo(p) : p;
    if (h.neg) base = base.neg();
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = i < 5 ? h.k[i] : 0u;
    return scalar_mul_window(base, k, kGlvNibbles);
}

}  // namespace g16
