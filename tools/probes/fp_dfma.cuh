// fp_dfma.cuh -- Montgomery product on the FP64 pipe (experimental second multiplier for Fp<>).
//
// B200 keeps a full-rate FP64 pipe (64 DFMA / clk / SM) next to the integer multiplier, whose IMAD.WIDE rate bounds the
// 8 x 32-bit CIOS product of fp.cuh at ~65 G mul/s.  A DFMA delivers half of a 52 x 52-bit product: with round-toward-zero
//     hi = fma_rz(a, b, 2^104)            -> mantissa = floor(a*b / 2^52)
//     lo = fma   (a, b, (2^104 + 2^52) - hi) -> mantissa = a*b mod 2^52          (exact)
// (Emmart, Zheng, Weems: "Faster modular exponentiation using double precision floating point arithmetic on the GPU",
// ARITH 2018.)  The raw bit patterns of hi / lo are summed as 64-bit integers into column accumulators; the exponent
// biases are subtracted once per column.  Operands are split into five 52-bit limbs; the reduction runs four 52-bit
// Montgomery rounds and one 48-bit round so that R stays 2^256 and the result is the same bit pattern the integer product
// returns (arkworks' Montgomery form).  25 + 25 limb products = 100 DFMA + 65 DADD per field product instead of 120
// IMAD.WIDE.
//
// The same source runs on the host (tests/host_fp_shim.cpp) with std::fma under FE_TOWARDZERO.
#pragma once
#include "../../crescent_credentials_b200/csrc/fp.cuh"
#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace g16 {

struct dfma {
    static constexpr uint64_t M52 = (1ull << 52) - 1;
    static constexpr uint64_t M48 = (1ull << 48) - 1;
    static constexpr uint64_t EXP52 = 0x4330000000000000ull;   // bits of 2^52
    static constexpr uint64_t EXP104 = 0x4670000000000000ull;  // bits of 2^104

    static G16_HD double from_bits(uint64_t b) {
#ifdef __CUDA_ARCH__
        return __longlong_as_double((long long)b);
#else
        double d;
        memcpy(&d, &b, 8);
        return d;
#endif
    }
    static G16_HD uint64_t to_bits(double d) {
#ifdef __CUDA_ARCH__
        return (uint64_t)__double_as_longlong(d);
#else
        uint64_t b;
        memcpy(&b, &d, 8);
        return b;
#endif
    }
    // integer < 2^52 -> the same value as a double
    static G16_HD double to_double(uint64_t x) { return from_bits(x | EXP52) - from_bits(EXP52); }

    // bits(hi) = EXP104 + floor(a*b / 2^52), bits(lo) = EXP52 + (a*b mod 2^52)
    static G16_HD void mul_hi_lo(double a, double b, uint64_t& hi, uint64_t& lo) {
        const double c1 = from_bits(EXP104);
        const double c2 = from_bits(EXP104) + from_bits(EXP52);  // 2^104 + 2^52: exact
#ifdef __CUDA_ARCH__
        double h = __fma_rz(a, b, c1);
        double l = __fma_rz(a, b, c2 - h);
#else
        double h = std::fma(a, b, c1);  // caller runs under FE_TOWARDZERO
        double l = std::fma(a, b, c2 - h);
#endif
        hi = to_bits(h);
        lo = to_bits(l);
    }

    // 8 x u32 -> 5 x 52-bit limbs (top limb: 48 bits)
    static G16_HD void split52(const uint32_t* v, uint64_t* l) {
        uint64_t w0 = (uint64_t)v[0] | ((uint64_t)v[1] << 32);
        uint64_t w1 = (uint64_t)v[2] | ((uint64_t)v[3] << 32);
        uint64_t w2 = (uint64_t)v[4] | ((uint64_t)v[5] << 32);
        uint64_t w3 = (uint64_t)v[6] | ((uint64_t)v[7] << 32);
        l[0] = w0 & M52;
        l[1] = ((w0 >> 52) | (w1 << 12)) & M52;
        l[2] = ((w1 >> 40) | (w2 << 24)) & M52;
        l[3] = ((w2 >> 28) | (w3 << 36)) & M52;
        l[4] = w3 >> 16;
    }
};

// PR-specific 52-bit constants: modulus limbs and -p^-1 mod 2^52, derived at compile time from PR::P / PR::INV.
template <class PR>
struct Dfma52 {
    static G16_HD constexpr uint64_t p64(int i) { return (uint64_t)PR::P(2 * i) | ((uint64_t)PR::P(2 * i + 1) << 32); }
    static G16_HD constexpr uint64_t P52(int i) {
        return i == 0   ? (p64(0) & dfma::M52)
               : i == 1 ? (((p64(0) >> 52) | (p64(1) << 12)) & dfma::M52)
               : i == 2 ? (((p64(1) >> 40) | (p64(2) << 24)) & dfma::M52)
               : i == 3 ? (((p64(2) >> 28) | (p64(3) << 36)) & dfma::M52)
                        : (p64(3) >> 16);
    }
    // Newton iteration for -p^-1 mod 2^64 from the 32-bit constant, truncated to 52 bits
    static G16_HD constexpr uint64_t NP52() {
        uint64_t p0 = p64(0);
        uint64_t x = (uint64_t)(0u - PR::INV);  // p^-1 mod 2^32
        x = x * (2 - p0 * x);                    // mod 2^64
        return (0 - x) & dfma::M52;
    }
};

// Montgomery product a * b * 2^-256 mod p on the FP64 pipe; same result bits as Fp::mul_cios.
template <class PR>
G16_HD Fp<PR> mul_dfma(const Fp<PR>& a, const Fp<PR>& b) {
    typedef Dfma52<PR> K;
    uint64_t al[5], bl[5];
    dfma::split52(a.v, al);
    dfma::split52(b.v, bl);
    double ad[5], bd[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ad[i] = dfma::to_double(al[i]);
        bd[i] = dfma::to_double(bl[i]);
    }
    // column accumulators, pre-loaded with minus the exponent biases they are going to receive:
    // column k gets lo terms of (i + j == k) and hi terms of (i + j == k - 1), once for a*b and once for q*p
    uint64_t col[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        int nlo_ab = (k <= 4) ? k + 1 : (k <= 8 ? 9 - k : 0);
        int nhi_ab = (k >= 1) ? ((k - 1 <= 4) ? k : (k - 1 <= 8 ? 10 - k : 0)) : 0;
        // q_i * p_j lands in columns i + j (lo) and i + j + 1 (hi) for i, j in 0..4 as well
        uint64_t bias = (uint64_t)(2 * nlo_ab) * dfma::EXP52 + (uint64_t)(2 * nhi_ab) * dfma::EXP104;
        col[k] = 0 - bias;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            dfma::mul_hi_lo(ad[i], bd[j], hi, lo);
            col[i + j] += lo;
            col[i + j + 1] += hi;
        }
    }
    const double pd[5] = {(double)K::P52(0), (double)K::P52(1), (double)K::P52(2), (double)K::P52(3), (double)K::P52(4)};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        // q = -col[i] / p mod 2^52 (2^48 in the last round: 4 * 52 + 48 = 256)
        uint64_t q = (col[i] * K::NP52()) & (i < 4 ? dfma::M52 : dfma::M48);
        double qd = dfma::to_double(q);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            dfma::mul_hi_lo(qd, pd[j], hi, lo);
            col[i + j] += lo;
            col[i + j + 1] += hi;
        }
        if (i < 4) col[i + 1] += col[i] >> 52;  // low 52 bits are zero now
    }
    // value = (col[4] >> 48) + 2^4 * sum_k col[5 + k] * 2^(52 k): renormalise the columns to 52-bit digits, then repack
    uint64_t u = col[4] >> 48;  // < 2^10
    uint64_t e[5];
    uint64_t c = u >> 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t w = col[5 + k] + c;
        e[k] = w & dfma::M52;
        c = w >> 52;
    }
    e[4] = col[9] + c;  // result < 2p < 2^255: the top digit needs no mask
    uint64_t w0 = (u & 15) | (e[0] << 4) | (e[1] << 56);
    uint64_t w1 = (e[1] >> 8) | (e[2] << 44);
    uint64_t w2 = (e[2] >> 20) | (e[3] << 32);
    uint64_t w3 = (e[3] >> 32) | (e[4] << 20);
    Fp<PR> r;
    r.v[0] = (uint32_t)w0;
    r.v[1] = (uint32_t)(w0 >> 32);
    r.v[2] = (uint32_t)w1;
    r.v[3] = (uint32_t)(w1 >> 32);
    r.v[4] = (uint32_t)w2;
    r.v[5] = (uint32_t)(w2 >> 32);
    r.v[6] = (uint32_t)w3;
    r.v[7] = (uint32_t)(w3 >> 32);
    Fp<PR>::reduce_once(r.v, 0);
    return r;
}

}  // namespace g16
