// fp_wide.cuh -- 8x8-limb products without reduction (16 limbs) and the separate Montgomery reduction.
//
// The integer pipe is the roofline of every heavy kernel (ncu: FMA-heavy pipe 73 % busy in the bucket loop, DRAM < 9 %),
// and one IMAD.WIDE.U32 keeps it busy for 4 cycles, so the way to go faster is to issue fewer of them:
//   * product by one level of subtractive Karatsuba: 3 x (4x4 limbs) = 48 wide multiplies instead of 64,
//   * reduction as a separate step (64 wide multiplies), which also lets Fq2 and sums of products share reductions.
// The additions Karatsuba costs go to the ALU pipe (IADD3), which has slack.
// Every carry chain is one asm statement with a faithful C emulation for the host build (tests/host_fp_shim.cpp).
#pragma once
#include <stdint.h>

namespace g16 {

// acc[0..3] = {x0, x1} * b  (two disjoint 64-bit lanes)
G16_HD void lanes2_mul(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %6;\n\t"
        "mul.hi.u32 %3, %5, %6;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3])
        : "r"(x0), "r"(x1), "r"(b));
#else
    uint64_t p = (uint64_t)x0 * b, q = (uint64_t)x1 * b;
    acc[0] = (uint32_t)p;
    acc[1] = (uint32_t)(p >> 32);
    acc[2] = (uint32_t)q;
    acc[3] = (uint32_t)(q >> 32);
#endif
}

// acc[0..3] += {x0, x1} * b as one carry chain; returns the carry out of acc[3]
G16_HD uint32_t lanes2_mad(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t b) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(cy)
        : "r"(x0), "r"(x1), "r"(b));
#else
    const uint32_t x[2] = {x0, x1};
    uint64_t c = 0;
    for (int k = 0; k < 2; k++) {
        uint64_t p = (uint64_t)x[k] * b;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    cy = (uint32_t)c;
#endif
    return cy;
}

// r[0..3] = a - b, returns borrow
G16_HD uint32_t sub4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    bw &= 1u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t t = (uint64_t)a[i] - b[i] - c;
        r[i] = (uint32_t)t;
        c = (t >> 32) & 1u;
    }
    bw = (uint32_t)c;
#endif
    return bw;
}

// r = (x ^ mask) + sel over 4 limbs: conditional two's-complement negation (mask = 0 / ~0, sel = 0 / 1)
G16_HD void cneg4(uint32_t* r, const uint32_t* x, uint32_t mask, uint32_t sel) {
#ifdef __CUDA_ARCH__
    uint32_t t0 = x[0] ^ mask, t1 = x[1] ^ mask, t2 = x[2] ^ mask, t3 = x[3] ^ mask;
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, 0;\n\t"
        "addc.cc.u32 %2, %6, 0;\n\t"
        "addc.u32 %3, %7, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3])
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(sel));
#else
    uint64_t c = sel;
    for (int i = 0; i < 4; i++) {
        uint64_t t = (uint64_t)(x[i] ^ mask) + c;
        r[i] = (uint32_t)t;
        c = t >> 32;
    }
#endif
}

// r[0..7] = a + b + cin (cin in {0,1}), returns carry out
G16_HD uint32_t add8c(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "add.cc.u32 t, %25, 0xffffffff;\n\t"  // CC = cin
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        "}"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cy)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
#else
    uint64_t c = cin;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)t;
        c = t >> 32;
    }
    cy = (uint32_t)c;
#endif
    return cy;
}

// r[0..3] = a[0..3] + small (a 32-bit value); no carry out expected by the callers
G16_HD void add4s(uint32_t* r, const uint32_t* a, uint32_t small) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, 0;\n\t"
        "addc.cc.u32 %2, %6, 0;\n\t"
        "addc.u32 %3, %7, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(small));
#else
    uint64_t c = small;
    for (int i = 0; i < 4; i++) {
        uint64_t t = (uint64_t)a[i] + c;
        r[i] = (uint32_t)t;
        c = t >> 32;
    }
#endif
}

// 4 x 4 limbs -> 8 limbs: 16 wide multiplies on two interleaved accumulators (even / odd limb alignment)
G16_HD void mul4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t E[8], O[8];  // E: positions 0..7; O: position k+1 for limb k
    uint32_t cy;
    lanes2_mul(E, a[0], a[2], b[0]);           // a0b0 @0, a2b0 @2
    lanes2_mul(O, a[1], a[3], b[0]);           // a1b0 @1, a3b0 @3
    E[4] = E[5] = E[6] = E[7] = 0;
    O[4] = O[5] = O[6] = O[7] = 0;
    cy = lanes2_mad(O, a[0], a[2], b[1]);      // a0b1 @1, a2b1 @3
    O[4] = cy;
    lanes2_mad(E + 2, a[1], a[3], b[1]);       // a1b1 @2, a3b1 @4  (E[4], E[5] fresh: no carry out)
    cy = lanes2_mad(E + 2, a[0], a[2], b[2]);  // a0b2 @2, a2b2 @4
    E[6] = cy;
    lanes2_mad(O + 2, a[1], a[3], b[2]);       // a1b2 @3, a3b2 @5  (O[5] fresh, O[4] <= 1: no carry out)
    cy = lanes2_mad(O + 2, a[0], a[2], b[3]);  // a0b3 @3, a2b3 @5
    O[6] = cy;
    lanes2_mad(E + 4, a[1], a[3], b[3]);       // a1b3 @4, a3b3 @6  (product < 2^256: no carry out)
    // r = E + (O << 32)
    uint32_t sh[8];
    sh[0] = 0;
#pragma unroll
    for (int k = 1; k < 8; k++) sh[k] = O[k - 1];
    add8(r, E, sh);
}

// 8 x 8 limbs -> 16 limbs, subtractive Karatsuba:
//   a*b = z0 + (z0 + z2 + (aL - aH)(bH - bL)) * 2^128 + z2 * 2^256
G16_HD void mul8_wide(uint32_t* P, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], m[8], da[4], db[4], t[8];
    mul4(z0, a, b);
    mul4(z2, a + 4, b + 4);
    uint32_t sa = sub4(da, a, a + 4);      // aL - aH
    uint32_t sb = sub4(db, b + 4, b);      // bH - bL
    cneg4(da, da, 0u - sa, sa);            // |aL - aH|
    cneg4(db, db, 0u - sb, sb);            // |bH - bL|
    mul4(m, da, db);
    uint32_t neg = sa ^ sb;                // 1: the cross term is negative -> subtract m
    uint32_t c1 = add8(t, z0, z2);
    uint32_t mask = 0u - neg;
#pragma unroll
    for (int k = 0; k < 8; k++) m[k] ^= mask;
    uint32_t c = add8c(t, t, m, neg);      // t +/- m  (two's complement when neg)
    c1 = c1 + c - neg;                     // top bit of the 257-bit middle term: 0 or 1
    // P = z0 + t * 2^128 + z2 * 2^256
    uint32_t mid[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        P[k] = z0[k];
        mid[k] = z0[k + 4];
        mid[k + 4] = z2[k];
    }
    uint32_t c2 = add8(P + 4, mid, t);
    add4s(P + 12, z2 + 4, c1 + c2);
}

}  // namespace g16
