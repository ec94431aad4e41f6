// fp.cuh -- BN254 Fr / Fq Montgomery arithmetic on 8 x 32-bit limbs for sm_100a.
//
// Replaces the arithmetic the reference gets from ark-ff 0.4 `Fp<MontBackend<_,4>,4>` (used at
// forks/groth16/src/r1cs_to_qap.rs:16-45,187,205-208 and prover.rs:63-65).  The in-memory form is
// identical: 4 x u64 LE limbs of a*2^256 mod p == 8 x u32 LE limbs.
//
// Multiplication is a word-serial (CIOS) Montgomery product.  A 32x32 product occupies a 64-bit
// lane, so partial products of the even limbs of an operand never overlap each other, nor do those of
// the odd limbs: we keep two accumulators ("ev" aligned at limb 0, "od" aligned at limb 1) and every
// row of the product is two straight carry chains of mad.lo.cc / madc.hi.cc that ptxas turns into
// IMAD.WIDE.U32 with carry-in/out.  Each chain lives inside ONE asm statement, so we never rely on the
// carry flag surviving between statements.
//
// The same source compiles for the host (gcc) with the chains emulated in C: tests/host_fp_test.cpp
// runs exactly this control flow against the big-integer oracle without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define G16_HD __host__ __device__ __forceinline__
#define G16_D __device__ __forceinline__
// big cold-path routines (full additions, doublings, inversions) are real calls: keeps ptxas time and code size sane
#define G16_HD_NOINLINE __host__ __device__ __noinline__
#else
#define G16_HD inline
#define G16_D inline
#define G16_HD_NOINLINE inline
#endif

// The Montgomery product is ~180 straight-line instructions (2.9 KB of SASS).  The pairing of verify.cu holds ~260 inlined
// copies and runs with 2 warps per scheduler; ncu shows stall_no_instruction 1.8 per issue at saturation.  A translation unit
// may define G16_MUL_AS_CALL before including this header to make the product ONE shared function instead (verify.cu does
// with -DG16_VERIFY_MUL_CALL): 147 k -> 94 k SASS instructions, but MEASURED SLOWER (754 k vs 807 k proofs/s, 26.9 vs 23.0 ms
// for one proof): the call sequence costs more than the fetch stalls it removes.  Kept as a switch for the record.
#ifdef G16_MUL_AS_CALL
#define G16_MUL_ATTR G16_HD_NOINLINE
#else
#define G16_MUL_ATTR G16_HD
#endif

namespace g16 {

struct limbs8 {
    uint32_t v[8];
};

// ---------------------------------------------------------------------------------------------
// carry-chain primitives.  acc is 8 limbs = four 64-bit lanes; (x0,x1,x2,x3) are the four 32-bit
// multiplicands feeding those lanes.
// ---------------------------------------------------------------------------------------------

// acc = {x0,x1,x2,x3} * b   (no carries: lanes are disjoint)
G16_HD void lanes_mul(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]),
          "=&r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    for (int k = 0; k < 4; k++) {
        uint64_t p = (uint64_t)x[k] * b;
        acc[2 * k] = (uint32_t)p;
        acc[2 * k + 1] = (uint32_t)(p >> 32);
    }
#endif
}

// acc += {x0,x1,x2,x3} * b as one carry chain; returns the carry out of limb 7 (0 or 1)
G16_HD uint32_t lanes_mad(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(cy)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t c = 0;
    for (int k = 0; k < 4; k++) {
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

// The per-row "fold + shift + multiply-add" step of the interleaved Montgomery product.
//   e0  += sh[1]                      (stray limb of the accumulator being shifted out; carry -> next limb)
//   sh[k] = sh[k+2] + lane product    (accumulator moves down 64 bits while the new row is added)
//   sh[6],sh[7] = top lane product (+ carry)
G16_HD void lanes_fold_shift_mad(uint32_t& e0, uint32_t* sh, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                                 uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(e0), "+r"(sh[0]), "+r"(sh[1]), "+r"(sh[2]), "+r"(sh[3]), "+r"(sh[4]), "+r"(sh[5]), "+r"(sh[6]),
          "+r"(sh[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t t = (uint64_t)e0 + sh[1];
    e0 = (uint32_t)t;
    uint64_t c = t >> 32;
    for (int k = 0; k < 4; k++) {
        uint64_t p = (uint64_t)x[k] * b;
        uint32_t add_lo = (k < 3) ? sh[2 * k + 2] : 0u;
        uint32_t add_hi = (k < 3) ? sh[2 * k + 3] : 0u;
        uint64_t lo = (uint64_t)add_lo + (uint32_t)p + c;
        uint64_t hi = (uint64_t)add_hi + (uint32_t)(p >> 32) + (lo >> 32);
        sh[2 * k] = (uint32_t)lo;
        sh[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
#endif
}

// r = a + b (8 limbs), returns carry
G16_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=r"(cy)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)t;
        c = t >> 32;
    }
    cy = (uint32_t)c;
#endif
    return cy;
}

// r = a - b (8 limbs), returns borrow (1 if a < b)
G16_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    bw &= 1u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] - b[i] - c;
        r[i] = (uint32_t)t;
        c = (t >> 32) & 1u;
    }
    bw = (uint32_t)c;
#endif
    return bw;
}

// e0 += sh[1] (carry into the next limb); sh = sh >> 64 with that carry added at limb 0; sh[6] = sh[7] = 0.
// The accumulator shift of the stand-alone Montgomery reduction (no product row rides on it).
G16_HD void lanes_fold_shift(uint32_t& e0, uint32_t* sh) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.cc.u32 %1, %3, 0;\n\t"
        "addc.cc.u32 %2, %4, 0;\n\t"
        "addc.cc.u32 %3, %5, 0;\n\t"
        "addc.cc.u32 %4, %6, 0;\n\t"
        "addc.cc.u32 %5, %7, 0;\n\t"
        "addc.u32 %6, %8, 0;"
        : "+r"(e0), "+r"(sh[0]), "+r"(sh[1]), "+r"(sh[2]), "+r"(sh[3]), "+r"(sh[4]), "+r"(sh[5])
        : "r"(sh[6]), "r"(sh[7]));
#else
    uint64_t t = (uint64_t)e0 + sh[1];
    e0 = (uint32_t)t;
    uint64_t c = t >> 32;
    for (int k = 0; k < 6; k++) {
        uint64_t v = (uint64_t)sh[k + 2] + c;
        sh[k] = (uint32_t)v;
        c = v >> 32;
    }
#endif
    sh[6] = 0;
    sh[7] = 0;
}

// lo += x with the carry going into hi (hi cannot overflow in the callers)
G16_HD void add_carry_into(uint32_t& lo, uint32_t& hi, uint32_t x) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, 0;"
        : "+r"(lo), "+r"(hi)
        : "r"(x));
#else
    uint64_t t = (uint64_t)lo + x;
    lo = (uint32_t)t;
    hi += (uint32_t)(t >> 32);
#endif
}

}  // namespace g16
#include "fp_wide.cuh"
namespace g16 {

// ---------------------------------------------------------------------------------------------
// Field parameters.  P = modulus limbs, INV = -p^-1 mod 2^32, R1 = 2^256 mod p (Montgomery one),
// R2 = 2^512 mod p.  Values: SURVEY appendix, re-derived by oracle/pyref.py in tests/test_constants.py.
// ---------------------------------------------------------------------------------------------
struct FrParams {
    static G16_HD constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;
    static G16_HD constexpr uint32_t R1(int i) {
        constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    static G16_HD constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return t[i];
    }
    static G16_HD constexpr uint32_t R3(int i) {  // 2^768 mod r
        constexpr uint32_t t[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return t[i];
    }
};
struct FqParams {
    static G16_HD constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;
    static G16_HD constexpr uint32_t R1(int i) {
        constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    static G16_HD constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return t[i];
    }
    static G16_HD constexpr uint32_t R3(int i) {  // 2^768 mod q
        constexpr uint32_t t[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
        return t[i];
    }
};

template <class PR>
struct alignas(16) Fp {
    uint32_t v[8];  // Montgomery form, fully reduced: 0 <= v < p

    static G16_HD void load_p(uint32_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = PR::P(i);
    }
    static G16_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    static G16_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = PR::R1(i);
        return r;
    }
    static G16_HD Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = PR::R2(i);
        return r;
    }
    G16_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    G16_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    G16_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // conditional subtract of p: x in [0, 2p) (+ optional carry bit) -> [0, p)
    static G16_HD void reduce_once(uint32_t* x, uint32_t carry) {
        uint32_t t[8], p[8];
        load_p(p);
        uint32_t bw = sub8(t, x, p);
        // keep t when there was no borrow, or when the 257-bit value had its top bit set
        bool take = (bw == 0) | (carry != 0);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = take ? t[i] : x[i];
    }

    friend G16_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t cy = add8(r.v, a.v, b.v);  // p < 2^254 so cy is always 0; kept for generality
        reduce_once(r.v, cy);
        return r;
    }
    friend G16_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t bw = sub8(r.v, a.v, b.v);
        uint32_t t[8], p[8];
        load_p(p);
        add8(t, r.v, p);
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = bw ? t[i] : r.v[i];
        return r;
    }
    G16_HD Fp neg() const {
        Fp r;
        uint32_t p[8];
        load_p(p);
        sub8(r.v, p, v);
        bool z = is_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
        return r;
    }
    G16_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*2^-256 mod p, word-serial CIOS (128 wide multiplies); kept as the reference formulation
    static G16_HD Fp mul_cios(const Fp& a, const Fp& b) {
        uint32_t X[8], Y[8];  // the two accumulators; their even/odd roles alternate every row
        uint32_t m, cy;
        // row 0
        lanes_mul(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);  // even-aligned
        lanes_mul(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);  // odd-aligned (limb k of Y sits at position k+1)
        m = X[0] * PR::INV;
        lanes_mad(Y, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);  // cannot carry out (total < 2^288)
        cy = lanes_mad(X, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
        Y[7] += cy;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t* ev = (i & 1) ? Y : X;  // becomes the even-aligned accumulator after the 32-bit shift
            uint32_t* od = (i & 1) ? X : Y;  // old even accumulator: limb 0 is zero, limb 1 is folded into ev[0]
            lanes_fold_shift_mad(ev[0], od, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            cy = lanes_mad(ev, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
            od[7] += cy;
            m = ev[0] * PR::INV;
            lanes_mad(od, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);
            cy = lanes_mad(ev, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
            od[7] += cy;
        }
        // after row 7 the even accumulator is Y (i=7 odd -> ev=Y), odd is X; result = (ev >> 32) + od
        Fp r;
        uint32_t sh[8];
#pragma unroll
        for (int k = 0; k < 7; k++) sh[k] = Y[k + 1];
        sh[7] = 0;
        cy = add8(r.v, sh, X);
        reduce_once(r.v, cy);
        return r;
    }
    // Three independent Montgomery products in lock-step: row i of every product before row i + 1, so that the six carry chains
    // (two per product) sit next to each other in the instruction stream.  ptxas keeps the rows of ONE product in order (each
    // is an asm block with a carry chain) and, given three separate mul_cios calls, emits the products one after the other;
    // written this way a thread has three times as many independent IMAD.WIDE chains in flight.  Same arithmetic, same result
    // bits as three mul_cios calls.  Used by Fq2::operator* when G16_FQ2_MUL3 is defined (off by default: prepared at the end
    // of round 1, checked on the host against the oracle, not yet measured on the GPU).
    static G16_HD void mul_cios3(const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, Fp& r0, Fp& r1,
                                 Fp& r2) {
        uint32_t A[3][8], B[3][8];  // operand copies: fully unrolled below, so these live in registers
#pragma unroll
        for (int k = 0; k < 8; k++) {
            A[0][k] = a0.v[k], A[1][k] = a1.v[k], A[2][k] = a2.v[k];
            B[0][k] = b0.v[k], B[1][k] = b1.v[k], B[2][k] = b2.v[k];
        }
        uint32_t X[3][8], Y[3][8];
        uint32_t m[3], cy[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            lanes_mul(X[j], A[j][0], A[j][2], A[j][4], A[j][6], B[j][0]);
            lanes_mul(Y[j], A[j][1], A[j][3], A[j][5], A[j][7], B[j][0]);
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            m[j] = X[j][0] * PR::INV;
            lanes_mad(Y[j], PR::P(1), PR::P(3), PR::P(5), PR::P(7), m[j]);
            cy[j] = lanes_mad(X[j], PR::P(0), PR::P(2), PR::P(4), PR::P(6), m[j]);
            Y[j][7] += cy[j];
        }
#pragma unroll
        for (int i = 1; i < 8; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                uint32_t* ev = (i & 1) ? Y[j] : X[j];
                uint32_t* od = (i & 1) ? X[j] : Y[j];
                lanes_fold_shift_mad(ev[0], od, A[j][1], A[j][3], A[j][5], A[j][7], B[j][i]);
                cy[j] = lanes_mad(ev, A[j][0], A[j][2], A[j][4], A[j][6], B[j][i]);
                od[7] += cy[j];
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                uint32_t* ev = (i & 1) ? Y[j] : X[j];
                uint32_t* od = (i & 1) ? X[j] : Y[j];
                m[j] = ev[0] * PR::INV;
                lanes_mad(od, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m[j]);
                cy[j] = lanes_mad(ev, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m[j]);
                od[7] += cy[j];
            }
        }
        Fp out[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            uint32_t sh[8];
#pragma unroll
            for (int k = 0; k < 7; k++) sh[k] = Y[j][k + 1];
            sh[7] = 0;
            uint32_t c = add8(out[j].v, sh, X[j]);
            reduce_once(out[j].v, c);
        }
        r0 = out[0];
        r1 = out[1];
        r2 = out[2];
    }
    // Montgomery reduction of a 16-limb value P < p * 2^256:  P * 2^-256 mod p, fully reduced.
    // Same two-accumulator scheme as mul_cios with the product rows removed: the high limbs of P enter one per round
    // at the top of the accumulator.
    static G16_HD Fp reduce_wide(const uint32_t* P) {
        uint32_t X[8], Y[8];
        uint32_t m, cy;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            X[k] = P[k];
            Y[k] = 0;
        }
        m = X[0] * PR::INV;
        lanes_mad(Y, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);
        cy = lanes_mad(X, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
        Y[7] += cy;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t* ev = (i & 1) ? Y : X;
            uint32_t* od = (i & 1) ? X : Y;
            lanes_fold_shift(ev[0], od);
            add_carry_into(ev[7], od[7], P[7 + i]);
            m = ev[0] * PR::INV;
            lanes_mad(od, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);
            cy = lanes_mad(ev, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
            od[7] += cy;
        }
        Fp r;
        uint32_t sh[8];
#pragma unroll
        for (int k = 0; k < 7; k++) sh[k] = Y[k + 1];
        sh[7] = P[15];
        cy = add8(r.v, sh, X);
        reduce_once(r.v, cy);
        return r;
    }

    // Montgomery product.  Default: word-serial CIOS (123 IMAD.WIDE + 8 IMAD.HI + 9 IMAD + 39 IADD3 in SASS).
    // G16_MUL_KARATSUBA selects the 16-limb Karatsuba product + separate reduction (103 IMAD.WIDE + 15 IMAD.HI but
    // 132 IADD3 + 44 SEL/LOP3): measured SLOWER on B200 (56.2 vs 64.9 G mul/s, bucket loop 8.5 vs 6.3 ms) because the
    // extra ALU work and code size cost more than the 10 % of wide multiplies they save.  Kept for the record and
    // because reduce_wide is the building block for sharing reductions between products.
    friend G16_MUL_ATTR Fp operator*(const Fp& a, const Fp& b) {
#ifdef G16_MUL_KARATSUBA
        uint32_t P[16];
        mul8_wide(P, a.v, b.v);
        return reduce_wide(P);
#else
        return mul_cios(a, b);
#endif
    }
    static G16_HD Fp mul_karatsuba(const Fp& a, const Fp& b) {
        uint32_t P[16];
        mul8_wide(P, a.v, b.v);
        return reduce_wide(P);
    }
    G16_HD Fp sqr() const { return (*this) * (*this); }

    // to / from Montgomery form
    G16_HD Fp to_mont() const { return (*this) * r2(); }
    G16_HD Fp from_mont() const {
        Fp o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = (i == 0) ? 1u : 0u;
        return (*this) * o;
    }

    // a^(p-2); exponent bits are compile-time constants of the field
    G16_HD_NOINLINE Fp inverse() const {
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = PR::P(i);
        e[0] -= 2u;  // both moduli have low limb >= 2, no borrow
        Fp acc = one();
        for (int i = 7; i >= 0; i--) {
            for (int bit = 31; bit >= 0; bit--) {
                acc = acc.sqr();
                if ((e[i] >> bit) & 1u) acc = acc * (*this);
            }
        }
        return acc;
    }

    // Same inverse by the binary extended Euclidean algorithm: shifts and subtractions only.  A Montgomery product is a
    // ~1000-cycle dependent chain for a lone warp, so the 380-product Fermat ladder is ~0.2 ms of pure latency; this loop
    // runs ~380 iterations of a few carry chains each (~8x less latency).  Used where an inversion sits alone on the
    // critical path (the shared inversions of the batched-affine bucket accumulation).  inverse of 0 is 0.
    G16_HD_NOINLINE Fp inverse_bgcd() const {
        if (is_zero()) return zero();
        // invariants: x1 * a == u, x2 * a == w (mod p) for the integer a = this->v (i.e. value * R)
        uint32_t u[8], w[8];
        Fp x1, x2;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u[i] = v[i];
            w[i] = PR::P(i);
            x1.v[i] = (i == 0) ? 1u : 0u;
            x2.v[i] = 0u;
        }
        auto is_one = [](const uint32_t* x) {
            uint32_t o = x[0] ^ 1u;
#pragma unroll
            for (int i = 1; i < 8; i++) o |= x[i];
            return o == 0;
        };
        auto shr1 = [](uint32_t* x) {
#pragma unroll
            for (int i = 0; i < 7; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
            x[7] >>= 1;
        };
        auto half = [&](Fp& x) {  // x / 2 mod p
            if (x.v[0] & 1u) {
                uint32_t p[8], t[8];
                load_p(p);
                add8(t, x.v, p);  // < 2^255: no carry out
#pragma unroll
                for (int i = 0; i < 8; i++) x.v[i] = t[i];
            }
            shr1(x.v);
        };
        for (int it = 0; it < 1024; it++) {
            if (is_one(u) || is_one(w)) break;
            if ((u[0] & 1u) && (w[0] & 1u)) {
                uint32_t t[8];
                uint32_t bw = sub8(t, u, w);
                if (!bw) {
#pragma unroll
                    for (int i = 0; i < 8; i++) u[i] = t[i];
                    x1 = x1 - x2;
                } else {
                    sub8(t, w, u);
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = t[i];
                    x2 = x2 - x1;
                }
            }
            if (!(u[0] & 1u)) {
                shr1(u);
                half(x1);
            } else {
                shr1(w);
                half(x2);
            }
        }
        Fp y = is_one(u) ? x1 : x2;  // (value * R)^-1 as a plain integer = value^-1 * R^-1
        Fp r3;
#pragma unroll
        for (int i = 0; i < 8; i++) r3.v[i] = PR::R3(i);
        return y * r3;  // * R^3 / R  =>  value^-1 * R
    }

    // ---- inverse by divsteps (Bernstein-Yang "safegcd", the variable-time form of libsecp256k1's modinv32) -----------------
    // The binary GCD above touches all eight limbs of four 256-bit values in every one of its ~400-500 iterations: ~100 us for
    // the lone lane that inverts in k_ba_invert -- five of those per MSM chain are the largest fixed latency of a small MSM
    // (25 x 0.12 ms per proof).  divsteps decide everything from the LOW bits of (f, g): 30 steps run on one
    // 32-bit word each and yield a 2x2 transition matrix, which is then applied once to the full-size values -- about 20 rounds of
    // ~40 word-multiplications instead of ~450 rounds of multi-limb shifts and subtractions.  Same value as inverse() /
    // inverse_bgcd() (the inverse is unique); inverse of 0 is 0.  Values travel as nine signed 30-bit limbs.
    struct S30 {
        int32_t v[9];
    };
    static G16_HD unsigned ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
        return (unsigned)(__ffs((int)x) - 1);
#else
        return (unsigned)__builtin_ctz(x);
#endif
    }
    static G16_HD uint32_t p_inv30() {  // p^-1 mod 2^30 (Newton from p = p^-1 mod 8)
        uint32_t p0 = PR::P(0), x = p0;
        x *= 2u - p0 * x;
        x *= 2u - p0 * x;
        x *= 2u - p0 * x;
        x *= 2u - p0 * x;
        return x & 0x3FFFFFFFu;
    }
    static G16_HD void to_s30(S30& o, const uint32_t* w) {  // 8 x 32 unsigned -> 9 x 30
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
            uint64_t two = w[k];
            if (k + 1 < 8) two |= (uint64_t)w[k + 1] << 32;
            o.v[i] = (int32_t)((uint32_t)(two >> sh) & 0x3FFFFFFFu);
        }
    }
    // 30 divsteps on the low words; returns the new eta and the transition matrix t = (u v; q r), entries in (-2^30, 2^30]
    static G16_HD int32_t divsteps_30(int32_t eta, uint32_t f0, uint32_t g0, int32_t* t) {
        uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
        int i = 30;
        for (;;) {
            const unsigned zeros = ctz32(g | (0xFFFFFFFFu << i));  // sentinel: never more than i
            g >>= zeros;
            u <<= zeros;
            v <<= zeros;
            eta -= (int32_t)zeros;
            i -= (int)zeros;
            if (i == 0) break;
            if (eta < 0) {
                uint32_t tmp;
                eta = -eta;
                tmp = f, f = g, g = 0u - tmp;
                tmp = u, u = q, q = 0u - tmp;
                tmp = v, v = r, r = 0u - tmp;
            }
            // cancel up to min(i, eta + 1, 6) low bits of g with a multiple of f:  w = -g / f mod 2^limit,  -1/f = f (f^2 - 2) mod 64
            int limit = (eta + 1) > i ? i : (eta + 1);
            const uint32_t m = (0xFFFFFFFFu >> (32 - limit)) & 63u;
            const uint32_t w = (f * g * (f * f - 2u)) & m;
            g += f * w;
            q += u * w;
            r += v * w;
        }
        t[0] = (int32_t)u, t[1] = (int32_t)v, t[2] = (int32_t)q, t[3] = (int32_t)r;
        return eta;
    }
    // (f, g) <- t (f, g) / 2^30, exactly
    static G16_HD void update_fg_30(S30& f, S30& g, const int32_t* t) {
        const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
        int64_t cf = u * f.v[0] + v * g.v[0];
        int64_t cg = q * f.v[0] + r * g.v[0];
        cf >>= 30;
        cg >>= 30;
#pragma unroll
        for (int i = 1; i < 9; i++) {
            const int64_t fi = f.v[i], gi = g.v[i];
            cf += u * fi + v * gi;
            cg += q * fi + r * gi;
            f.v[i - 1] = (int32_t)cf & 0x3FFFFFFF;
            cf >>= 30;
            g.v[i - 1] = (int32_t)cg & 0x3FFFFFFF;
            cg >>= 30;
        }
        f.v[8] = (int32_t)cf;
        g.v[8] = (int32_t)cg;
    }
    // (d, e) <- t (d, e) / 2^30 mod p, staying in (-2p, p)
    static G16_HD void update_de_30(S30& d, S30& e, const int32_t* t, const S30& P, uint32_t pinv30) {
        const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
        const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;  // all-ones when negative
        int32_t md = (u & sd) + (v & se);
        int32_t me = (q & sd) + (r & se);
        int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
        int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
        // multiples of p that zero the low 30 bits
        md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & 0x3FFFFFFFu);
        me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & 0x3FFFFFFFu);
        cd += (int64_t)P.v[0] * md;
        ce += (int64_t)P.v[0] * me;
        cd >>= 30;
        ce >>= 30;
#pragma unroll
        for (int i = 1; i < 9; i++) {
            cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i] + (int64_t)P.v[i] * md;
            ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i] + (int64_t)P.v[i] * me;
            d.v[i - 1] = (int32_t)cd & 0x3FFFFFFF;
            cd >>= 30;
            e.v[i - 1] = (int32_t)ce & 0x3FFFFFFF;
            ce >>= 30;
        }
        d.v[8] = (int32_t)cd;
        e.v[8] = (int32_t)ce;
    }
    G16_HD_NOINLINE Fp inverse_safegcd() const {
        if (is_zero()) return zero();
        S30 P, f, g, d, e;
        uint32_t pw[8];
        load_p(pw);
        to_s30(P, pw);
        f = P;
        to_s30(g, v);
#pragma unroll
        for (int i = 0; i < 9; i++) d.v[i] = 0, e.v[i] = 0;
        e.v[0] = 1;
        const uint32_t pinv30 = p_inv30();
        int32_t eta = -1;
        for (int it = 0; it < 40; it++) {  // <= 25 rounds suffice for 256-bit inputs; the bound only guards against misuse
            int32_t t[4];
            eta = divsteps_30(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
            update_de_30(d, e, t, P, pinv30);
            update_fg_30(f, g, t);
            int32_t nz = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) nz |= g.v[i];
            if (nz == 0) break;
        }
        // now f = +-1 and d = +-(this->v as an integer)^-1 mod p, d in (-2p, p): fix the sign, then bring d into [0, p)
        const int32_t fneg = f.v[8] >> 31;
        int32_t c = 0;
        const int32_t dneg = d.v[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d.v[i] += P.v[i] & dneg;         // d < 0: add p  -> (-p, p)
#pragma unroll
        for (int i = 0; i < 9; i++) d.v[i] = (d.v[i] ^ fneg) - fneg;  // f = -1: negate
#pragma unroll
        for (int i = 0; i < 8; i++) {                                  // propagate carries, limbs back to [0, 2^30)
            d.v[i] += c;
            c = d.v[i] >> 30;
            d.v[i] &= 0x3FFFFFFF;
        }
        d.v[8] += c;
        const int32_t dneg2 = d.v[8] >> 31;
        c = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {                                  // still negative: add p once more -> [0, p)
            d.v[i] += (P.v[i] & dneg2) + c;
            c = i < 8 ? d.v[i] >> 30 : 0;
            if (i < 8) d.v[i] &= 0x3FFFFFFF;
        }
        Fp y;
#pragma unroll
        for (int k = 0; k < 8; k++) {  // 9 x 30 -> 8 x 32
            const int bit = 32 * k, i = bit / 30, sh = bit % 30;
            uint64_t two = (uint64_t)(uint32_t)d.v[i] | ((uint64_t)(uint32_t)d.v[i + 1] << 30);
            y.v[k] = (uint32_t)(two >> sh);
        }
        Fp r3;
#pragma unroll
        for (int i = 0; i < 8; i++) r3.v[i] = PR::R3(i);
        return y * r3;  // (value * R)^-1 * R^3 / R = value^-1 * R
    }
    // what the kernels call where one inversion sits alone on a critical path
    G16_HD Fp inverse_fast() const {
#ifdef G16_INV_BGCD
        return inverse_bgcd();
#else
        return inverse_safegcd();
#endif
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// ---------------------------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2+1)   (ark-bn254 Fq2Config: NONRESIDUE = -1)
// ---------------------------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    static G16_HD Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    static G16_HD Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    G16_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    G16_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    G16_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
    friend G16_HD Fq2 operator+(const Fq2& a, const Fq2& b) { return Fq2{a.c0 + b.c0, a.c1 + b.c1}; }
    friend G16_HD Fq2 operator-(const Fq2& a, const Fq2& b) { return Fq2{a.c0 - b.c0, a.c1 - b.c1}; }
    G16_HD Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
    G16_HD Fq2 dbl() const { return Fq2{c0.dbl(), c1.dbl()}; }
    // Karatsuba: 3 Fq multiplications.  Deliberately a real call on the device: with everything inlined the G2 bucket
    // loop is 28 Fq products (~110 KB of SASS), overflows the instruction cache (ncu: stall_no_instruction second
    // largest) and needs 255 registers; as calls it runs 12 % faster at 168 registers (profiles/r01_notes.md).
    friend G16_HD_NOINLINE Fq2 operator*(const Fq2& a, const Fq2& b) {
#ifdef G16_FQ2_MUL3
        Fq v0, v1, s;  // the three products in lock-step (mul_cios3): six carry chains in flight instead of two
        Fq::mul_cios3(a.c0, b.c0, a.c1, b.c1, a.c0 + a.c1, b.c0 + b.c1, v0, v1, s);
#else
        Fq v0 = a.c0 * b.c0;
        Fq v1 = a.c1 * b.c1;
        Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
#endif
        return Fq2{v0 - v1, s - v0 - v1};
    }
    // complex squaring: 2 Fq multiplications
    G16_HD_NOINLINE Fq2 sqr() const {
        Fq t = c0 * c1;
        return Fq2{(c0 + c1) * (c0 - c1), t.dbl()};
    }
    G16_HD_NOINLINE Fq2 inverse() const {
        Fq n = (c0.sqr() + c1.sqr()).inverse();
        return Fq2{c0 * n, (c1 * n).neg()};
    }
    G16_HD_NOINLINE Fq2 inverse_fast() const {  // same value, low-latency Fq inversion
        Fq n = (c0.sqr() + c1.sqr()).inverse_fast();
        return Fq2{c0 * n, (c1 * n).neg()};
    }
};

}  // namespace g16
