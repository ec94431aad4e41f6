// ec.cuh -- BN254 G1 (over Fq) and G2 (over Fq2) group arithmetic for the MSM and assembly kernels.
//
// Replaces ark-ec 0.4 short_weierstrass::{Affine, Projective} as used by the reference at
// forks/groth16/src/prover.rs:66,74,76-80,94-98,116-118,124-135,256-274.  Curve: y^2 = x^3 + b, a = 0.
// Bucket accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): a mixed addition is
// 8M + 2S and never needs the curve constant.  Results are exact group elements, so they equal the
// reference's Jacobian results after normalisation.
//
// Affine infinity is encoded as (0, 0) (not on either curve since b != 0); XYZZ infinity as ZZ == 0.
#pragma once
#include "fp.cuh"

namespace g16 {

template <class F>
struct Affine {
    F x, y;
    G16_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static G16_HD Affine inf() { return Affine{F::zero(), F::zero()}; }
    G16_HD Affine neg() const { return Affine{x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    G16_HD bool is_inf() const { return zz.is_zero(); }
    static G16_HD XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    static G16_HD XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
    G16_HD XYZZ neg() const { return XYZZ{x, y.neg(), zz, zzz}; }

    // doubling, dbl-2008-s-1 with a = 0: 6M + 3S.  dbl() is a real call (code size); dbl_inl() is the same body inlined, for
    // the latency-bound single-lane chains (bucket reduction) where the call's trip through local memory costs more than the code
    G16_HD_NOINLINE XYZZ dbl() const { return dbl_inl(); }
    G16_HD XYZZ dbl_inl() const {
        if (is_inf()) return *this;
        F u = y.dbl();
        if (u.is_zero()) return inf();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // doubling of an affine point (mdbl-2008-s-1)
    static G16_HD_NOINLINE XYZZ dbl_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        F u = p.y.dbl();
        if (u.is_zero()) return inf();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * p.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // mixed addition acc += p (madd-2008-s): 8M + 2S
    G16_HD void madd(const Affine<F>& p) {
        if (p.is_inf()) return;
        if (is_inf()) {
            x = p.x;
            y = p.y;
            zz = F::one();
            zzz = F::one();
            return;
        }
        F u2 = p.x * zz;
        F s2 = p.y * zzz;
        F pp = u2 - x;
        F r = s2 - y;
        if (pp.is_zero()) {
            if (r.is_zero()) {
                *this = dbl_affine(p);
            } else {
                *this = inf();
            }
            return;
        }
        F p2 = pp.sqr();
        F p3 = pp * p2;
        F q = x * p2;
        F nx = r.sqr() - p3 - q.dbl();
        y = r * (q - nx) - y * p3;
        x = nx;
        zz = zz * p2;
        zzz = zzz * p3;
    }

    // full addition acc += o (add-2008-s): 12M + 2S.  add() is a real call, add_inl() the same body inlined (see dbl_inl)
    G16_HD_NOINLINE void add(const XYZZ& o) { add_inl(o); }
    G16_HD void add_inl(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F pp = u2 - u1;
        F r = s2 - s1;
        if (pp.is_zero()) {
            if (r.is_zero()) {
                *this = dbl();
            } else {
                *this = inf();
            }
            return;
        }
        F p2 = pp.sqr();
        F p3 = pp * p2;
        F q = u1 * p2;
        F nx = r.sqr() - p3 - q.dbl();
        y = r * (q - nx) - s1 * p3;
        x = nx;
        zz = zz * o.zz * p2;
        zzz = zzz * o.zzz * p3;
    }

    G16_HD_NOINLINE Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        // one inversion: i3 = 1/zzz, and zz^3 = zzz^2  =>  1/zz = zz^2 / zzz^2 = (zz * i3)^2
        F i3 = zzz.inverse_fast();
        F t = zz * i3;
        F i2 = t.sqr();
        return Affine<F>{x * i2, y * i3};
    }
};

// k * P for a 256-bit little-endian canonical (non-Montgomery) scalar held in 8 limbs; plain
// double-and-add, used only by the handful of scalar multiplications in proof assembly
// (prover.rs:76-80,94,98,104,116,118) and by the fixed-base key generator.
template <class F>
G16_HD_NOINLINE XYZZ<F> scalar_mul(const XYZZ<F>& p, const uint32_t* k) {
    XYZZ<F> acc = XYZZ<F>::inf();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = acc.dbl();
            if ((k[i] >> bit) & 1u) {
                acc.add(p);
                started = true;
            }
        }
    }
    return acc;
}

// Same product for the one place where a fresh-point scalar multiplication is a lone lane's critical path (k_scale_point:
// s * MSM_a, r * MSM_b1): signed 4-bit windows -- a table of 1P..8P (4 doublings + 3 additions), then per nibble 4 doublings +
// at most 1 addition -- with the group operations inlined, none behind a call.  `nibbles` = how many low nibbles of k count
// (64 for a full 256-bit scalar: ~3100 dependent field products instead of ~4060; 33 for the halves of a GLV split).
template <class F>
G16_HD XYZZ<F> scalar_mul_window(const XYZZ<F>& p, const uint32_t* k, int nibbles = 64) {
    XYZZ<F> m[9];  // m[d] = d * P
    m[1] = p;
    m[2] = p.dbl_inl();
    m[3] = m[2];
    m[3].add_inl(p);
    m[4] = m[2].dbl_inl();
    m[5] = m[4];
    m[5].add_inl(p);
    m[6] = m[3].dbl_inl();
    m[7] = m[6];
    m[7].add_inl(p);
    m[8] = m[4].dbl_inl();
    // signed digits in [-7, 8], least significant first; one more digit takes the last carry
    signed char digit[65];
    int carry = 0;
#pragma unroll 1
    for (int j = 0; j < nibbles; j++) {
        int d = (int)((k[j >> 3] >> (4 * (j & 7))) & 15u) + carry;
        carry = d > 8;
        digit[j] = (signed char)(carry ? d - 16 : d);
    }
    digit[nibbles] = (signed char)carry;
    XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
    for (int j = nibbles; j >= 0; j--) {
        if (j != nibbles) {
#pragma unroll 1
            for (int t = 0; t < 4; t++) acc = acc.dbl_inl();
        }
        const int d = digit[j];
        if (d > 0) acc.add_inl(m[d]);
        else if (d < 0) acc.add_inl(m[-d].neg());
    }
    return acc;
}

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace g16
