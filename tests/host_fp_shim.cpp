// Host build of csrc/fp.cuh (carry chains emulated in C) exposed for ctypes; lets the CPU test-suite
// exercise the exact control flow of the device Montgomery code against the big-integer oracle.
#include "../crescent_credentials_b200/csrc/glv.cuh"
#include <string.h>
using namespace g16;
template <class F> static void bin(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    F x, y, z;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    switch (op) {
        case 0: z = x * y; break;
        case 1: z = x + y; break;
        case 2: z = x - y; break;
        case 3: z = x.neg(); break;
        case 4: z = x.inverse(); break;
        case 5: z = x.to_mont(); break;
        case 6: z = x.from_mont(); break;
        case 7: z = x.sqr(); break;
        case 8: z = F::mul_karatsuba(x, y); break;
        case 9: z = x.inverse_bgcd(); break;
        case 10: z = x.inverse_safegcd(); break;
        default: z = F::zero();
    }
    memcpy(r, z.v, 32);
}
extern "C" void host_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) { bin<Fr>(op, a, b, r); }
extern "C" void host_fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) { bin<Fq>(op, a, b, r); }
extern "C" void host_fq2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fq2 x, y, z;
    memcpy(&x, a, 64);
    memcpy(&y, b, 64);
    switch (op) {
        case 0: z = x * y; break;
        case 1: z = x + y; break;
        case 2: z = x - y; break;
        case 3: z = x.neg(); break;
        case 4: z = x.inverse(); break;
        case 7: z = x.sqr(); break;
        default: z = Fq2::zero();
    }
    memcpy(r, &z, 64);
}

// k * P for an affine G1 point (16 words x || y, Montgomery) and a canonical 256-bit scalar: which = 0 plain double-and-add
// (scalar_mul), 1 signed 4-bit windows with inlined group operations (scalar_mul_window); out = affine Montgomery, (0, 0) = infinity
extern "C" void host_g1_scalar_mul(const uint32_t* p, const uint32_t* k, int which, uint32_t* out) {
    G1Affine a;
    memcpy(&a, p, 64);
    G1XYZZ x = G1XYZZ::from_affine(a);
    G1XYZZ r = which ? scalar_mul_window(x, k) : scalar_mul(x, k);
    G1Affine o = r.to_affine();
    memcpy(out, &o, 64);
}

// GLV split of a canonical scalar (glv.cuh): out = k1 (5 words) | neg1 | k2 (5 words) | neg2 | ok
extern "C" void host_glv_decompose(const uint32_t* k, uint32_t* out) {
    GlvScalar s = glv_decompose(k);
    for (int h = 0; h < 2; h++) {
        for (int i = 0; i < 5; i++) out[6 * h + i] = s.h[h].k[i];
        out[6 * h + 5] = (uint32_t)s.h[h].neg;
    }
    out[12] = (uint32_t)s.ok;
}
// k * P through the split: k1 * P + k2 * phi(P), the two halves one after the other on the host
extern "C" void host_g1_scalar_mul_glv(const uint32_t* p, const uint32_t* k, uint32_t* out) {
    G1Affine a;
    memcpy(&a, p, 64);
    G1XYZZ x = G1XYZZ::from_affine(a);
    GlvScalar s = glv_decompose(k);
    G1XYZZ r = glv_half_mul(x, s.h[0], 0);
    r.add_inl(glv_half_mul(x, s.h[1], 1));
    G1Affine o = r.to_affine();
    memcpy(out, &o, 64);
}
