// Host build of csrc/fp.cuh (carry chains emulated in C) exposed for ctypes; lets the CPU test-suite
// exercise the exact control flow of the device Montgomery code against the big-integer oracle.
#include "../crescent_credentials_b200/csrc/ec.cuh"
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
