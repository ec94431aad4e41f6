// Host build of csrc/pairing.cuh (same source as the verifier kernels, carry chains emulated in C) exposed for ctypes: the
// CPU test-suite runs the device pairing code against the big-integer oracle (oracle/pairing.py) without a GPU.
// All field elements cross as Montgomery 8 x u32 limbs; an Fq12 is 12 of them in ark-serialize order.
#include "../crescent_credentials_b200/csrc/pairing.cuh"
#include <string.h>
#include <vector>
using namespace g16;

extern "C" {
// op: 0 mul, 1 sqr, 2 inverse, 3 conj, 4 cyclotomic_sqr, 5/6/7 frobenius 1/2/3, 8 cyclotomic_exp_x
void host_f12_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fq12 x, y, z;
    memcpy(&x, a, sizeof(Fq12));
    memcpy(&y, b, sizeof(Fq12));
    switch (op) {
        case 0: z = x * y; break;
        case 1: z = x.sqr(); break;
        case 2: z = x.inverse(); break;
        case 3: z = x.conj(); break;
        case 4: z = x.cyclotomic_sqr(); break;
        case 5: z = x.frobenius(1); break;
        case 6: z = x.frobenius(2); break;
        case 7: z = x.frobenius(3); break;
        case 8: z = cyclotomic_exp_x(x); break;
        default: z = Fq12::one();
    }
    memcpy(out, &z, sizeof(Fq12));
}
void host_f6_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fq6 x, y, z;
    memcpy(&x, a, sizeof(Fq6));
    memcpy(&y, b, sizeof(Fq6));
    switch (op) {
        case 0: z = x * y; break;
        case 2: z = x.inverse(); break;
        case 9: z = x.mul_by_01(y.c0, y.c1); break;
        default: z = Fq6::one();
    }
    memcpy(out, &z, sizeof(Fq6));
}
int host_ell_coeffs(void) { return kEllCoeffs; }
void host_g2_prepare(const uint32_t* q, uint32_t* out) {
    G2Affine Q;
    memcpy(&Q, q, sizeof(Q));
    std::vector<EllCoeff> c(kEllCoeffs);
    g2_prepare(Q, c.data());
    memcpy(out, c.data(), sizeof(EllCoeff) * kEllCoeffs);
}
// multi_miller_loop over three pairs: (p0, q0) on the fly, (p1, t1) and (p2, t2) from prepared tables; act = 3 flags
void host_miller3(const uint32_t* p, const int* act, const uint32_t* q0, const uint32_t* t1, const uint32_t* t2, uint32_t* out) {
    G1Affine P[3];
    memcpy(P, p, sizeof(P));
    bool a[3] = {act[0] != 0, act[1] != 0, act[2] != 0};
    G2Affine Q;
    memcpy(&Q, q0, sizeof(Q));
    Fq12 f = miller_loop3(P, a, Q, (const EllCoeff*)t1, (const EllCoeff*)t2);
    memcpy(out, &f, sizeof(f));
}
int host_final_exp(const uint32_t* f, uint32_t* out) {
    Fq12 x;
    memcpy(&x, f, sizeof(x));
    bool ok;
    Fq12 r = final_exponentiation(x, ok);
    memcpy(out, &r, sizeof(r));
    return ok ? 1 : 0;
}
// window tables of one gamma_abc point, built the way the kernel does (one entry = scalar_mul + to_affine)
void host_abc_table(const uint32_t* point, uint32_t* out /* 32 * 255 points */) {
    G1Affine P;
    memcpy(&P, point, sizeof(P));
    G1Affine* o = (G1Affine*)out;
    for (int w = 0; w < kAbcWindows; w++)
        for (int d = 1; d <= kAbcDigits; d++) {
            uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            k[w >> 2] = (uint32_t)d << (8 * (w & 3));
            o[w * kAbcDigits + d - 1] = scalar_mul(G1XYZZ::from_affine(P), k).to_affine();
        }
}
void host_prepare_inputs(const uint32_t* abc0, const uint32_t* tbl, const uint32_t* inputs, size_t n, uint32_t* out) {
    G1Affine a0;
    memcpy(&a0, abc0, sizeof(a0));
    G1Affine r = prepare_inputs_one(a0, (const G1Affine*)tbl, (const Fr*)inputs, n);
    memcpy(out, &r, sizeof(r));
}
int host_verify_one(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* prepared, const uint32_t* ng,
                    const uint32_t* nd, const uint32_t* alpha_beta) {
    G1Affine A, C, PI;
    G2Affine B;
    Fq12 ab;
    memcpy(&A, a, sizeof(A));
    memcpy(&B, b, sizeof(B));
    memcpy(&C, c, sizeof(C));
    memcpy(&PI, prepared, sizeof(PI));
    memcpy(&ab, alpha_beta, sizeof(ab));
    return verify_one(A, B, C, PI, (const EllCoeff*)ng, (const EllCoeff*)nd, ab);
}
}
