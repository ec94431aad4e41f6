// witness.cu -- sparse R1CS evaluation (K2), the pointwise H-quotient (K4) and the R1CS->QAP witness map.
//
// Follows LibsnarkReduction::witness_map_from_matrices (forks/groth16/src/r1cs_to_qap.rs:150-213) and, as the
// secondary variant, CircomReduction::witness_map_from_matrices (forks/circom-compat/src/circom/qap.rs:25-90).
// Row evaluation is evaluate_constraint (r1cs_to_qap.rs:16-45) including its coeff == 1 add-only fast path.
#include "internal.cuh"

namespace g16 {

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    Fr x;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    *reinterpret_cast<uint4*>(&x.v[0]) = __ldg(q);
    *reinterpret_cast<uint4*>(&x.v[4]) = __ldg(q + 1);
    return x;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& x) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = *reinterpret_cast<const uint4*>(&x.v[0]);
    q[1] = *reinterpret_cast<const uint4*>(&x.v[4]);
}

// out[perm(i)] = <row_i, z>; perm = bit reversal over log_n bits when log_n_rev != 0 (feeds the DIT iNTT directly)
__global__ void k_spmv(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                       const Fr* __restrict__ val, const Fr* __restrict__ z, Fr* __restrict__ out, uint64_t nc,
                       unsigned log_n_rev) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    uint64_t lo = row_ptr[i], hi = row_ptr[i + 1];
    const Fr one = Fr::one();
    Fr sum = Fr::zero();
    for (uint64_t k = lo; k < hi; k++) {
        Fr c = ld_fr(val + k);
        Fr w = ld_fr(z + col[k]);
        if (c == one)
            sum = sum + w;
        else
            sum = sum + w * c;
    }
    uint64_t o = log_n_rev ? (uint64_t)(__brev((unsigned)i) >> (32 - log_n_rev)) : i;
    st_fr(out + o, sum);
}

// a[perm(nc + i)] = z[i], i < ni   (r1cs_to_qap.rs:173-177)
__global__ void k_place_inputs(const Fr* __restrict__ z, Fr* __restrict__ a, uint64_t nc, uint64_t ni,
                               unsigned log_n_rev) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    uint64_t p = nc + i;
    uint64_t o = log_n_rev ? (uint64_t)(__brev((unsigned)p) >> (32 - log_n_rev)) : p;
    st_fr(a + o, ld_fr(z + i));
}

// c = a*b  (qap.rs:52-58) ;  a = a*b - c (qap.rs:74-88)
__global__ void k_mul_into(const Fr* __restrict__ a, const Fr* __restrict__ b, Fr* __restrict__ c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(c + i, ld_fr(a + i) * ld_fr(b + i));
}
__global__ void k_mul_sub(Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, ld_fr(a + i) * ld_fr(b + i) - ld_fr(c + i));
}

static int spmv(g16_ctx* ctx, int k, Fr* out, unsigned log_n_rev, cudaStream_t st) {
    if (ctx->nc == 0) return G16_OK;
    unsigned blocks = (unsigned)((ctx->nc + 127) / 128);
    G16_LAUNCH(ctx, k_spmv, blocks, 128, 0, st, ctx->mat[k].row_ptr, ctx->mat[k].col, ctx->mat[k].val, ctx->d_z, out,
               ctx->nc, log_n_rev);
    return G16_OK;
}

int r1cs_eval_dev(g16_ctx* ctx, Fr* az, Fr* bz, Fr* cz, bool bitrev, cudaStream_t st) {
    unsigned lr = bitrev ? ctx->log_n : 0;
    if (az) G16_TRY(spmv(ctx, 0, az, lr, st));
    if (bz) G16_TRY(spmv(ctx, 1, bz, lr, st));
    if (cz) G16_TRY(spmv(ctx, 2, cz, lr, st));
    return G16_OK;
}

int witness_map_dev(g16_ctx* ctx, int reduction, cudaStream_t st) {
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map: no R1CS loaded");
    NttTables* t;
    G16_TRY(ntt_get_tables(ctx, ctx->log_n, &t));
    const size_t n = (size_t)1 << ctx->log_n;
    const unsigned lr = ctx->log_n;  // log_n == 0: __brev path disabled, identity
    Fr *a = ctx->d_a, *b = ctx->d_b, *c = ctx->d_c;
    G16_CUDA(ctx, cudaMemsetAsync(a, 0, n * sizeof(Fr), st));
    G16_CUDA(ctx, cudaMemsetAsync(b, 0, n * sizeof(Fr), st));
    G16_CUDA(ctx, cudaMemsetAsync(c, 0, n * sizeof(Fr), st));
    const unsigned eb = (unsigned)((n + 255) / 256);
    if (reduction == G16_REDUCTION_LIBSNARK) {
        if (!t->zinv_ok) return set_err(ctx, G16_ERR_VANISHING_ZERO, "g^n - 1 == 0");
        G16_TRY(r1cs_eval_dev(ctx, a, b, c, true, st));
        G16_LAUNCH(ctx, k_place_inputs, (unsigned)((ctx->ni + 127) / 128), 128, 0, st, ctx->d_z, a, ctx->nc, ctx->ni, lr);
        // The reference runs seven transforms: iFFT and coset FFT of a, b and c, the pointwise (a*b - c) / Z, one coset
        // iFFT (r1cs_to_qap.rs:179-210).  Z is CONSTANT on the coset (Z(g w^i) = g^n - 1), so by linearity
        //     h = coset_iFFT((A'*B' - C') / Z) = (coset_iFFT(A'*B') - iFFT(c)) / (g^n - 1)
        // as exact field identities, whatever the witness: c needs ONE transform, and the quotient kernel disappears.
        //   a, b: iNTT (1/n deferred into the coset table) -> coset NTT, evaluations bit-reversed
        G16_TRY(ntt_dit(ctx, a, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dit(ctx, b, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dif(ctx, a, t, false, t->coset_scaled, st));
        G16_TRY(ntt_dif(ctx, b, t, false, t->coset_scaled, st));
        //   c: coefficients / (g^n - 1), natural order
        G16_TRY(ntt_dit(ctx, c, t, true, nullptr, &t->zinv_n, st));
        //   coset iNTT of a*b (product taken on the first load; g^-i / (n (g^n - 1)) and the subtraction on the last store)
        G16_TRY(ntt_dit(ctx, a, t, true, t->coset_inv_z, nullptr, st, b, c));
    } else if (reduction == G16_REDUCTION_CIRCOM) {
        if (ctx->log_n >= 28) return set_err(ctx, G16_ERR_DEGREE_TOO_LARGE, "circom reduction needs a 2n domain");
        G16_TRY(r1cs_eval_dev(ctx, a, b, nullptr, true, st));
        G16_LAUNCH(ctx, k_place_inputs, (unsigned)((ctx->ni + 127) / 128), 128, 0, st, ctx->d_z, a, ctx->nc, ctx->ni, lr);
        // c_i = a_i * b_i for i < nc; b is zero on the padding rows so the product is zero there as the reference's c
        G16_LAUNCH(ctx, k_mul_into, eb, 256, 0, st, a, b, c, n);
        G16_TRY(ntt_dit(ctx, a, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dit(ctx, b, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dit(ctx, c, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dif(ctx, a, t, false, t->odd_scaled, st));
        G16_TRY(ntt_dif(ctx, b, t, false, t->odd_scaled, st));
        G16_TRY(ntt_dif(ctx, c, t, false, t->odd_scaled, st));
        G16_LAUNCH(ctx, k_mul_sub, eb, 256, 0, st, a, b, c, n);
        G16_TRY(bitrev_permute(ctx, a, ctx->log_n, st));  // evaluations are the output here: natural order needed
    } else {
        return set_err(ctx, G16_ERR_BAD_ARG, "unknown reduction %d", reduction);
    }
    return G16_OK;
}

}  // namespace g16
