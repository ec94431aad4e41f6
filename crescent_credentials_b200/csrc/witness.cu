// witness.cu -- sparse R1CS evaluation (K2), the pointwise H-quotient (K4) and the R1CS->QAP witness map.
//
// Follows LibsnarkReduction::witness_map_from_matrices (forks/groth16/src/r1cs_to_qap.rs:150-213) and, as the
// secondary variant, CircomReduction::witness_map_from_matrices (forks/circom-compat/src/circom/qap.rs:25-90).
// Row evaluation is evaluate_constraint (r1cs_to_qap.rs:16-45) including its coeff == 1 add-only fast path.
#include <algorithm>
#include <vector>

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

// rows nc .. n-1 of the three vectors: a[perm(nc + i)] = z[i] for i < ni (r1cs_to_qap.rs:173-177), zero padding above and
// in b / c.  Rows below nc are written by the SpMV, so no vector needs a memset of its own.
__global__ void k_pad_rows(const Fr* __restrict__ z, Fr* __restrict__ a, Fr* __restrict__ b, Fr* __restrict__ c, uint64_t nc,
                           uint64_t ni, uint64_t n, unsigned log_n_rev) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t p = nc + i;
    if (p >= n) return;
    uint64_t o = log_n_rev ? (uint64_t)(__brev((unsigned)p) >> (32 - log_n_rev)) : p;
    if (a) st_fr(a + o, i < ni ? ld_fr(z + i) : Fr::zero());
    if (b) st_fr(b + o, Fr::zero());
    if (c) st_fr(c + o, Fr::zero());
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

// ---- sliced-ELL SpMV ----------------------------------------------------------------------------------------------------
// The row-per-thread CSR kernel above is bound by divergence, not by memory: R1CS rows are short on average (3-5 terms)
// with a long tail, so a warp loops to the length of its longest row (~4x the mean) and reads coefficients with a
// row-length stride.  Here the rows are sorted by length once at load time and cut into slices of 32 (one warp): every lane
// of a warp has (almost) the same trip count, entry k of the 32 rows of a slice is contiguous (coalesced column and
// coefficient reads), and the coefficient class -- +1, -1, anything else -- rides in the top bits of the column word so
// that the ~70-90 % unit coefficients of a circom system cost no coefficient read and no product (the coeff == 1 fast path of
// evaluate_constraint, r1cs_to_qap.rs:16-45, extended to -1).  The sum of a row is exact field arithmetic: order-free.
enum { SELL_GENERAL = 0, SELL_ONE = 1, SELL_MINUS_ONE = 2, SELL_PAD = 3 };

__global__ void k_sell_build(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const Fr* __restrict__ val,
                             const uint32_t* __restrict__ sell_row, const uint64_t* __restrict__ sell_ptr, size_t slices,
                             uint32_t* __restrict__ sell_col, Fr* __restrict__ sell_val) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t slice = t >> 5;
    if (slice >= slices) return;
    unsigned lane = (unsigned)(t & 31);
    uint64_t base = sell_ptr[slice];
    uint64_t width = (sell_ptr[slice + 1] - base) >> 5;
    uint32_t row = sell_row[t];
    uint64_t lo = 0, len = 0;
    if (row != kSellNoRow) {
        lo = row_ptr[row];
        len = row_ptr[row + 1] - lo;
    }
    const Fr one = Fr::one(), minus_one = Fr::one().neg();
    for (uint64_t k = 0; k < width; k++) {
        uint64_t e = base + (k << 5) + lane;
        if (k < len) {
            Fr c = val[lo + k];
            uint32_t cls = c == one ? SELL_ONE : (c == minus_one ? SELL_MINUS_ONE : SELL_GENERAL);
            sell_col[e] = col[lo + k] | (cls << 30);
            sell_val[e] = c;
        } else {
            sell_col[e] = (uint32_t)SELL_PAD << 30;
            sell_val[e] = Fr::zero();
        }
    }
}

__global__ void __launch_bounds__(128)
    k_spmv_sell(const uint32_t* __restrict__ sell_row, const uint64_t* __restrict__ sell_ptr, const uint32_t* __restrict__ sell_col,
                const Fr* __restrict__ sell_val, size_t slices, const Fr* __restrict__ z, Fr* __restrict__ out,
                unsigned log_n_rev) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t slice = t >> 5;
    if (slice >= slices) return;
    unsigned lane = (unsigned)(t & 31);
    uint64_t e = sell_ptr[slice] + lane, end = sell_ptr[slice + 1];
    Fr sum = Fr::zero();
    uint32_t cw = e < end ? __ldg(sell_col + e) : ((uint32_t)SELL_PAD << 30);
    for (; e < end; e += 32) {
        uint32_t cur = cw;
        if (e + 32 < end) cw = __ldg(sell_col + e + 32);  // next step's column word while this step's gather is in flight
        uint32_t cls = cur >> 30;
        if (cls == SELL_PAD) continue;
        Fr w = ld_fr(z + (cur & 0x3FFFFFFFu));
        if (cls == SELL_ONE)
            sum = sum + w;
        else if (cls == SELL_MINUS_ONE)
            sum = sum - w;
        else
            sum = sum + w * ld_fr(sell_val + e);
    }
    uint32_t row = sell_row[t];
    if (row == kSellNoRow) return;
    uint64_t o = log_n_rev ? (uint64_t)(__brev(row) >> (32 - log_n_rev)) : row;
    st_fr(out + o, sum);
}

void sell_free(CsrDev* m) {
    dev_free(m->sell_row);
    dev_free(m->sell_ptr);
    dev_free(m->sell_col);
    dev_free(m->sell_val);
    m->sell_row = nullptr;
    m->sell_ptr = nullptr;
    m->sell_col = nullptr;
    m->sell_val = nullptr;
    m->slices = 0;
}

int sell_build(g16_ctx* ctx, int k, const uint64_t* rp, cudaStream_t st) {
    CsrDev& m = ctx->mat[k];
    sell_free(&m);
    const uint64_t nc = ctx->nc;
    if (nc == 0 || nc >= kSellNoRow) return G16_OK;
    // rows by descending length (stable: equal lengths keep their order, so neighbouring rows stay neighbours)
    std::vector<uint32_t> order(nc);
    for (uint64_t i = 0; i < nc; i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [rp](uint32_t a, uint32_t b) { return rp[a + 1] - rp[a] > rp[b + 1] - rp[b]; });
    const size_t slices = (size_t)((nc + 31) / 32);
    std::vector<uint32_t> rows(slices * 32, kSellNoRow);
    std::vector<uint64_t> ptr(slices + 1, 0);
    for (size_t s = 0; s < slices; s++) {
        uint64_t first = (uint64_t)s * 32;
        uint32_t r0 = order[first];
        uint64_t width = rp[r0 + 1] - rp[r0];  // the slice's longest row comes first
        for (uint64_t l = 0; l < 32 && first + l < nc; l++) rows[first + l] = order[first + l];
        ptr[s + 1] = ptr[s] + width * 32;
    }
    const uint64_t total = ptr[slices];
    G16_TRY(dev_alloc(ctx, &m.sell_row, slices * 32));
    G16_TRY(dev_alloc(ctx, &m.sell_ptr, slices + 1));
    G16_TRY(dev_alloc(ctx, &m.sell_col, (size_t)total));
    G16_TRY(dev_alloc(ctx, &m.sell_val, (size_t)total));
    G16_CUDA(ctx, cudaMemcpyAsync(m.sell_row, rows.data(), slices * 32 * 4, cudaMemcpyHostToDevice, st));
    G16_CUDA(ctx, cudaMemcpyAsync(m.sell_ptr, ptr.data(), (slices + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total)
        G16_LAUNCH(ctx, k_sell_build, (unsigned)((slices * 32 + 127) / 128), 128, 0, st, m.row_ptr, m.col, m.val, m.sell_row, m.sell_ptr,
                   slices, m.sell_col, m.sell_val);
    G16_CUDA(ctx, cudaStreamSynchronize(st));  // rows / ptr are host temporaries
    m.slices = slices;
    return G16_OK;
}

static int spmv(g16_ctx* ctx, int k, Fr* out, unsigned log_n_rev, cudaStream_t st) {
    if (ctx->nc == 0) return G16_OK;
    const CsrDev& m = ctx->mat[k];
    if (ctx->opt_spmv_sell && m.slices) {
        G16_LAUNCH(ctx, k_spmv_sell, (unsigned)((m.slices * 32 + 127) / 128), 128, 0, st, m.sell_row, m.sell_ptr, m.sell_col, m.sell_val,
                   m.slices, ctx->d_z, out, log_n_rev);
        return G16_OK;
    }
    unsigned blocks = (unsigned)((ctx->nc + 127) / 128);
    G16_LAUNCH(ctx, k_spmv, blocks, 128, 0, st, m.row_ptr, m.col, m.val, ctx->d_z, out, ctx->nc, log_n_rev);
    return G16_OK;
}

int r1cs_eval_dev(g16_ctx* ctx, Fr* az, Fr* bz, Fr* cz, bool bitrev, cudaStream_t st) {
    unsigned lr = bitrev ? ctx->log_n : 0;
    if (az) G16_TRY(spmv(ctx, 0, az, lr, st));
    if (bz) G16_TRY(spmv(ctx, 1, bz, lr, st));
    if (cz) G16_TRY(spmv(ctx, 2, cz, lr, st));
    return G16_OK;
}

// The LibsnarkReduction witness map cut into its three independent pipelines and the final transform, so that the host glue of
// a sharded run can place them on different GPUs (sharded.py, plan "wm_split"): parts is a mask of
//   G16_WM_PART_A      a <- coset_NTT(iNTT(A z ++ inputs))          (r1cs_to_qap.rs:164-185, a side)
//   G16_WM_PART_B      b <- coset_NTT(iNTT(B z))                    (:164-185, b side)
//   G16_WM_PART_C      c <- iNTT(C z) / (g^n - 1)                   (:191-199 with the constant-Z identity of witness_map_dev)
//   G16_WM_PART_FINAL  a <- coset_iNTT(a * b) / (g^n - 1) - c = h   (:187,201-210)
// Running all four in this order is witness_map_dev(LIBSNARK) bit for bit (same kernels, same tables).
int witness_map_part_dev(g16_ctx* ctx, int parts, cudaStream_t st) {
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map_part: no R1CS loaded");
    NttTables* t;
    G16_TRY(ntt_get_tables(ctx, ctx->log_n, &t));
    if (!t->zinv_ok) return set_err(ctx, G16_ERR_VANISHING_ZERO, "g^n - 1 == 0");
    const size_t n = (size_t)1 << ctx->log_n;
    const unsigned lr = ctx->log_n;
    Fr *a = ctx->d_a, *b = ctx->d_b, *c = ctx->d_c;
    const unsigned pb = (unsigned)((n - ctx->nc + 127) / 128);
    if (parts & G16_WM_PART_A) {
        G16_TRY(r1cs_eval_dev(ctx, a, nullptr, nullptr, true, st));
        G16_LAUNCH(ctx, k_pad_rows, pb, 128, 0, st, ctx->d_z, a, (Fr*)nullptr, (Fr*)nullptr, ctx->nc, ctx->ni, (uint64_t)n, lr);
        G16_TRY(ntt_dit(ctx, a, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dif(ctx, a, t, false, t->coset_scaled, st));
    }
    if (parts & G16_WM_PART_B) {
        G16_TRY(r1cs_eval_dev(ctx, nullptr, b, nullptr, true, st));
        G16_LAUNCH(ctx, k_pad_rows, pb, 128, 0, st, ctx->d_z, (Fr*)nullptr, b, (Fr*)nullptr, ctx->nc, ctx->ni, (uint64_t)n, lr);
        G16_TRY(ntt_dit(ctx, b, t, true, nullptr, nullptr, st));
        G16_TRY(ntt_dif(ctx, b, t, false, t->coset_scaled, st));
    }
    if (parts & G16_WM_PART_C) {
        G16_TRY(r1cs_eval_dev(ctx, nullptr, nullptr, c, true, st));
        G16_LAUNCH(ctx, k_pad_rows, pb, 128, 0, st, ctx->d_z, (Fr*)nullptr, (Fr*)nullptr, c, ctx->nc, ctx->ni, (uint64_t)n, lr);
        G16_TRY(ntt_dit(ctx, c, t, true, nullptr, &t->zinv_n, st));
    }
    if (parts & G16_WM_PART_FINAL) G16_TRY(ntt_dit(ctx, a, t, true, t->coset_inv_z, nullptr, st, b, c));
    return G16_OK;
}

int witness_map_dev(g16_ctx* ctx, int reduction, cudaStream_t st) {
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map: no R1CS loaded");
    NttTables* t;
    G16_TRY(ntt_get_tables(ctx, ctx->log_n, &t));
    const size_t n = (size_t)1 << ctx->log_n;
    const unsigned lr = ctx->log_n;  // log_n == 0: __brev path disabled, identity
    Fr *a = ctx->d_a, *b = ctx->d_b, *c = ctx->d_c;
    const unsigned eb = (unsigned)((n + 255) / 256);
    const unsigned pb = (unsigned)((n - ctx->nc + 127) / 128);  // n >= nc + ni > nc
    if (reduction == G16_REDUCTION_LIBSNARK) {
        if (!t->zinv_ok) return set_err(ctx, G16_ERR_VANISHING_ZERO, "g^n - 1 == 0");
        G16_TRY(r1cs_eval_dev(ctx, a, b, c, true, st));
        G16_LAUNCH(ctx, k_pad_rows, pb, 128, 0, st, ctx->d_z, a, b, c, ctx->nc, ctx->ni, (uint64_t)n, lr);
        // The reference runs seven transforms: iFFT and coset FFT of a, b and c, the pointwise (a*b - c) / Z, one coset
        // iFFT (r1cs_to_qap.rs:179-210).  Z is CONSTANT on the coset (Z(g w^i) = g^n - 1), so by linearity
        //     h = coset_iFFT((A'*B' - C') / Z) = (coset_iFFT(A'*B') - iFFT(c)) / (g^n - 1)
        // as exact field identities, whatever the witness: c needs ONE transform, and the quotient kernel disappears.
        //   a, b: iNTT (1/n deferred into the coset table) -> coset NTT, evaluations bit-reversed
        //   c:    iNTT with 1 / (n (g^n - 1)) on the last store: coefficients / (g^n - 1), natural order
        // a lone witness map (a shard rank without wire MSMs, serialised runs) batches the three inverse transforms into one
        // launch per pass and the two coset transforms into another (ntt.cu: NttBatch; 3.29 -> 3.21 ms).  Beside the MSM chains
        // the batched launches were measured to cost the whole proof +4 ms (33.7 vs 29.9 ms, profiles/
        // r02_ab_ba_add_ntt_batch.log): 3072-block launches of 64 KB-shared-memory blocks crowd the DRAM-bound MSM kernels out
        // of the SMs, so there the transforms stay one modest launch each.
        Fr* abc[3] = {a, b, c};
        if (ctx->opt_ntt_batch > 0 || (ctx->opt_ntt_batch < 0 && ctx->wm_alone)) {
            G16_TRY(ntt_dit_batch(ctx, abc, 3, t, true, nullptr, &t->zinv_n, 4u, st));
            G16_TRY(ntt_dif_batch(ctx, abc, 2, t, false, t->coset_scaled, st));
        } else {
            G16_TRY(ntt_dit(ctx, a, t, true, nullptr, nullptr, st));
            G16_TRY(ntt_dit(ctx, b, t, true, nullptr, nullptr, st));
            G16_TRY(ntt_dif(ctx, a, t, false, t->coset_scaled, st));
            G16_TRY(ntt_dif(ctx, b, t, false, t->coset_scaled, st));
            G16_TRY(ntt_dit(ctx, c, t, true, nullptr, &t->zinv_n, st));
        }
        //   coset iNTT of a*b (product taken on the first load; g^-i / (n (g^n - 1)) and the subtraction on the last store)
        G16_TRY(ntt_dit(ctx, a, t, true, t->coset_inv_z, nullptr, st, b, c));
    } else if (reduction == G16_REDUCTION_CIRCOM) {
        if (ctx->log_n >= 28) return set_err(ctx, G16_ERR_DEGREE_TOO_LARGE, "circom reduction needs a 2n domain");
        G16_TRY(r1cs_eval_dev(ctx, a, b, nullptr, true, st));
        G16_LAUNCH(ctx, k_pad_rows, pb, 128, 0, st, ctx->d_z, a, b, (Fr*)nullptr, ctx->nc, ctx->ni, (uint64_t)n, lr);
        // c_i = a_i * b_i for i < nc; b is zero on the padding rows so the product is zero there as the reference's c
        G16_LAUNCH(ctx, k_mul_into, eb, 256, 0, st, a, b, c, n);
        Fr* abc[3] = {a, b, c};
        if (ctx->opt_ntt_batch > 0 || (ctx->opt_ntt_batch < 0 && ctx->wm_alone)) {
            G16_TRY(ntt_dit_batch(ctx, abc, 3, t, true, nullptr, nullptr, 0u, st));
            G16_TRY(ntt_dif_batch(ctx, abc, 3, t, false, t->odd_scaled, st));
        } else {
            G16_TRY(ntt_dit(ctx, a, t, true, nullptr, nullptr, st));
            G16_TRY(ntt_dit(ctx, b, t, true, nullptr, nullptr, st));
            G16_TRY(ntt_dit(ctx, c, t, true, nullptr, nullptr, st));
            G16_TRY(ntt_dif(ctx, a, t, false, t->odd_scaled, st));
            G16_TRY(ntt_dif(ctx, b, t, false, t->odd_scaled, st));
            G16_TRY(ntt_dif(ctx, c, t, false, t->odd_scaled, st));
        }
        G16_LAUNCH(ctx, k_mul_sub, eb, 256, 0, st, a, b, c, n);
        G16_TRY(bitrev_permute(ctx, a, ctx->log_n, st));  // evaluations are the output here: natural order needed
    } else {
        return set_err(ctx, G16_ERR_BAD_ARG, "unknown reduction %d", reduction);
    }
    return G16_OK;
}

}  // namespace g16
