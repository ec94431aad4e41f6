// ntt.cu -- radix-2 NTT / iNTT / coset variants over BN254 Fr (K3).
//
// Replaces ark-poly 0.4 Radix2EvaluationDomain::{fft,ifft}_in_place and the coset domain used by the reference at
// forks/groth16/src/r1cs_to_qap.rs:179-185,198-199,210 (and forks/circom-compat/src/circom/qap.rs:60-83).
// Conventions (same as the oracle): omega_n = rho^(2^(28-log n)), rho = 5^((r-1)/2^28); the inverse transform carries
// 1/n; coset FFT multiplies coefficient i by g^i first, coset iFFT multiplies by g^-i afterwards.
//
// Structure: a transform is 1-3 passes over HBM.  Each pass stages a tile of 2^11 elements (64 KB) in shared memory --
// as two 16-byte half-planes so that unit-stride accesses are bank-conflict free -- and runs up to 11 butterfly levels
// there.  The forward transform is decimation-in-frequency (natural -> bit-reversed), the inverse decimation-in-time
// (bit-reversed -> natural): the witness map chains them so that no bit-reversal pass is ever executed; the
// element-wise coset / 1/n scalings ride on the first load or the last store of a transform.
#include "internal.cuh"

namespace g16 {

constexpr unsigned kTileLog = 11;
constexpr int kNttThreads = 512;

// ---- host-side scalar helpers (fp.cuh compiles for the host too) --------------------------------------------------
static Fr fr_from_canonical_words(const uint32_t* w) {
    Fr x;
    for (int i = 0; i < 8; i++) x.v[i] = w[i];
    return x.to_mont();
}
static Fr fr_from_u64(uint64_t v) {
    uint32_t w[8] = {(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0};
    return fr_from_canonical_words(w);
}
static Fr fr_pow_u64(Fr b, uint64_t e) {
    Fr acc = Fr::one();
    while (e) {
        if (e & 1) acc = acc * b;
        b = b.sqr();
        e >>= 1;
    }
    return acc;
}
// rho = 5^((r-1)/2^28), canonical limbs (re-derived by oracle/pyref.py FR_ROOT_2_28)
static Fr fr_root_2_28() {
    static const uint32_t w[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu,
                                  0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};
    return fr_from_canonical_words(w);
}
static Fr fr_omega(unsigned log_n) {
    Fr w = fr_root_2_28();
    for (unsigned i = log_n; i < 28; i++) w = w.sqr();
    return w;
}

struct PowLadder {
    Fr sq[32];  // base^(2^k)
};

__global__ void k_pow_table(Fr* __restrict__ out, size_t n, PowLadder lad, Fr scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        Fr acc = scale;
        uint32_t e = (uint32_t)i;
#pragma unroll 1
        for (int k = 0; k < 32 && e; k++, e >>= 1)
            if (e & 1u) acc = acc * lad.sq[k];
        out[i] = acc;
    }
}

int pow_table_dev(g16_ctx* ctx, Fr* out, size_t n, Fr base, Fr scale, cudaStream_t st);
static int make_pow_table(g16_ctx* ctx, Fr** out, size_t n, Fr base, Fr scale, cudaStream_t st) {
    G16_TRY(dev_alloc(ctx, out, n));
    return pow_table_dev(ctx, *out, n, base, scale, st);
}
int pow_table_dev(g16_ctx* ctx, Fr* dst, size_t n, Fr base, Fr scale, cudaStream_t st) {
    Fr** out = &dst;
    if (n == 0) return G16_OK;
    PowLadder lad;
    Fr b = base;
    for (int k = 0; k < 32; k++) {
        lad.sq[k] = b;
        b = b.sqr();
    }
    size_t g = (n + 255) / 256;
    if (g > (size_t)kNumSMs * 8) g = (size_t)kNumSMs * 8;
    G16_LAUNCH(ctx, k_pow_table, (int)g, 256, 0, st, *out, n, lad, scale);
    return G16_OK;
}

int ntt_get_tables(g16_ctx* ctx, unsigned log_n, NttTables** out) {
    if (log_n > 28) return set_err(ctx, G16_ERR_DEGREE_TOO_LARGE, "domain 2^%u exceeds Fr two-adicity 28", log_n);
    auto it = ctx->ntt.find(log_n);
    if (it != ctx->ntt.end()) {
        *out = &it->second;
        return G16_OK;
    }
    NttTables t;
    t.log_n = log_n;
    size_t n = (size_t)1 << log_n;
    Fr w = fr_omega(log_n);
    Fr wi = w.inverse();
    Fr g = fr_from_u64(5);  // Fr::GENERATOR (r1cs_to_qap.rs:182)
    Fr gi = g.inverse();
    t.n_inv = fr_from_u64(n).inverse();
    Fr gn = fr_pow_u64(g, n);
    Fr zv = gn - Fr::one();  // evaluate_vanishing_polynomial(g) = g^n - 1 (r1cs_to_qap.rs:201-204)
    t.zinv_ok = !zv.is_zero();
    t.zinv = t.zinv_ok ? zv.inverse() : Fr::zero();
    size_t half = n > 1 ? n / 2 : 1;
    G16_TRY(make_pow_table(ctx, &t.tw, half, w, Fr::one(), ctx->main));
    G16_TRY(make_pow_table(ctx, &t.tw_inv, half, wi, Fr::one(), ctx->main));
    G16_TRY(make_pow_table(ctx, &t.coset, n, g, Fr::one(), ctx->main));
    G16_TRY(make_pow_table(ctx, &t.coset_inv, n, gi, t.n_inv, ctx->main));
    G16_TRY(make_pow_table(ctx, &t.coset_scaled, n, g, t.n_inv, ctx->main));
    if (log_n < 28) G16_TRY(make_pow_table(ctx, &t.odd_scaled, n, fr_omega(log_n + 1), t.n_inv, ctx->main));
    t.zinv_n = t.zinv * t.n_inv;
    if (t.zinv_ok) G16_TRY(make_pow_table(ctx, &t.coset_inv_z, n, gi, t.zinv_n, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    auto res = ctx->ntt.emplace(log_n, t);
    *out = &res.first->second;
    return G16_OK;
}

// Up to three transforms of the same kind run as ONE launch (blockIdx.y = member): the witness map inverts a, b, c together
// and coset-transforms a, b together.  A 2^21 transform is only 1024 tiles for 148 SMs x 3 resident blocks (2.3 waves: the
// last wave is a third full); batched, the same passes run 6.9 and 4.6 waves deep and half of the launches disappear.
struct NttBatch {
    Fr* data[3];
    unsigned scalar_mask;  // bit m: member m gets the uniform post_scalar on the last DIT pass
};

// ---- the shared-memory pass ---------------------------------------------------------------------------------------
// Levels [s_lo, s_lo + nlev) of a 2^log_n transform.  A tile is 2^nlev rows (the index bits being transformed) by
// 2^t_log adjacent columns (low index bits, for contiguous global accesses); element (j, c) lives at tile slot j*T + c.
template <bool DIT>
__global__ void __launch_bounds__(kNttThreads, 2)
    k_ntt_pass(NttBatch batch, const Fr* __restrict__ tw, unsigned log_n, unsigned s_lo, unsigned nlev,
               unsigned t_log, const Fr* __restrict__ pre, const Fr* __restrict__ post, Fr post_scalar,
               int has_post_scalar, const Fr* __restrict__ sub) {
    extern __shared__ uint4 smem[];
    Fr* __restrict__ data = batch.data[blockIdx.y];
    has_post_scalar = has_post_scalar && ((batch.scalar_mask >> blockIdx.y) & 1u);
    const unsigned tile_log = nlev + t_log;
    const unsigned tile = 1u << tile_log;
    uint4* sm_lo = smem;
    uint4* sm_hi = smem + tile;
    const unsigned T = 1u << t_log;
    const size_t blk = blockIdx.x;
    const size_t lo_tiles = (size_t)1 << (s_lo - t_log);
    const size_t lo_base = (blk % lo_tiles) << t_log;
    const size_t hi = blk / lo_tiles;
    const size_t gbase = (hi << (s_lo + nlev)) | lo_base;

    for (unsigned e = threadIdx.x; e < tile; e += kNttThreads) {
        unsigned c = e & (T - 1), j = e >> t_log;
        size_t gi = gbase | ((size_t)j << s_lo) | c;
        const uint4* src = reinterpret_cast<const uint4*>(data + gi);
        uint4 a = src[0], b = src[1];
        if (pre) {
            Fr x, p = pre[gi];
            *reinterpret_cast<uint4*>(&x.v[0]) = a;
            *reinterpret_cast<uint4*>(&x.v[4]) = b;
            x = x * p;
            a = *reinterpret_cast<uint4*>(&x.v[0]);
            b = *reinterpret_cast<uint4*>(&x.v[4]);
        }
        sm_lo[e] = a;
        sm_hi[e] = b;
    }
    __syncthreads();

    const unsigned nb = tile >> 1;
    for (unsigned lv = 0; lv < nlev; lv++) {
        const unsigned q = DIT ? lv : (nlev - 1 - lv);
        const unsigned s = s_lo + q;
        const unsigned tw_shift = log_n - s - 1;
        for (unsigned bf = threadIdx.x; bf < nb; bf += kNttThreads) {
            unsigned c = bf & (T - 1), jb = bf >> t_log;
            unsigned jl = jb & ((1u << q) - 1);
            unsigned j0 = ((jb >> q) << (q + 1)) | jl;
            unsigned i0 = (j0 << t_log) | c;
            unsigned i1 = i0 | (1u << (q + t_log));
            size_t low = ((size_t)jl << s_lo) | (lo_base + c);  // (global index of i0) mod 2^s
            Fr x, y, w;
            *reinterpret_cast<uint4*>(&x.v[0]) = sm_lo[i0];
            *reinterpret_cast<uint4*>(&x.v[4]) = sm_hi[i0];
            *reinterpret_cast<uint4*>(&y.v[0]) = sm_lo[i1];
            *reinterpret_cast<uint4*>(&y.v[4]) = sm_hi[i1];
            const uint4* wp = reinterpret_cast<const uint4*>(tw + (low << tw_shift));
            *reinterpret_cast<uint4*>(&w.v[0]) = __ldg(wp);
            *reinterpret_cast<uint4*>(&w.v[4]) = __ldg(wp + 1);
            Fr u, v;
            if (DIT) {
                Fr t = y * w;
                u = x + t;
                v = x - t;
            } else {
                u = x + y;
                v = (x - y) * w;
            }
            sm_lo[i0] = *reinterpret_cast<uint4*>(&u.v[0]);
            sm_hi[i0] = *reinterpret_cast<uint4*>(&u.v[4]);
            sm_lo[i1] = *reinterpret_cast<uint4*>(&v.v[0]);
            sm_hi[i1] = *reinterpret_cast<uint4*>(&v.v[4]);
        }
        __syncthreads();
    }

    for (unsigned e = threadIdx.x; e < tile; e += kNttThreads) {
        unsigned c = e & (T - 1), j = e >> t_log;
        size_t gi = gbase | ((size_t)j << s_lo) | c;
        uint4 a = sm_lo[e], b = sm_hi[e];
        if (post || has_post_scalar) {
            Fr x;
            *reinterpret_cast<uint4*>(&x.v[0]) = a;
            *reinterpret_cast<uint4*>(&x.v[4]) = b;
            x = x * (post ? post[gi] : post_scalar);
            if (sub) x = x - sub[gi];
            a = *reinterpret_cast<uint4*>(&x.v[0]);
            b = *reinterpret_cast<uint4*>(&x.v[4]);
        }
        uint4* dst = reinterpret_cast<uint4*>(data + gi);
        dst[0] = a;
        dst[1] = b;
    }
}

// ---- the same pass with two butterfly levels per shared-memory round trip ---------------------------------------------
// A thread keeps the four elements whose row indices differ in bits q and q+1 in registers and runs levels q and q+1 on
// them (4 products, as two radix-2 levels would: in a prime field the fourth root of unity is no cheaper than any other
// twiddle).  Halves the shared-memory traffic, the barriers and the index arithmetic of the radix-2 loop; an odd level
// count leaves one radix-2 level.  Results are bit-identical to k_ntt_pass (exact arithmetic, same twiddle table).
constexpr int kNtt4Threads = 256;

__device__ __forceinline__ Fr sm_ld(const uint4* lo, const uint4* hi, unsigned i) {
    Fr x;
    *reinterpret_cast<uint4*>(&x.v[0]) = lo[i];
    *reinterpret_cast<uint4*>(&x.v[4]) = hi[i];
    return x;
}
__device__ __forceinline__ void sm_st(uint4* lo, uint4* hi, unsigned i, const Fr& x) {
    lo[i] = *reinterpret_cast<const uint4*>(&x.v[0]);
    hi[i] = *reinterpret_cast<const uint4*>(&x.v[4]);
}
__device__ __forceinline__ Fr tw_ld(const Fr* p) {
    Fr w;
    const uint4* wp = reinterpret_cast<const uint4*>(p);
    *reinterpret_cast<uint4*>(&w.v[0]) = __ldg(wp);
    *reinterpret_cast<uint4*>(&w.v[4]) = __ldg(wp + 1);
    return w;
}

template <bool DIT>
__global__ void __launch_bounds__(kNtt4Threads, 3)
    k_ntt_pass4(NttBatch batch, const Fr* __restrict__ tw, unsigned log_n, unsigned s_lo, unsigned nlev,
                unsigned t_log, const Fr* __restrict__ pre, const Fr* __restrict__ post, Fr post_scalar,
                int has_post_scalar, const Fr* __restrict__ sub) {
    extern __shared__ uint4 smem[];
    Fr* __restrict__ data = batch.data[blockIdx.y];
    has_post_scalar = has_post_scalar && ((batch.scalar_mask >> blockIdx.y) & 1u);
    const unsigned tile_log = nlev + t_log;
    const unsigned tile = 1u << tile_log;
    uint4* sm_lo = smem;
    uint4* sm_hi = smem + tile;
    const unsigned T = 1u << t_log;
    const size_t blk = blockIdx.x;
    const size_t lo_tiles = (size_t)1 << (s_lo - t_log);
    const size_t lo_base = (blk % lo_tiles) << t_log;
    const size_t hi = blk / lo_tiles;
    const size_t gbase = (hi << (s_lo + nlev)) | lo_base;

    for (unsigned e = threadIdx.x; e < tile; e += kNtt4Threads) {
        unsigned c = e & (T - 1), j = e >> t_log;
        size_t gi = gbase | ((size_t)j << s_lo) | c;
        const uint4* src = reinterpret_cast<const uint4*>(data + gi);
        uint4 a = src[0], b = src[1];
        if (pre) {
            Fr x, p = pre[gi];
            *reinterpret_cast<uint4*>(&x.v[0]) = a;
            *reinterpret_cast<uint4*>(&x.v[4]) = b;
            x = x * p;
            a = *reinterpret_cast<uint4*>(&x.v[0]);
            b = *reinterpret_cast<uint4*>(&x.v[4]);
        }
        sm_lo[e] = a;
        sm_hi[e] = b;
    }
    __syncthreads();

    // one radix-2 level q (the leftover of an odd level count)
    auto radix2 = [&](unsigned q) {
        const unsigned s = s_lo + q;
        const unsigned tw_shift = log_n - s - 1;
        for (unsigned bf = threadIdx.x; bf < (tile >> 1); bf += kNtt4Threads) {
            unsigned c = bf & (T - 1), jb = bf >> t_log;
            unsigned jl = jb & ((1u << q) - 1);
            unsigned j0 = ((jb >> q) << (q + 1)) | jl;
            unsigned i0 = (j0 << t_log) | c;
            unsigned i1 = i0 | (1u << (q + t_log));
            size_t low = ((size_t)jl << s_lo) | (lo_base + c);
            Fr x = sm_ld(sm_lo, sm_hi, i0), y = sm_ld(sm_lo, sm_hi, i1);
            Fr w = tw_ld(tw + (low << tw_shift));
            Fr u, v;
            if (DIT) {
                Fr t = y * w;
                u = x + t;
                v = x - t;
            } else {
                u = x + y;
                v = (x - y) * w;
            }
            sm_st(sm_lo, sm_hi, i0, u);
            sm_st(sm_lo, sm_hi, i1, v);
        }
        __syncthreads();
    };
    // levels q and q + 1 together
    auto radix4 = [&](unsigned q) {
        const unsigned s = s_lo + q;
        const unsigned sh1 = log_n - s - 1, sh2 = log_n - s - 2;
        for (unsigned g = threadIdx.x; g < (tile >> 2); g += kNtt4Threads) {
            unsigned c = g & (T - 1), jb = g >> t_log;
            unsigned jl = jb & ((1u << q) - 1);
            unsigned j00 = ((jb >> q) << (q + 2)) | jl;
            unsigned i00 = (j00 << t_log) | c;
            unsigned b0 = 1u << (q + t_log), b1 = b0 << 1;
            size_t low = ((size_t)jl << s_lo) | (lo_base + c);  // (global index of i00) mod 2^s
            const Fr* w1p = tw + (low << sh1);                          // level s, both pairs
            const Fr* w2ap = tw + (low << sh2);                         // level s + 1, pair (00, 10)
            const Fr* w2bp = tw + ((low + ((size_t)1 << s)) << sh2);    // level s + 1, pair (01, 11)
            if (DIT) {
                Fr w1 = tw_ld(w1p);
                Fr t1 = sm_ld(sm_lo, sm_hi, i00 | b0) * w1;
                Fr t3 = sm_ld(sm_lo, sm_hi, i00 | b0 | b1) * w1;
                Fr e0 = sm_ld(sm_lo, sm_hi, i00), e2 = sm_ld(sm_lo, sm_hi, i00 | b1);
                Fr a0 = e0 + t1, a1 = e0 - t1, a2 = e2 + t3, a3 = e2 - t3;
                Fr u = a2 * tw_ld(w2ap);
                sm_st(sm_lo, sm_hi, i00, a0 + u);
                sm_st(sm_lo, sm_hi, i00 | b1, a0 - u);
                Fr v = a3 * tw_ld(w2bp);
                sm_st(sm_lo, sm_hi, i00 | b0, a1 + v);
                sm_st(sm_lo, sm_hi, i00 | b0 | b1, a1 - v);
            } else {
                Fr e0 = sm_ld(sm_lo, sm_hi, i00), e2 = sm_ld(sm_lo, sm_hi, i00 | b1);
                Fr u0 = e0 + e2;
                Fr d0 = (e0 - e2) * tw_ld(w2ap);
                Fr e1 = sm_ld(sm_lo, sm_hi, i00 | b0), e3 = sm_ld(sm_lo, sm_hi, i00 | b0 | b1);
                Fr u1 = e1 + e3;
                Fr d1 = (e1 - e3) * tw_ld(w2bp);
                Fr w1 = tw_ld(w1p);
                sm_st(sm_lo, sm_hi, i00, u0 + u1);
                sm_st(sm_lo, sm_hi, i00 | b0, (u0 - u1) * w1);
                sm_st(sm_lo, sm_hi, i00 | b1, d0 + d1);
                sm_st(sm_lo, sm_hi, i00 | b0 | b1, (d0 - d1) * w1);
            }
        }
        __syncthreads();
    };
    if (DIT) {  // ascending levels
        unsigned q = 0;
        for (; q + 1 < nlev; q += 2) radix4(q);
        if (q < nlev) radix2(q);
    } else {  // descending levels
        unsigned q = nlev;
        for (; q >= 2; q -= 2) radix4(q - 2);
        if (q == 1) radix2(0);
    }

    for (unsigned e = threadIdx.x; e < tile; e += kNtt4Threads) {
        unsigned c = e & (T - 1), j = e >> t_log;
        size_t gi = gbase | ((size_t)j << s_lo) | c;
        uint4 a = sm_lo[e], b = sm_hi[e];
        if (post || has_post_scalar) {
            Fr x;
            *reinterpret_cast<uint4*>(&x.v[0]) = a;
            *reinterpret_cast<uint4*>(&x.v[4]) = b;
            x = x * (post ? post[gi] : post_scalar);
            if (sub) x = x - sub[gi];
            a = *reinterpret_cast<uint4*>(&x.v[0]);
            b = *reinterpret_cast<uint4*>(&x.v[4]);
        }
        uint4* dst = reinterpret_cast<uint4*>(data + gi);
        dst[0] = a;
        dst[1] = b;
    }
}

// the element-wise part of a transform when there is no butterfly to ride on (n == 1)
__global__ void k_scale_table(Fr* __restrict__ data, const Fr* __restrict__ pre, const Fr* __restrict__ tbl, Fr scalar,
                              int scale, const Fr* __restrict__ sub, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = data[i];
    if (pre) x = x * pre[i];
    if (scale) x = x * (tbl ? tbl[i] : scalar);
    if (sub) x = x - sub[i];
    data[i] = x;
}

struct Pass {
    unsigned s_lo, nlev, t_log;
};

// DIT order (ascending levels); DIF walks the list backwards.
static std::vector<Pass> plan_passes(unsigned log_n) {
    std::vector<Pass> p;
    if (log_n == 0) return p;
    unsigned first = log_n < kTileLog ? log_n : kTileLog;
    p.push_back({0, first, 0});
    unsigned rem = log_n - first;
    if (rem) {
        unsigned np = (rem + kTileLog - 1) / kTileLog;
        unsigned per = (rem + np - 1) / np;
        unsigned s = first;
        while (rem) {
            unsigned nl = rem < per ? rem : per;
            unsigned t = kTileLog - nl;
            if (t > s) t = s;
            p.push_back({s, nl, t});
            s += nl;
            rem -= nl;
        }
    }
    return p;
}

static bool g_smem_attr_set = false;
static int ensure_smem_attr(g16_ctx* ctx) {
    if (g_smem_attr_set) return G16_OK;
    size_t bytes = ((size_t)1 << kTileLog) * 32;
    G16_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    G16_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    G16_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    G16_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    g_smem_attr_set = true;
    return G16_OK;
}

// `count` (<= 3) transforms of the same kind in one launch per pass.  post / pre_mul / post_sub apply to every member (only
// used with count == 1); post_scalar applies to the members selected by scalar_mask.
int ntt_dit_batch(g16_ctx* ctx, Fr* const* data, int count, NttTables* t, bool inverse_root, const Fr* post, const Fr* post_scalar,
                  unsigned scalar_mask, cudaStream_t st, const Fr* pre_mul, const Fr* post_sub) {
    if (post_sub && !post) return set_err(ctx, G16_ERR_BAD_ARG, "ntt_dit: post_sub needs a post table");
    if (count < 1 || count > 3) return set_err(ctx, G16_ERR_BAD_ARG, "ntt batch of %d", count);
    G16_TRY(ensure_smem_attr(ctx));
    auto passes = plan_passes(t->log_n);
    const Fr* tw = inverse_root ? t->tw_inv : t->tw;
    Fr ps = post_scalar ? *post_scalar : Fr::zero();
    NttBatch batch{{data[0], count > 1 ? data[1] : nullptr, count > 2 ? data[2] : nullptr}, scalar_mask};
    if (passes.empty()) {  // n == 1: only the element-wise work remains
        for (int m = 0; m < count; m++)
            G16_LAUNCH(ctx, k_scale_table, 1, 32, 0, st, data[m], pre_mul, post, ps,
                       (int)(post != nullptr || (post_scalar != nullptr && ((scalar_mask >> m) & 1u))), post_sub, (size_t)1);
        return G16_OK;
    }
    for (size_t k = 0; k < passes.size(); k++) {
        const Pass& p = passes[k];
        bool last = (k + 1 == passes.size());
        size_t blocks = ((size_t)1 << t->log_n) >> (p.nlev + p.t_log);
        size_t smem = ((size_t)1 << (p.nlev + p.t_log)) * 32;
        dim3 grid((unsigned)blocks, (unsigned)count);
        if (ctx->opt_ntt_radix4 >= 0 ? ctx->opt_ntt_radix4 != 0 : ctx->wm_alone)
            G16_LAUNCH(ctx, k_ntt_pass4<true>, grid, kNtt4Threads, smem, st, batch, tw, t->log_n, p.s_lo, p.nlev,
                       p.t_log, k == 0 ? pre_mul : (const Fr*)nullptr, last ? post : (const Fr*)nullptr, ps,
                       (int)(last && post_scalar != nullptr && post == nullptr), last ? post_sub : (const Fr*)nullptr);
        else
            G16_LAUNCH(ctx, k_ntt_pass<true>, grid, kNttThreads, smem, st, batch, tw, t->log_n, p.s_lo, p.nlev,
                       p.t_log, k == 0 ? pre_mul : (const Fr*)nullptr, last ? post : (const Fr*)nullptr, ps,
                       (int)(last && post_scalar != nullptr && post == nullptr), last ? post_sub : (const Fr*)nullptr);
    }
    return G16_OK;
}
int ntt_dit(g16_ctx* ctx, Fr* data, NttTables* t, bool inverse_root, const Fr* post, const Fr* post_scalar,
            cudaStream_t st, const Fr* pre_mul, const Fr* post_sub) {
    Fr* one[1] = {data};
    return ntt_dit_batch(ctx, one, 1, t, inverse_root, post, post_scalar, 1u, st, pre_mul, post_sub);
}

int ntt_dif_batch(g16_ctx* ctx, Fr* const* data, int count, NttTables* t, bool inverse_root, const Fr* pre, cudaStream_t st) {
    if (count < 1 || count > 3) return set_err(ctx, G16_ERR_BAD_ARG, "ntt batch of %d", count);
    G16_TRY(ensure_smem_attr(ctx));
    auto passes = plan_passes(t->log_n);
    const Fr* tw = inverse_root ? t->tw_inv : t->tw;
    NttBatch batch{{data[0], count > 1 ? data[1] : nullptr, count > 2 ? data[2] : nullptr}, 0u};
    // n == 1: no pass, and every pre table starts with 1 (g^0 / n): nothing to do
    for (size_t k = passes.size(); k-- > 0;) {
        const Pass& p = passes[k];
        bool first = (k + 1 == passes.size());
        size_t blocks = ((size_t)1 << t->log_n) >> (p.nlev + p.t_log);
        size_t smem = ((size_t)1 << (p.nlev + p.t_log)) * 32;
        dim3 grid((unsigned)blocks, (unsigned)count);
        if (ctx->opt_ntt_radix4 >= 0 ? ctx->opt_ntt_radix4 != 0 : ctx->wm_alone)
            G16_LAUNCH(ctx, k_ntt_pass4<false>, grid, kNtt4Threads, smem, st, batch, tw, t->log_n, p.s_lo, p.nlev,
                       p.t_log, first ? pre : (const Fr*)nullptr, (const Fr*)nullptr, Fr::zero(), 0, (const Fr*)nullptr);
        else
            G16_LAUNCH(ctx, k_ntt_pass<false>, grid, kNttThreads, smem, st, batch, tw, t->log_n, p.s_lo, p.nlev,
                       p.t_log, first ? pre : (const Fr*)nullptr, (const Fr*)nullptr, Fr::zero(), 0, (const Fr*)nullptr);
    }
    return G16_OK;
}
int ntt_dif(g16_ctx* ctx, Fr* data, NttTables* t, bool inverse_root, const Fr* pre, cudaStream_t st) {
    Fr* one[1] = {data};
    return ntt_dif_batch(ctx, one, 1, t, inverse_root, pre, st);
}

__global__ void k_bitrev(Fr* __restrict__ data, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        size_t r = (size_t)(__brev((unsigned)i) >> (32 - log_n));
        if (i < r) {
            uint4* pi = reinterpret_cast<uint4*>(data + i);
            uint4* pr = reinterpret_cast<uint4*>(data + r);
            uint4 a0 = pi[0], a1 = pi[1], b0 = pr[0], b1 = pr[1];
            pi[0] = b0;
            pi[1] = b1;
            pr[0] = a0;
            pr[1] = a1;
        }
    }
}

int bitrev_permute(g16_ctx* ctx, Fr* data, unsigned log_n, cudaStream_t st) {
    if (log_n == 0) return G16_OK;
    size_t n = (size_t)1 << log_n;
    size_t g = (n + 255) / 256;
    if (g > (size_t)kNumSMs * 16) g = (size_t)kNumSMs * 16;
    G16_LAUNCH(ctx, k_bitrev, (unsigned)g, 256, 0, st, data, log_n);
    return G16_OK;
}

// arkworks-semantics transform, natural order in and out
int ntt_api(g16_ctx* ctx, Fr* d, unsigned log_n, int inverse, int coset, cudaStream_t st) {
    NttTables* t;
    G16_TRY(ntt_get_tables(ctx, log_n, &t));
    if (log_n == 0) return G16_OK;  // size-1 transform is the identity (g^0 = 1, 1/n = 1)
    if (!inverse) {
        G16_TRY(ntt_dif(ctx, d, t, false, coset ? t->coset : nullptr, st));
        G16_TRY(bitrev_permute(ctx, d, log_n, st));
    } else {
        G16_TRY(bitrev_permute(ctx, d, log_n, st));
        G16_TRY(ntt_dit(ctx, d, t, true, coset ? t->coset_inv : nullptr, coset ? nullptr : &t->n_inv, st));
    }
    return G16_OK;
}

}  // namespace g16
