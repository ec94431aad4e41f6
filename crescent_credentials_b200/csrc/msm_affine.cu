// msm_affine.cu -- batched-affine bucket accumulation for the Pippenger MSMs (first stage of K7).
//
// Same job as the bucket loop of ark-ec 0.4 VariableBaseMSM::msm_bigint (reference call sites
// forks/groth16/src/prover.rs:66,74,266): add up the points that fall into each bucket.  The result of a bucket is an
// exact group element, so the order and the coordinate system of the additions are free.
//
// An XYZZ mixed addition costs 10 Fq products (8M + 2S).  An affine addition costs 1 inversion + 3 products, and with
// Montgomery's trick the inversion of M independent denominators costs 3(M-1) products plus ONE inversion -- about 6.5
// products per addition once M is large.  Independent additions are what a bucket tree provides: level k adds the points
// of every bucket pairwise (2i, 2i+1), halving each bucket, so level k of ALL buckets is one flat list of independent
// additions.  Per level, three kernels:
//     k_ba_products  thread t walks its K consecutive output slots (K per level, ba_slots_per_thread), multiplies the denominators x2 - x1 into a
//                    running product (stored per slot: the prefix products) and emits its total T[t];
//     k_ba_invert    one warp per 32 * kBaGroup totals: prefix products, a shuffle product tree, ONE binary-GCD inversion,
//                    the tree walked back down, back-substitution => 1/T[t];
//     k_ba_add       thread t walks its slots backwards: 1/d_j = (running inverse) * prefix_{j-1}, then
//                    lambda = (y2 - y1)/d, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1.
// For G2 the shared inversion runs in Fq on the norms: 1/d = conj(d) / (d.c0^2 + d.c1^2).
// After `levels` levels (default 5: 97 % of the additions) the few survivors per bucket are summed by the XYZZ task
// kernel of msm.cu, which also absorbs whatever skew the scalars have.
//
// Exceptional pairs are decided identically in k_ba_products and k_ba_add (ba_classify): an operand at infinity or
// P + (-P) needs no inversion (denominator treated as 1); P + P uses the tangent (d = 2y, numerator 3x^2).
#include "msm_internal.cuh"

namespace g16 {

// ---- the field the shared inversion runs in is always Fq ------------------------------------------------------------------
template <class F>
struct BaField;
template <>
struct BaField<Fq> {
    static __device__ __forceinline__ Fq den(const Fq& d) { return d; }
    static __device__ __forceinline__ Fq inv(const Fq& d, const Fq& inv_den) {
        (void)d;
        return inv_den;
    }
};
template <>
struct BaField<Fq2> {
    static __device__ __forceinline__ Fq den(const Fq2& d) { return d.c0.sqr() + d.c1.sqr(); }
    static __device__ __forceinline__ Fq2 inv(const Fq2& d, const Fq& inv_den) {
        return Fq2{d.c0 * inv_den, (d.c1 * inv_den).neg()};
    }
};

enum { BA_ADD = 0, BA_DBL = 1, BA_TRIVIAL = 2 };

// kind of the addition p1 + p2 and its denominator
template <class F>
__device__ __forceinline__ int ba_classify(const Affine<F>& p1, const Affine<F>& p2, F& d) {
    if (p1.is_inf() || p2.is_inf()) return BA_TRIVIAL;
    d = p2.x - p1.x;
    if (!d.is_zero()) return BA_ADD;
    if (p1.y == p2.y && !p1.y.is_zero()) {
        d = p1.y.dbl();
        return BA_DBL;
    }
    return BA_TRIVIAL;  // P + (-P)
}
template <class F>
__device__ __forceinline__ Affine<F> ba_trivial_sum(const Affine<F>& p1, const Affine<F>& p2) {
    if (p1.is_inf()) return p2;
    if (p2.is_inf()) return p1;
    return Affine<F>::inf();
}

// input points of a level: level 0 gathers from the base table through the sorted references (sign in the top bit)
template <class F, bool L0>
__device__ __forceinline__ Affine<F> ba_load(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ vals, uint32_t s) {
    if (L0) {
        uint32_t ref = vals[s];
        Affine<F> p = ldg_vec(src + (ref & ~kNegBit));
        if (ref & kNegBit) p.y = p.y.neg();
        return p;
    }
    return ld_vec(src + s);
}
// ---- level tables -------------------------------------------------------------------------------------------------------------
// lvl row 0 = level-0 point counts per bucket; rows 1..levels = counts after each level (scanned in place afterwards into
// offsets; entry [nbuckets] of a row then holds the row total).  start0 = first sorted position of every bucket.
__global__ void k_ba_counts(const uint32_t* __restrict__ bucket_start, unsigned nseg, uint32_t nb, int levels,
                            uint32_t* __restrict__ start0, uint32_t* __restrict__ lvl, size_t stride) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)nseg * nb;
    if (g > total) return;
    if (g == total) {
        for (int k = 0; k <= levels; k++) lvl[(size_t)k * stride + g] = 0;
        return;
    }
    unsigned seg = (unsigned)(g / nb);
    uint32_t b = (uint32_t)(g - (size_t)seg * nb);
    const uint32_t* st = bucket_start + (size_t)seg * (nb + 1);
    uint32_t s0 = st[b], c = st[b + 1] - s0;
    start0[g] = s0;
    lvl[g] = c;
    for (int k = 1; k <= levels; k++) {
        c = (c + 1) >> 1;
        lvl[(size_t)k * stride + g] = c;
    }
}

// first / last bucket touched by every thread of the level kernels (blockIdx.y = output level - 1)
struct BaKs {
    int k[kBaMaxLevels];
};
__global__ void k_ba_thread_buckets(const uint32_t* __restrict__ lvl, size_t stride, size_t nbuckets,
                                    uint32_t* __restrict__ tb, size_t tstride, BaKs ks) {
    const unsigned k = blockIdx.y + 1;
    const int kBaK = ks.k[blockIdx.y];
    const uint32_t* off = lvl + (size_t)k * stride;
    const uint32_t total = off[nbuckets];
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t o0 = t * kBaK;
    if (o0 >= total) return;
    uint32_t o1 = (uint32_t)(o0 + kBaK < total ? o0 + kBaK : total) - 1;
    auto find = [&](uint32_t o) {  // largest g with off[g] <= o  (that bucket is non-empty because off[g + 1] > o)
        size_t lo = 0, hi = nbuckets;  // invariant: off[lo] <= o < off[hi]
        while (hi - lo > 1) {
            size_t mid = (lo + hi) >> 1;
            if (off[mid] <= o) lo = mid;
            else hi = mid;
        }
        return (uint32_t)lo;
    };
    uint32_t* row = tb + (size_t)blockIdx.y * 2 * tstride;
    row[t] = find((uint32_t)o0);
    row[tstride + t] = find(o1);
}

// ---- the three kernels of a level ---------------------------------------------------------------------------------------
// Level 0 (L0) reads its points from the base table through the sorted references.  k_ba_products touches only the x
// coordinates (one 32-byte sector per point; staging whole points for k_ba_add was measured slower: it doubles the random
// traffic of this kernel, which is bound by it).
template <class F, bool L0>
__device__ __forceinline__ F ba_load_x(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ vals, uint32_t s) {
    if (L0) return ldg_vec(&src[vals[s] & ~kNegBit].x);
    return ld_vec(&src[s].x);
}

template <class F, bool L0>
__global__ void __launch_bounds__(128)
    k_ba_products(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ in_start,
                  const uint32_t* __restrict__ in_cnt, const uint32_t* __restrict__ out_off, size_t nbuckets,
                  const uint32_t* __restrict__ tb_first, Fq* __restrict__ pre, Fq* __restrict__ T, Fq* __restrict__ den_out,
                  const int kBaK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = out_off[nbuckets];
    const size_t o0 = t * kBaK;
    if (o0 >= total) return;
    uint32_t g = tb_first[t];
    uint32_t g_off = out_off[g], g_end = out_off[g + 1];
    uint32_t s_base = in_start[g];
    uint32_t s_cnt = in_cnt ? in_cnt[g] : in_start[g + 1] - s_base;
    Fq run = Fq::one();
#pragma unroll 1
    for (int j = 0; j < kBaK; j++) {
        uint32_t o = (uint32_t)o0 + j;
        if (o >= total) break;
        while (o >= g_end) {
            g++;
            g_off = g_end;
            g_end = out_off[g + 1];
            s_base = in_start[g];
            s_cnt = in_cnt ? in_cnt[g] : in_start[g + 1] - s_base;
        }
        uint32_t i = o - g_off;
        if (2 * i + 1 < s_cnt) {
            uint32_t s = s_base + 2 * i;
            F x1 = ba_load_x<F, L0>(src, vals, s);
            F x2 = ba_load_x<F, L0>(src, vals, s + 1);
            F d = x2 - x1;
            int kind = BA_ADD;
            if (d.is_zero() || x1.is_zero() || x2.is_zero()) {  // rare: needs the y coordinates to decide
                Affine<F> p1 = ba_load<F, L0>(src, vals, s), p2 = ba_load<F, L0>(src, vals, s + 1);
                kind = ba_classify(p1, p2, d);
            }
            if (kind != BA_TRIVIAL) {
                const Fq den = BaField<F>::den(d);
                // G2: the Fq norm costs two squarings; k_ba_add reads it back instead of recomputing it (G1: den is d itself)
                if (sizeof(F) != sizeof(Fq)) st_vec(den_out + o, den);
                run = run * den;
            }
        }
        st_vec(pre + o, run);
    }
    st_vec(T + t, run);
}

// Q[i] = 1 / T[i] for the nT = ceil(total / kBaK) thread totals of a level.  One warp inverts 32 * kBaGroup totals with a
// single field inversion: every lane multiplies its kBaGroup totals up (prefix products kept in Q), a shuffle tree
// multiplies the 32 lane products, lane 31 inverts the warp product (binary-GCD inversion: this kernel is pure latency),
// the tree is walked back down handing every lane the inverse of its own product, and the lanes back-substitute.
__device__ __forceinline__ Fq shfl_fq(const Fq& a, unsigned src_lane) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src_lane);
    return r;
}
__global__ void __launch_bounds__(128)
    k_ba_invert(const uint32_t* __restrict__ out_off, size_t nbuckets, const Fq* __restrict__ T, Fq* __restrict__ Q, const int kBaK) {
    const uint32_t total = out_off[nbuckets];
    const size_t nT = ((size_t)total + kBaK - 1) / kBaK;
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const size_t lo = u * kBaGroup;
    if ((u - lane) * kBaGroup >= nT) return;  // whole warp beyond the end
    const size_t hi = lo >= nT ? lo : (lo + kBaGroup < nT ? lo + kBaGroup : nT);
    Fq q = Fq::one();
#pragma unroll 1
    for (size_t i = lo; i < hi; i++) {
        q = q * ld_vec(T + i);
        st_vec(Q + i, q);
    }
    // up-sweep: after step s, lanes with the low s+1 bits set hold the product of their 2^(s+1)-lane block
    Fq sub[5];  // sub[s] = product of this lane's block before step s (what the sibling needs on the way down)
    Fq acc = q;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        sub[s] = acc;
        Fq other = shfl_fq(acc, lane ^ (1u << s));  // sibling block's product (both siblings hold their own after step s-1)
        acc = acc * other;                           // every lane of the merged block now holds the merged product
    }
    // acc = product of all 32 lanes (identical in every lane); one inversion per warp
    Fq inv;
    if (lane == 31) inv = acc.inverse_fast();
    inv = shfl_fq(inv, 31);
    // down-sweep: inverse of a block's product = inverse of the merged product * the sibling's product
#pragma unroll
    for (int s = 4; s >= 0; s--) {
        Fq other = shfl_fq(sub[s], lane ^ (1u << s));
        inv = inv * other;
    }
    // inv = 1 / (this lane's product); back-substitute
#pragma unroll 1
    for (size_t i = hi; i-- > lo;) {
        Fq prev = i > lo ? ld_vec(Q + i - 1) : Fq::one();
        Fq ti = ld_vec(T + i);
        st_vec(Q + i, inv * prev);
        inv = inv * ti;
    }
}

static int ba_invert(g16_ctx* ctx, const uint32_t* out_off, size_t nbuckets, size_t threads, Fq* T, Fq* Q, int kBaK, cudaStream_t st) {
    size_t inv_threads = (threads + kBaGroup - 1) / kBaGroup;
    inv_threads = (inv_threads + 31) / 32 * 32;
    G16_LAUNCH(ctx, k_ba_invert, (unsigned)((inv_threads + 127) / 128), 128, 0, st, out_off, nbuckets, (const Fq*)T, Q, kBaK);
    return G16_OK;
}

// p1 <- p1 + p2 for one slot of a level: `prefix` is the product of the denominators of the thread's earlier slots (one for
// the thread's first slot), `run` the running inverse (updated).  den_in: the Fq norms k_ba_products stored (G2 only).
template <class F>
__device__ __forceinline__ void ba_add_pair(Affine<F>& p1, const Affine<F>& p2, const Fq& prefix, bool first, Fq& run,
                                            const Fq* __restrict__ den_in, uint32_t o) {
    F d;
    int kind = ba_classify(p1, p2, d);
    if (kind == BA_TRIVIAL) {
        p1 = ba_trivial_sum(p1, p2);
        return;
    }
    Fq den = sizeof(F) == sizeof(Fq) ? BaField<F>::den(d) : ld_vec(den_in + o);  // G2: the norm k_ba_products stored
    Fq inv_den;
    inv_den = first ? run : run * prefix;
    run = run * den;
    F inv_d = BaField<F>::inv(d, inv_den);
    F num;
    if (kind == BA_ADD) {
        num = p2.y - p1.y;
    } else {
        F xx = p1.x.sqr();
        num = xx.dbl() + xx;
    }
    F lam = num * inv_d;
    F x3 = lam.sqr() - p1.x - p2.x;
    p1.y = lam * (p1.x - x3) - p1.y;
    p1.x = x3;
}

// G1 held to 80 registers (6 blocks per SM, 36 bytes spilled) rather than the 92 ptxas would take (5 blocks): h-query level 0
// 1.537 against 1.554 ms (profiles/r02_final_pass2.log); G2 to 168 (3 blocks)
template <class F, bool L0>
__global__ void __launch_bounds__(128, sizeof(F) == sizeof(Fq) ? 6 : 3)
    k_ba_add(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ in_start,
             const uint32_t* __restrict__ in_cnt, const uint32_t* __restrict__ out_off, size_t nbuckets,
             const uint32_t* __restrict__ tb_last, const Fq* __restrict__ pre, const Fq* __restrict__ Tinv,
             Affine<F>* __restrict__ out, const Fq* __restrict__ den_in, const int kBaK) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = out_off[nbuckets];
    const size_t o0 = t * kBaK;
    if (o0 >= total) return;
    const uint32_t o_last = (uint32_t)(o0 + kBaK < total ? o0 + kBaK : total) - 1;
    uint32_t g = tb_last[t];
    uint32_t g_off = out_off[g];
    uint32_t s_base = in_start[g];
    uint32_t s_cnt = in_cnt ? in_cnt[g] : in_start[g + 1] - s_base;
    Fq run = ld_vec(Tinv + t);
    // Measured and not adopted (profiles/r02_ab_ba_add_ntt_batch.log, h-query level 0 timed alone / whole proof): issuing a
    // slot's loads back to back with the next slot's references prefetched under the gathers -- 1.548 ms / 33.7 ms at 80
    // registers with spills, 1.482 / 33.3 at 96 registers and 5 blocks per SM, against 1.462 / 33.8 for this loop; the two
    // products off the running inverse row-interleaved (four carry chains) -- 1.574 / 34.3.  Level 0 is bound by the random
    // 128-byte DRAM granules of its gathers (379 B of DRAM traffic per slot, ncu), not by exposed latency or by ILP.
#pragma unroll 1
    for (uint32_t o = o_last;; o--) {
        while (o < g_off) {
            g--;
            g_off = out_off[g];
            s_base = in_start[g];
            s_cnt = in_cnt ? in_cnt[g] : in_start[g + 1] - s_base;
        }
        uint32_t i = o - g_off;
        uint32_t s = s_base + 2 * i;
        Affine<F> p1 = ba_load<F, L0>(src, vals, s);
        if (2 * i + 1 < s_cnt) {
            Affine<F> p2 = ba_load<F, L0>(src, vals, s + 1);
            Fq prefix = (o == (uint32_t)o0) ? Fq::one() : ld_vec(pre + o - 1);
            ba_add_pair<F>(p1, p2, prefix, o == (uint32_t)o0, run, den_in, o);
        }
        st_vec(out + o, p1);
        if (o == (uint32_t)o0) break;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
// Output slots per thread at level k (0-based): one shared-inversion chain per thread.  16 at full size (the inversion is then
// amortised over 16 x 256 additions); fewer when the level is small -- a thread walks its slots one after the other, so on a
// level of a few hundred thousand slots 16 per thread meant ~1 wave of blocks each running a 16-deep dependent chain (the
// small shards of an 8-GPU run, small circuits).  Chosen from the level's capacity on the host: the launch sequence stays static.
int ba_slots_per_thread(size_t items, size_t nbuckets, int level) {
    const size_t cap = ba_level_cap(items, nbuckets, level + 1);
    const size_t target_threads = (size_t)kNumSMs * 6 * 128 * 3;  // three waves of resident k_ba_add blocks
    size_t k = (cap + target_threads - 1) / target_threads;
    if (k < 2) k = 2;
    if (k > (size_t)kBaKMax) k = kBaKMax;
    return (int)k;
}

size_t ba_level_cap(size_t items, size_t nbuckets, int level) {
    size_t cap = items;
    for (int k = 0; k < level; k++) cap = (cap + nbuckets) / 2 + 1;
    return cap;
}

void ba_free(MsmScratch* sc) {
    dev_free(sc->ba_start0);
    dev_free(sc->ba_lvl);
    dev_free(sc->ba_tb);
    dev_free(sc->ba_buf_a);
    dev_free(sc->ba_buf_b);
    dev_free(sc->ba_pre);
    dev_free(sc->ba_T);
    dev_free(sc->ba_Q);
    dev_free(sc->ba_den);
    sc->ba_den = nullptr;
    sc->ba_start0 = sc->ba_lvl = sc->ba_tb = nullptr;
    sc->ba_buf_a = sc->ba_buf_b = nullptr;
    sc->ba_pre = sc->ba_T = sc->ba_Q = nullptr;
    sc->ba_levels = 0;
}

int ba_alloc(g16_ctx* ctx, int group, MsmScratch* sc, size_t items, size_t nbuckets, int levels) {
    sc->ba_levels = levels;
    sc->ba_stride = nbuckets + 1;
    // the level tables exist even with levels == 0: row 0 (+ start0) feeds the task builder
    G16_TRY(dev_alloc(ctx, &sc->ba_start0, nbuckets + 1));
    G16_TRY(dev_alloc(ctx, &sc->ba_lvl, (size_t)(levels + 1) * sc->ba_stride));
    if (levels == 0) return G16_OK;
    size_t cap1 = ba_level_cap(items, nbuckets, 1);
    size_t threads = 0;
    for (int k = 0; k < levels; k++) {
        const size_t K = (size_t)ba_slots_per_thread(items, nbuckets, k);
        const size_t th = (ba_level_cap(items, nbuckets, k + 1) + K - 1) / K;
        if (th > threads) threads = th;
    }
    sc->ba_items = items;
    sc->ba_tstride = threads + 1;
    G16_TRY(dev_alloc(ctx, &sc->ba_tb, (size_t)levels * 2 * sc->ba_tstride));
    size_t pb = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    G16_CUDA(ctx, cudaMalloc(&sc->ba_buf_a, cap1 * pb));
    if (levels > 1) G16_CUDA(ctx, cudaMalloc(&sc->ba_buf_b, ba_level_cap(items, nbuckets, 2) * pb));
    Fq* p;
    G16_TRY(dev_alloc(ctx, &p, cap1));
    sc->ba_pre = p;
    G16_TRY(dev_alloc(ctx, &p, threads + 1));
    sc->ba_T = p;
    G16_TRY(dev_alloc(ctx, &p, threads + 1));
    sc->ba_Q = p;
    if (group == 2) {  // per-slot Fq norms of the G2 denominators, written by k_ba_products and read back by k_ba_add
        G16_TRY(dev_alloc(ctx, &p, cap1));
        sc->ba_den = p;
    }
    return G16_OK;
}

int ba_build_levels(g16_ctx* ctx, MsmScratch* dg, unsigned nseg, uint32_t nb, int levels, cudaStream_t st) {
    const size_t nbuckets = (size_t)nseg * nb;
    G16_LAUNCH(ctx, k_ba_counts, (unsigned)((nbuckets + 1 + 255) / 256), 256, 0, st, dg->bucket_start, nseg, nb, levels,
               dg->ba_start0, dg->ba_lvl, dg->ba_stride);
    if (levels == 0) return G16_OK;
    G16_TRY(exclusive_scan_batched(ctx, dg->ba_lvl + dg->ba_stride, nbuckets + 1, dg->ba_stride, (unsigned)levels, dg->task_tmp, st));
    size_t threads = dg->ba_tstride - 1;
    BaKs ks;
    for (int k = 0; k < kBaMaxLevels; k++) ks.k[k] = k < levels ? ba_slots_per_thread(dg->ba_items, nbuckets, k) : kBaKMax;
    G16_LAUNCH(ctx, k_ba_thread_buckets, dim3((unsigned)((threads + 255) / 256), (unsigned)levels), 256, 0, st, dg->ba_lvl,
               dg->ba_stride, nbuckets, dg->ba_tb, dg->ba_tstride, ks);
    return G16_OK;
}

template <class F>
static int ba_run_levels_t(g16_ctx* ctx, const MsmBases* mb, MsmScratch* sc, const MsmScratch* dg, size_t items, size_t nbuckets,
                           int levels, const void** out_pts, cudaStream_t st, cudaEvent_t add0_ev0, cudaEvent_t add0_ev1) {
    const Affine<F>* src = (const Affine<F>*)mb->pts;
    for (int k = 0; k < levels; k++) {
        size_t cap = ba_level_cap(items, nbuckets, k + 1);
        const int kBaK = ba_slots_per_thread(dg->ba_items, nbuckets, k);  // the K the thread -> bucket tables were built with
        size_t threads = (cap + kBaK - 1) / kBaK;
        unsigned grid = (unsigned)((threads + 127) / 128);
        const uint32_t* out_off = dg->ba_lvl + (size_t)(k + 1) * dg->ba_stride;
        const uint32_t* in_start = k == 0 ? dg->ba_start0 : dg->ba_lvl + (size_t)k * dg->ba_stride;
        const uint32_t* in_cnt = k == 0 ? dg->ba_lvl : nullptr;
        const uint32_t* tb_first = dg->ba_tb + (size_t)k * 2 * dg->ba_tstride;
        const uint32_t* tb_last = tb_first + dg->ba_tstride;
        Affine<F>* dst = (Affine<F>*)((k & 1) ? sc->ba_buf_b : sc->ba_buf_a);
        Fq* pre = (Fq*)sc->ba_pre;
        Fq* T = (Fq*)sc->ba_T;
        Fq* Q = (Fq*)sc->ba_Q;
        Fq* den = (Fq*)sc->ba_den;  // G2 only
        if (k == 0) {
            G16_LAUNCH(ctx, (k_ba_products<F, true>), grid, 128, 0, st, src, dg->s_vals, in_start, in_cnt, out_off, nbuckets, tb_first, pre, T,
                       den, kBaK);
            G16_TRY(ba_invert(ctx, out_off, nbuckets, threads, T, Q, kBaK, st));
            if (add0_ev0) G16_CUDA(ctx, cudaEventRecord(add0_ev0, st));
            G16_LAUNCH(ctx, (k_ba_add<F, true>), grid, 128, 0, st, src, dg->s_vals, in_start, in_cnt, out_off, nbuckets, tb_last,
                           (const Fq*)pre, (const Fq*)Q, dst, (const Fq*)den, kBaK);
            if (add0_ev1) G16_CUDA(ctx, cudaEventRecord(add0_ev1, st));
        } else {
            G16_LAUNCH(ctx, (k_ba_products<F, false>), grid, 128, 0, st, src, (const uint32_t*)nullptr, in_start, in_cnt, out_off, nbuckets,
                       tb_first, pre, T, den, kBaK);
            G16_TRY(ba_invert(ctx, out_off, nbuckets, threads, T, Q, kBaK, st));
            G16_LAUNCH(ctx, (k_ba_add<F, false>), grid, 128, 0, st, src, (const uint32_t*)nullptr, in_start, in_cnt, out_off, nbuckets,
                           tb_last, (const Fq*)pre, (const Fq*)Q, dst, (const Fq*)den, kBaK);
        }
        src = dst;
    }
    *out_pts = src;
    return G16_OK;
}

int ba_run_levels(g16_ctx* ctx, const MsmBases* mb, MsmScratch* sc, const MsmScratch* dg, size_t items, size_t nbuckets,
                  int levels, const void** out_pts, cudaStream_t st, cudaEvent_t add0_ev0, cudaEvent_t add0_ev1) {
    if (mb->group == 1) return ba_run_levels_t<Fq>(ctx, mb, sc, dg, items, nbuckets, levels, out_pts, st, add0_ev0, add0_ev1);
    return ba_run_levels_t<Fq2>(ctx, mb, sc, dg, items, nbuckets, levels, out_pts, st, add0_ev0, add0_ev1);
}

}  // namespace g16
