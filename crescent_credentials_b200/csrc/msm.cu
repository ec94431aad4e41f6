// msm.cu -- Pippenger multi-scalar multiplication for BN254 G1 (Fq) and G2 (Fq2)  (K5-K8).
//
// Replaces ark-ec 0.4 VariableBaseMSM::msm_bigint at the reference call sites forks/groth16/src/prover.rs:66,74,266.
// The result is the exact group element Sigma s_i * P_i, so it equals the reference's after normalisation whatever
// the evaluation order.
//
// Pipeline (all on the device, stream-ordered, no host synchronisation):
//   digit stage (depends on the scalars only; MSMs over the same scalars share it -- a/l and b_g1/b_g2 of a proof):
//   1. k_digits        scalar (Montgomery) -> canonical -> W signed c-bit digits; emits (bucket key, point ref) pairs.
//                      Zero digits and points at infinity get the EMPTY key.
//   2. radix sort      msm_sort.cu: LSD radix sort of the pairs by bucket.
//   3. k_bucket_bounds bucket start offsets from the sorted keys (no atomics).
//   4. level tables    msm_affine.cu: per-bucket counts/offsets after each batched-affine level.
//   5. tasks           what is left of every bucket after the affine levels is cut into tasks of <= kTaskLen points;
//                      tasks are radix-sorted by length (descending) so that the lanes of a warp run equally long loops
//                      whatever the scalar skew.
//   point stage (per MSM):
//   6. batched-affine levels (msm_affine.cu): pairwise affine additions with shared inversions, ~6.5 Fq-mul each.
//   7. k_accumulate    one thread per task: XYZZ mixed additions (8M + 2S each), next point prefetched.
//   8. k_bucket_combine one warp per bucket folds that bucket's task partials (shuffle tree when there are several).
//   9. k_chunk_reduce / k_tree_sum / k_finish
//                      Sigma (b+1) * B_b per bucket set by chunked running sums + short double-and-add, tree sums,
//                      and (without precomputation) the Horner combination of the windows.
//
// With `precomp` the bases hold 2^(c*w) * P_i for every window w, all windows share ONE bucket set and step 9's
// Horner tail disappears: HBM capacity (180 GB) is traded for integer-pipe work.
#include <stdlib.h>

#include "msm_internal.cuh"

namespace g16 {

// ---- 1. signed-digit decomposition ---------------------------------------------------------------------------------------
// pair position = w * n + i.  value = base index | sign; base index = i (plain) or w * n_bases + i (precomputed table).
__global__ void k_digits(const Fr* __restrict__ scalars, size_t n, size_t n_bases, unsigned c, unsigned windows,
                         int precomp, const uint8_t* __restrict__ skip, uint32_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ldg_vec(scalars + i).from_mont();
    const uint32_t nb = 1u << (c - 1);
    const bool dead = skip && skip[i];
    uint32_t carry = 0;
    for (unsigned w = 0; w < windows; w++) {
        unsigned bit = w * c;
        unsigned limb = bit >> 5, off = bit & 31;
        uint64_t two = limb < 8 ? (uint64_t)s.v[limb] : 0;
        if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
        uint32_t d = (uint32_t)((two >> off) & ((1ull << c) - 1)) + carry;
        uint32_t neg = 0;
        if (d > nb) {
            d = (1u << c) - d;
            neg = kNegBit;
            carry = 1;
        } else {
            carry = 0;
        }
        size_t pos = (size_t)w * n + i;
        keys[pos] = (d == 0 || dead) ? nb : d - 1;
        vals[pos] = (uint32_t)(precomp ? (size_t)w * n_bases + i : i) | neg;
    }
}

// ---- 3. bucket boundaries -------------------------------------------------------------------------------------------------
// start[seg][b] for b in [0, nb]: first sorted position (relative to the whole array) whose key >= b
__global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, size_t seg_len, unsigned nseg, uint32_t nb,
                                uint32_t* __restrict__ start) {
    size_t total = seg_len * nseg;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    unsigned seg = (unsigned)(g / seg_len);
    size_t i = g - (size_t)seg * seg_len;
    uint32_t k = keys[g];
    int64_t kp = (i == 0) ? -1 : (int64_t)keys[g - 1];
    uint32_t* st = start + (size_t)seg * (nb + 1);
    for (int64_t b = kp + 1; b <= (int64_t)k; b++) st[b] = (uint32_t)g;
    if (i == seg_len - 1)
        for (int64_t b = (int64_t)k + 1; b <= (int64_t)nb; b++) st[b] = (uint32_t)(g + 1);
}

// ---- 5. tasks ----------------------------------------------------------------------------------------------------------------
// Bucket g holds cnt points at positions [start[g], start[g] + cnt) of the point list the accumulate kernel reads
// (cnt == nullptr: start is an offset table with nbuckets + 1 entries).
__device__ __forceinline__ uint32_t bucket_count(const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt, size_t g) {
    return cnt ? cnt[g] : start[g + 1] - start[g];
}
// buckets with at most `direct_max` points get no task: k_bucket_tail sums them with one thread each
__global__ void k_task_counts(const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt, size_t nbuckets,
                              uint32_t direct_max, uint32_t task_len, uint32_t* __restrict__ ntasks) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nbuckets) return;
    uint32_t c = bucket_count(start, cnt, g);
    ntasks[g] = c > direct_max ? (c + task_len - 1) / task_len : 0;
}
__global__ void k_fill_u32(uint32_t* __restrict__ p, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
// task t of bucket g covers positions [start + j*task_len, ...); sort key = task_len - len (longest first)
__global__ void k_make_tasks(const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt_tab,
                             const uint32_t* __restrict__ task_off, size_t nbuckets, uint32_t direct_max, uint32_t task_len,
                             uint32_t* __restrict__ tkeys, uint32_t* __restrict__ tvals, uint32_t* __restrict__ t_start,
                             uint32_t* __restrict__ t_len) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nbuckets) return;
    uint32_t s0 = start[g], cnt = bucket_count(start, cnt_tab, g);
    if (cnt <= direct_max) return;
    uint32_t t = task_off[g];
    for (uint32_t o = 0; o < cnt; o += task_len, t++) {
        uint32_t len = cnt - o < task_len ? cnt - o : task_len;
        t_start[t] = s0 + o;
        t_len[t] = len;
        tkeys[t] = task_len - len;
        tvals[t] = t;
    }
}

// ---- 7. bucket accumulation (XYZZ) ------------------------------------------------------------------------------------------
// DIRECT = false: points are gathered from the base table through the sorted references (sign in the top bit);
// DIRECT = true: the task ranges index `bases` itself (the survivors of the batched-affine levels, sign already applied).
template <class F, int MINB, bool DIRECT>
__global__ void __launch_bounds__(128, MINB)
    k_accumulate(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ vals,
                 const uint32_t* __restrict__ order_keys, const uint32_t* __restrict__ order,
                 const uint32_t* __restrict__ t_start, const uint32_t* __restrict__ t_len, XYZZ<F>* __restrict__ partial,
                 size_t tcap) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < tcap; p += stride) {
        if (order_keys[p] >= kTaskLen) break;  // padding entries sort last
        uint32_t t = order[p];
        uint32_t s0 = t_start[t], len = t_len[t];
        XYZZ<F> acc = XYZZ<F>::inf();
        uint32_t ref = DIRECT ? s0 : vals[s0];
        Affine<F> nxt = ldg_vec(bases + (ref & ~kNegBit));
        for (uint32_t k = 0; k < len; k++) {
            Affine<F> cur = nxt;
            uint32_t cref = ref;
            if (k + 1 < len) {
                ref = DIRECT ? s0 + k + 1 : vals[s0 + k + 1];
                nxt = ldg_vec(bases + (ref & ~kNegBit));
            }
            if (!DIRECT && (cref & kNegBit)) cur.y = cur.y.neg();
            acc.madd(cur);
        }
        st_vec(partial + t, acc);
    }
}

// Survivors of the batched-affine levels: almost every bucket holds one or two points, so one thread sums a whole
// bucket straight into bucket_sum (consecutive buckets are consecutive in memory: no task indirection, no combine).
// Buckets above kTailDirect points (skewed scalars) are left to the task path.
template <class F>
__global__ void __launch_bounds__(128)
    k_bucket_tail(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ off, size_t nbuckets,
                  XYZZ<F>* __restrict__ bucket_sum) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nbuckets) return;
    uint32_t s0 = off[g], cnt = off[g + 1] - s0;
    if (cnt > kTailDirect) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = 0; k < cnt; k++) acc.madd(ld_vec(pts + s0 + k));
    st_vec(bucket_sum + g, acc);
}

// ---- 8. per-bucket combine (one warp per bucket) ----------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ XYZZ<F> shfl_down_xyzz(const XYZZ<F>& a, unsigned delta) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&a);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(XYZZ<F>) / 4; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
    return r;
}

template <class F>
__global__ void __launch_bounds__(128)
    k_bucket_combine(const XYZZ<F>* __restrict__ partial, const uint32_t* __restrict__ task_off,
                     const uint32_t* __restrict__ total_tasks, size_t nbuckets, int keep_untasked,
                     XYZZ<F>* __restrict__ bucket_sum) {
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned lane = threadIdx.x & 31;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b < nbuckets; b += nwarps) {
        uint32_t t0 = task_off[b];
        uint32_t t1 = (b + 1 < nbuckets) ? task_off[b + 1] : *total_tasks;
        uint32_t cnt = t1 - t0;
        if (cnt == 0 && keep_untasked) continue;  // k_bucket_tail already wrote this bucket
        if (cnt <= 1) {
            if (lane == 0) {
                XYZZ<F> v = cnt ? ld_vec(partial + t0) : XYZZ<F>::inf();
                st_vec(bucket_sum + b, v);
            }
            continue;
        }
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t t = t0 + lane; t < t1; t += 32) acc.add(ld_vec(partial + t));
#pragma unroll 1
        for (unsigned d = 16; d >= 1; d >>= 1) {
            XYZZ<F> o = shfl_down_xyzz(acc, d);
            if (lane < d && d < cnt) acc.add(o);
        }
        if (lane == 0) st_vec(bucket_sum + b, acc);
    }
}

// ---- 9. bucket-set reduction:  S = Sigma_{b} (b + 1) * B_b ------------------------------------------------------------------
// chunk k of a set holds buckets [k*kChunk, (k+1)*kChunk): running sums give acc = Sigma (j+1) B_{lo+j}, run = Sigma B;
// contribution = acc + lo * run with lo = k*kChunk applied by a short double-and-add.
template <class F>
__global__ void __launch_bounds__(64)
    k_chunk_reduce(const XYZZ<F>* __restrict__ bucket_sum, unsigned nseg, uint32_t nb, XYZZ<F>* __restrict__ chunk_sum) {
    uint32_t chunks_per_seg = (nb + kChunk - 1) / kChunk;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)nseg * chunks_per_seg) return;
    unsigned seg = (unsigned)(g / chunks_per_seg);
    uint32_t k = (uint32_t)(g - (size_t)seg * chunks_per_seg);
    uint32_t lo = k * kChunk;
    uint32_t hi = lo + kChunk < nb ? lo + kChunk : nb;
    const XYZZ<F>* B = bucket_sum + (size_t)seg * nb;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    // This kernel is one dependent chain per thread (pure latency, ~0.5 ms per G1 MSM and 1.25 ms for G2 in round 1): the G1
    // build inlines the additions -- as calls every one of them moves two 128-byte points through local memory -- and the
    // next bucket is loaded while the current addition runs.  (G2 keeps the calls: inlined it would be ~100 KB of SASS.)
    constexpr bool INL = sizeof(F) == sizeof(Fq);
    XYZZ<F> nxt = ld_vec(B + hi - 1);
#pragma unroll 1
    for (uint32_t j = hi; j-- > lo;) {
        XYZZ<F> cur = nxt;
        if (j > lo) nxt = ld_vec(B + j - 1);
        if (INL) {
            run.add_inl(cur);
            acc.add_inl(run);
        } else {
            run.add(cur);
            acc.add(run);
        }
    }
    if (lo && !run.is_inf()) {
        XYZZ<F> t = XYZZ<F>::inf();
#pragma unroll 1
        for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
            t = INL ? t.dbl_inl() : t.dbl();
            if ((lo >> bit) & 1u) {
                if (INL) t.add_inl(run);
                else t.add(run);
            }
        }
        if (INL) acc.add_inl(t);
        else acc.add(t);
    }
    st_vec(chunk_sum + g, acc);
}

// sums `count` consecutive values per segment with `nblk` blocks per segment: out[seg][blk]
template <class F>
__global__ void __launch_bounds__(128)
    k_tree_sum(const XYZZ<F>* __restrict__ in, size_t count, unsigned nblk, XYZZ<F>* __restrict__ out) {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    unsigned seg = blockIdx.x / nblk, blk = blockIdx.x % nblk;
    const XYZZ<F>* src = in + (size_t)seg * count;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (size_t i = (size_t)blk * blockDim.x + threadIdx.x; i < count; i += (size_t)nblk * blockDim.x) acc.add(ld_vec(src + i));
    st_vec(sh + threadIdx.x, acc);
    __syncthreads();
    for (unsigned d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if (threadIdx.x < d) {
            acc.add(ld_vec(sh + threadIdx.x + d));
            st_vec(sh + threadIdx.x, acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_vec(out + blockIdx.x, acc);
}

// Horner over the window sums: result = Sigma_w 2^(c*w) * S_w  (single bucket set: result = S_0)
template <class F>
__global__ void k_finish(const XYZZ<F>* __restrict__ set_sum, unsigned nseg, unsigned c, XYZZ<F>* __restrict__ result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = ld_vec(set_sum + (nseg - 1));
    for (int w = (int)nseg - 2; w >= 0; w--) {
        for (unsigned k = 0; k < c; k++) acc = acc.dbl();
        acc.add(ld_vec(set_sum + w));
    }
    st_vec(result, acc);
}

// ---- base preparation -------------------------------------------------------------------------------------------------------
template <class F>
__global__ void k_mark_inf(const Affine<F>* __restrict__ pts, size_t n, uint8_t* __restrict__ skip) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) skip[i] = ldg_vec(pts + i).is_inf() ? 1 : 0;
}
// table[w*n + i] = 2^(c*w) * P_i, affine
template <class F>
__global__ void __launch_bounds__(128) k_precomp(Affine<F>* __restrict__ table, size_t n, unsigned c, unsigned windows) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> p = XYZZ<F>::from_affine(ld_vec(table + i));
    for (unsigned w = 1; w < windows; w++) {
        for (unsigned k = 0; k < c; k++) p = p.dbl();
        Affine<F> a = p.to_affine();
        st_vec(table + (size_t)w * n + i, a);
        p = XYZZ<F>::from_affine(a);
    }
}

// generator multiples: out[i] = k_i * G  (canonical generators, forks/groth16/src/generator.rs:34-35)
template <class F>
__global__ void __launch_bounds__(128)
    k_fixed_base(const Fr* __restrict__ scalars, size_t n, Affine<F> gen, Affine<F>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ldg_vec(scalars + i).from_mont();
    XYZZ<F> r = scalar_mul(XYZZ<F>::from_affine(gen), s.v);
    st_vec(out + i, r.to_affine());
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
int msm_pick_window(size_t n, int group, bool precomp) {
    (void)group;
    unsigned lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c;
    if (!precomp) {
        c = (int)lg - 5;
        if (c > 16) c = 16;
    } else if (lg < 17) {
        c = (int)lg - 1;
    } else {
        // with the window tables every digit of every point lands in ONE set of 2^(c-1) buckets, so a smaller window buys
        // fuller buckets (the pairwise levels halve them more effectively, the bucket reduction shrinks) for ~1/c more
        // (point, window) pairs.  Measured best at c = round(log2 n) - 3 from 2^17.6 to 2^21.5 points
        // (profiles/r02_ab_shard_window*.log, r02_ab_window_n1.log: S-rs256 28.3 -> 26.3 ms, one rank of 8 5.42 -> 5.17 ms)
        unsigned r = lg + ((double)n >= 1.41421356 * (double)((size_t)1 << lg) ? 1 : 0);  // log2 n, rounded
        c = (int)r - 3;
        if (c > 20) c = 20;
    }
    if (c < 4) c = 4;
    return c;
}

// XYZZ-only path: every non-empty bucket is cut into tasks of <= kTaskLen points.  After batched-affine levels only the
// buckets that still hold more than kTailDirect points (skewed scalars; the short top window) get tasks, of kTaskLenDirect
// points so that their few threads do not become a latency tail.
constexpr uint32_t kTaskLenDirect = 16;
static size_t task_capacity(size_t items, size_t nbuckets, int levels) {
    if (levels == 0) return items / kTaskLen + nbuckets + 1;
    size_t pts = ba_level_cap(items, nbuckets, levels);
    return pts / kTaskLenDirect + pts / kTailDirect + 2;
}

template <class F>
static int msm_alloc_scratch(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc) {
    const size_t n = mb->n;
    const unsigned W = mb->windows;
    const uint32_t nb = 1u << (mb->c - 1);
    const unsigned nseg = mb->precomp ? 1 : W;
    size_t items = n * W;
    size_t nbuckets = (size_t)nseg * nb;
    size_t tcap = task_capacity(items, nbuckets, mb->ba_levels);
    size_t sort_n = items > tcap ? items : tcap;
    size_t seg_len = mb->precomp ? items : n;
    size_t hist_n = radix_hist_words(seg_len, nseg);
    if (radix_hist_words(tcap, 1) > hist_n) hist_n = radix_hist_words(tcap, 1);
    sc->cap_items = items;
    sc->cap_buckets = nbuckets;
    sc->cap_tasks = tcap;
    sc->hist_cap = hist_n;
    G16_TRY(dev_alloc(ctx, &sc->keys_a, sort_n));
    G16_TRY(dev_alloc(ctx, &sc->keys_b, sort_n));
    G16_TRY(dev_alloc(ctx, &sc->vals_a, sort_n));
    G16_TRY(dev_alloc(ctx, &sc->vals_b, sort_n));
    G16_TRY(dev_alloc(ctx, &sc->hist, hist_n));
    G16_TRY(dev_alloc(ctx, &sc->bucket_start, (size_t)nseg * (nb + 1)));
    G16_TRY(dev_alloc(ctx, &sc->task_off, nbuckets + 1));
    size_t scan_tmp = scan_tmp_words(hist_n);
    if ((size_t)(mb->ba_levels + 1) * scan_tmp_words(nbuckets + 1) > scan_tmp) scan_tmp = (size_t)(mb->ba_levels + 1) * scan_tmp_words(nbuckets + 1);
    G16_TRY(dev_alloc(ctx, &sc->task_tmp, scan_tmp));
    G16_TRY(dev_alloc(ctx, &sc->tasks, 4 * tcap));  // t_start | t_len | task-sort alternates (keys, order)
    G16_TRY(dev_alloc(ctx, &sc->counters, 16));
    G16_TRY(ba_alloc(ctx, mb->group, sc, items, nbuckets, mb->ba_levels));
    size_t chunks = (size_t)nseg * ((nb + kChunk - 1) / kChunk);
    XYZZ<F>* p;
    G16_TRY(dev_alloc(ctx, &p, tcap));
    sc->partial = p;
    G16_TRY(dev_alloc(ctx, &p, nbuckets));
    sc->bucket_sum = p;
    G16_TRY(dev_alloc(ctx, &p, chunks));
    sc->chunk_sum = p;
    G16_TRY(dev_alloc(ctx, &p, (size_t)nseg * 64 + nseg));
    sc->block_sum = p;
    G16_TRY(dev_alloc(ctx, &p, 1));
    sc->result = p;
    sc->point_bytes = sizeof(Affine<F>);
    return G16_OK;
}

void msm_free(MsmBases* mb, MsmScratch* sc) {
    if (mb->owns_pts) dev_free(mb->pts);
    dev_free(mb->skip);
    *mb = MsmBases();
    dev_free(sc->keys_a);
    dev_free(sc->keys_b);
    dev_free(sc->vals_a);
    dev_free(sc->vals_b);
    dev_free(sc->hist);
    dev_free(sc->bucket_start);
    dev_free(sc->task_off);
    dev_free(sc->task_tmp);
    dev_free(sc->tasks);
    dev_free(sc->counters);
    ba_free(sc);
    dev_free(sc->partial);
    dev_free(sc->bucket_sum);
    dev_free(sc->chunk_sum);
    dev_free(sc->block_sum);
    dev_free(sc->result);
    *sc = MsmScratch();
}

template <class F>
static int msm_set_bases_t(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, const void* pts_dev, size_t n, int c, bool precomp,
                           cudaStream_t st) {
    msm_free(mb, sc);
    mb->group = sizeof(F) == sizeof(Fq) ? 1 : 2;
    mb->n = n;
    mb->c = c > 0 ? c : msm_pick_window(n, mb->group, precomp);
    if (mb->c < 2 || mb->c > 24) return set_err(ctx, G16_ERR_BAD_ARG, "window bits %d out of range", mb->c);
    mb->windows = (255 + mb->c - 1) / mb->c;
    mb->precomp = precomp;
    mb->ba_levels = ctx->opt_ba_levels < 0 ? kBaDefaultLevels : ctx->opt_ba_levels;
    if (mb->ba_levels > kBaMaxLevels) mb->ba_levels = kBaMaxLevels;
    if ((size_t)mb->windows * (n ? n : 1) >= ((size_t)1 << 31))
        return set_err(ctx, G16_ERR_BAD_ARG, "MSM of %zu points x %d windows exceeds 2^31 pair references", n, mb->windows);
    if (n == 0) return G16_OK;
    size_t copies = precomp ? mb->windows : 1;
    Affine<F>* tbl;
    G16_TRY(dev_alloc(ctx, &tbl, n * copies));
    mb->pts = tbl;
    mb->owns_pts = true;
    G16_CUDA(ctx, cudaMemcpyAsync(tbl, pts_dev, n * sizeof(Affine<F>), cudaMemcpyDeviceToDevice, st));
    G16_TRY(dev_alloc(ctx, &mb->skip, n));
    G16_LAUNCH(ctx, k_mark_inf<F>, (unsigned)((n + 255) / 256), 256, 0, st, tbl, n, mb->skip);
    if (precomp) G16_LAUNCH(ctx, k_precomp<F>, (unsigned)((n + 127) / 128), 128, 0, st, tbl, n, (unsigned)mb->c, (unsigned)mb->windows);
    G16_TRY(msm_alloc_scratch<F>(ctx, mb, sc));
    return G16_OK;
}

int msm_set_bases(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, int group, const void* pts_dev, size_t n, int c, bool precomp,
                  cudaStream_t st) {
    if (group == 1) return msm_set_bases_t<Fq>(ctx, mb, sc, pts_dev, n, c, precomp, st);
    if (group == 2) return msm_set_bases_t<Fq2>(ctx, mb, sc, pts_dev, n, c, precomp, st);
    return set_err(ctx, G16_ERR_BAD_ARG, "group must be 1 or 2");
}

// counts[0] = points a skips that b does not; counts[1] = points b skips that a does not
__global__ void k_skip_diff(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, unsigned long long* counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long only_a = 0, only_b = 0;
    for (; i < n; i += stride) {
        only_a += a[i] && !b[i];
        only_b += b[i] && !a[i];
    }
    if (only_a) atomicAdd(counts, only_a);
    if (only_b) atomicAdd(counts + 1, only_b);
}

int msm_can_share(g16_ctx* ctx, const MsmBases* a, const MsmBases* b, size_t max_extra_inf, bool* ok, cudaStream_t st) {
    *ok = false;
    if (a->n == 0 || a->n != b->n || a->c != b->c || a->windows != b->windows || a->precomp != b->precomp ||
        a->ba_levels != b->ba_levels)
        return G16_OK;
    unsigned long long* d = (unsigned long long*)((char*)ctx->d_small + 3072);  // past the assembly scratch
    G16_CUDA(ctx, cudaMemsetAsync(d, 0, 16, st));
    G16_LAUNCH(ctx, k_skip_diff, kNumSMs * 2, 256, 0, st, a->skip, b->skip, a->n, d);
    unsigned long long h[2];
    G16_CUDA(ctx, cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, st));
    G16_CUDA(ctx, cudaStreamSynchronize(st));
    // a point only `a` skips would be lost for `b`; a point only `b` skips is an infinity entry of b's table that the
    // addition kernels pass over (costs an idle lane, so only a few are tolerated)
    *ok = h[0] == 0 && h[1] <= max_extra_inf;
    return G16_OK;
}

// ---- digit stage: steps 1-5 ------------------------------------------------------------------------------------------------
static int msm_digit_stage(g16_ctx* ctx, const MsmBases* mb, MsmScratch* sc, const Fr* scalars, size_t n, cudaStream_t st) {
    const unsigned W = mb->windows, c = mb->c;
    const uint32_t nb = 1u << (c - 1);
    const unsigned nseg = mb->precomp ? 1 : W;
    const size_t items = n * W;
    const size_t seg_len = mb->precomp ? items : n;
    const size_t nbuckets = (size_t)nseg * nb;
    const int L = mb->ba_levels;
    const size_t tcap = task_capacity(items, nbuckets, L);
    unsigned key_bits = c;  // keys in [0, nb] need c bits (nb = 2^(c-1) is the EMPTY key)

    // 1. digits
    G16_LAUNCH(ctx, k_digits, (unsigned)((n + 127) / 128), 128, 0, st, scalars, n, mb->n, c, W, (int)mb->precomp, mb->skip,
               sc->keys_a, sc->vals_a);
    // 2. sort pairs by bucket
    uint32_t *keys = sc->keys_a, *vals = sc->vals_a, *keys2 = sc->keys_b, *vals2 = sc->vals_b;
    G16_TRY(radix_sort(ctx, &keys, &vals, &keys2, &vals2, seg_len, nseg, key_bits, sc->hist, sc->task_tmp, st));
    sc->s_vals = vals;
    // 3. bucket boundaries
    G16_LAUNCH(ctx, k_bucket_bounds, (unsigned)((items + 255) / 256), 256, 0, st, keys, seg_len, nseg, nb, sc->bucket_start);
    // 4. per-level bucket counts / offsets
    G16_TRY(ba_build_levels(ctx, sc, nseg, nb, L, st));
    // 5. tasks over what the affine levels leave: counts -> offsets -> descriptors -> sort by length (descending)
    const uint32_t* tstart = L ? sc->ba_lvl + (size_t)L * sc->ba_stride : sc->ba_start0;
    const uint32_t* tcnt = L ? nullptr : sc->ba_lvl;
    const uint32_t direct_max = L ? kTailDirect : 0;
    const uint32_t task_len = L ? kTaskLenDirect : kTaskLen;
    G16_LAUNCH(ctx, k_task_counts, (unsigned)((nbuckets + 255) / 256), 256, 0, st, tstart, tcnt, nbuckets, direct_max, task_len, sc->task_off);
    G16_TRY(exclusive_scan(ctx, sc->task_off, nbuckets, sc->task_tmp, sc->counters, st));
    uint32_t *tkeys = keys2, *tvals = vals2;  // the alternate sort buffers are free now
    G16_LAUNCH(ctx, k_fill_u32, kNumSMs * 4, 256, 0, st, tkeys, (uint32_t)kTaskLen, tcap);
    uint32_t* t_start = sc->tasks;
    uint32_t* t_len = sc->tasks + sc->cap_tasks;
    G16_LAUNCH(ctx, k_make_tasks, (unsigned)((nbuckets + 255) / 256), 256, 0, st, tstart, tcnt, sc->task_off, nbuckets, direct_max,
               task_len, tkeys, tvals, t_start, t_len);
    uint32_t* tk_alt = sc->tasks + 2 * sc->cap_tasks;  // dedicated: with one 9-bit pass the sorted order ENDS in the alternates,
    uint32_t* tv_alt = sc->tasks + 3 * sc->cap_tasks;  // which k_accumulate reads while it writes sc->partial
    G16_TRY(radix_sort(ctx, &tkeys, &tvals, &tk_alt, &tv_alt, tcap, 1, 9, sc->hist, sc->task_tmp, st));
    sc->s_tkeys = tkeys;
    sc->s_tvals = tvals;
    return G16_OK;
}

// ---- point stage: steps 6-9 --------------------------------------------------------------------------------------------------
template <class F>
static int msm_point_stage(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, const MsmScratch* dg, size_t n, cudaStream_t st,
                           cudaEvent_t ev0, cudaEvent_t ev1) {
    XYZZ<F>* result = (XYZZ<F>*)sc->result;
    const unsigned W = mb->windows, c = mb->c;
    const uint32_t nb = 1u << (c - 1);
    const unsigned nseg = mb->precomp ? 1 : W;
    const size_t items = n * W;
    const size_t nbuckets = (size_t)nseg * nb;
    const int L = mb->ba_levels;
    const size_t tcap = task_capacity(items, nbuckets, L);
    const uint32_t* t_start = dg->tasks;
    const uint32_t* t_len = dg->tasks + dg->cap_tasks;
    // kernel_events == 1: the events bracket the whole bucket accumulation (steps 6-7); == 2: only the first level's k_ba_add
    const bool one_kernel = ctx->opt_kernel_events == 2 && L > 0;
    if (ev0 && !one_kernel) G16_CUDA(ctx, cudaEventRecord(ev0, st));
    // 6. batched-affine levels
    const void* pts = mb->pts;
    if (L) G16_TRY(ba_run_levels(ctx, mb, sc, dg, items, nbuckets, L, &pts, st, one_kernel ? ev0 : nullptr, one_kernel ? ev1 : nullptr));
    // 7. accumulate the survivors
    {
        size_t want = (tcap + 127) / 128;
        unsigned grid = (unsigned)(want < (size_t)kNumSMs * 8 ? want : (size_t)kNumSMs * 8);
        if (L) {
            G16_LAUNCH(ctx, k_bucket_tail<F>, (unsigned)((nbuckets + 127) / 128), 128, 0, st, (const Affine<F>*)pts,
                       dg->ba_lvl + (size_t)L * dg->ba_stride, nbuckets, (XYZZ<F>*)sc->bucket_sum);
            G16_LAUNCH(ctx, (k_accumulate<F, 3, true>), grid, 128, 0, st, (const Affine<F>*)pts, (const uint32_t*)nullptr, dg->s_tkeys,
                       dg->s_tvals, t_start, t_len, (XYZZ<F>*)sc->partial, tcap);
        } else
            G16_LAUNCH(ctx, (k_accumulate<F, 3, false>), grid, 128, 0, st, (const Affine<F>*)pts, dg->s_vals, dg->s_tkeys, dg->s_tvals,
                       t_start, t_len, (XYZZ<F>*)sc->partial, tcap);
    }
    if (ev1 && !one_kernel) G16_CUDA(ctx, cudaEventRecord(ev1, st));
    if (getenv("G16_DEBUG_MSM")) {  // diagnostics only: synchronises
        uint32_t ntasks = 0, totals[kBaMaxLevels + 1] = {};
        cudaStreamSynchronize(st);
        cudaMemcpy(&ntasks, dg->counters, 4, cudaMemcpyDeviceToHost);
        for (int k = 1; k <= L; k++) cudaMemcpy(&totals[k], dg->ba_lvl + (size_t)k * dg->ba_stride + nbuckets, 4, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[g16 msm] group %d n %zu items %zu buckets %zu levels %d tcap %zu tasks %u level totals:", mb->group, n, items,
                nbuckets, L, tcap, ntasks);
        for (int k = 1; k <= L; k++) fprintf(stderr, " %u", totals[k]);
        fprintf(stderr, "\n");
    }
    // 8. combine task partials per bucket
    {
        size_t want = (nbuckets * 32 + 127) / 128;
        unsigned grid = (unsigned)(want < (size_t)kNumSMs * 32 ? want : (size_t)kNumSMs * 32);
        G16_LAUNCH(ctx, k_bucket_combine<F>, grid, 128, 0, st, (const XYZZ<F>*)sc->partial, dg->task_off, dg->counters,
                   nbuckets, (int)(L != 0), (XYZZ<F>*)sc->bucket_sum);
    }
    // 9. reduce every bucket set, then combine the windows
    {
        uint32_t chunks_per_seg = (nb + kChunk - 1) / kChunk;
        size_t chunks = (size_t)nseg * chunks_per_seg;
        G16_LAUNCH(ctx, k_chunk_reduce<F>, (unsigned)((chunks + 63) / 64), 64, 0, st, (const XYZZ<F>*)sc->bucket_sum, nseg, nb,
                   (XYZZ<F>*)sc->chunk_sum);
        unsigned nblk = (unsigned)((chunks_per_seg + 127) / 128);
        if (nblk > 64) nblk = 64;
        XYZZ<F>* blk = (XYZZ<F>*)sc->block_sum;
        XYZZ<F>* setsum = blk + (size_t)nseg * 64;
        size_t smem = 128 * sizeof(XYZZ<F>);
        G16_LAUNCH(ctx, k_tree_sum<F>, nseg * nblk, 128, smem, st, (const XYZZ<F>*)sc->chunk_sum, (size_t)chunks_per_seg, nblk, blk);
        G16_LAUNCH(ctx, k_tree_sum<F>, nseg, 128, smem, st, (const XYZZ<F>*)blk, (size_t)nblk, 1u, setsum);
        G16_LAUNCH(ctx, k_finish<F>, 1, 32, 0, st, (const XYZZ<F>*)setsum, nseg, c, result);
    }
    return G16_OK;
}

int msm_run(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, const Fr* scalars_dev, size_t n, cudaStream_t st, cudaEvent_t ev0,
            cudaEvent_t ev1, const MsmScratch* digits, cudaEvent_t ev_digits_done) {
    if (mb->group != 1 && mb->group != 2) return set_err(ctx, G16_ERR_BAD_ARG, "msm: bases not set");
    if (n > mb->n) return set_err(ctx, G16_ERR_BAD_ARG, "msm: %zu scalars for %zu bases", n, mb->n);
    if (n == 0 || mb->n == 0) {
        if (sc->result) G16_CUDA(ctx, cudaMemsetAsync(sc->result, 0, mb->group == 1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ), st));
        return G16_OK;
    }
    if (!digits) {
        G16_TRY(msm_digit_stage(ctx, mb, sc, scalars_dev, n, st));
        if (ev_digits_done) G16_CUDA(ctx, cudaEventRecord(ev_digits_done, st));
        digits = sc;
    } else if (digits->cap_items != sc->cap_items || digits->cap_buckets != sc->cap_buckets || digits->ba_levels != sc->ba_levels) {
        return set_err(ctx, G16_ERR_BAD_ARG, "msm: shared digit stage has a different geometry");
    }
    if (mb->group == 1) return msm_point_stage<Fq>(ctx, mb, sc, digits, n, st, ev0, ev1);
    return msm_point_stage<Fq2>(ctx, mb, sc, digits, n, st, ev0, ev1);
}

// canonical generators in Montgomery form are produced on the host from their canonical coordinates
static Fq fq_from_words(const uint32_t* w) {
    Fq x;
    for (int i = 0; i < 8; i++) x.v[i] = w[i];
    return x.to_mont();
}
G1Affine g1_generator() {
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, two[8] = {2, 0, 0, 0, 0, 0, 0, 0};
    return G1Affine{fq_from_words(one), fq_from_words(two)};
}
G2Affine g2_generator() {
    // forks/circom-compat/src/zkey.rs:442-462 (decimal) == forks/halo2curves/src/bn256/curve.rs:98-128
    static const uint32_t x0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    static const uint32_t x1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    static const uint32_t y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    static const uint32_t y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    return G2Affine{Fq2{fq_from_words(x0), fq_from_words(x1)}, Fq2{fq_from_words(y0), fq_from_words(y1)}};
}

int fixed_base_dev(g16_ctx* ctx, int group, const Fr* scalars_dev, size_t n, void* out, cudaStream_t st) {
    if (n == 0) return G16_OK;
    unsigned grid = (unsigned)((n + 127) / 128);
    if (group == 1)
        G16_LAUNCH(ctx, k_fixed_base<Fq>, grid, 128, 0, st, scalars_dev, n, g1_generator(), (G1Affine*)out);
    else if (group == 2)
        G16_LAUNCH(ctx, k_fixed_base<Fq2>, grid, 128, 0, st, scalars_dev, n, g2_generator(), (G2Affine*)out);
    else
        return set_err(ctx, G16_ERR_BAD_ARG, "group must be 1 or 2");
    return G16_OK;
}

}  // namespace g16
