// msm.cu -- Pippenger multi-scalar multiplication for BN254 G1 (Fq) and G2 (Fq2)  (K5-K8).
//
// Replaces ark-ec 0.4 VariableBaseMSM::msm_bigint at the reference call sites forks/groth16/src/prover.rs:66,74,266.
// The result is the exact group element Sigma s_i * P_i, so it equals the reference's after normalisation whatever
// the evaluation order.
//
// Pipeline (all on the device, stream-ordered, no host synchronisation):
//   1. k_digits        scalar (Montgomery) -> canonical -> W signed c-bit digits; emits (bucket key, point ref) pairs.
//                      Zero digits and points at infinity get the EMPTY key.
//   2. radix sort      hand-written LSD radix sort, 8 bits per pass, stable block-local ranking with warp match;
//                      segments (= windows) never mix because histograms are laid out [segment][bin][tile].
//   3. k_bucket_bounds bucket start offsets from the sorted keys (no atomics).
//   4. tasks           every bucket is cut into tasks of <= kTaskLen points; tasks are radix-sorted by length
//                      (descending) so that the lanes of a warp run equally long loops whatever the scalar skew.
//   5. k_accumulate    one thread per task: XYZZ mixed additions (8M + 2S each), next point prefetched.
//   6. k_bucket_combine one warp per bucket folds that bucket's task partials (shuffle tree when there are several).
//   7. k_chunk_reduce / k_tree_sum / k_finish
//                      Sigma (b+1) * B_b per bucket set by chunked running sums + short double-and-add, tree sums,
//                      and (without precomputation) the Horner combination of the windows.
//
// With `precomp` the bases hold 2^(c*w) * P_i for every window w, all windows share ONE bucket set and steps 7's
// Horner tail disappears: HBM capacity (180 GB) is traded for integer-pipe work.
#include "internal.cuh"

namespace g16 {

constexpr unsigned kTaskLen = 256;  // max points per accumulate task
constexpr unsigned kChunk = 16;     // buckets per reduction chunk
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096
constexpr uint32_t kNegBit = 0x80000000u;

template <class F>
struct PointBytes {
    static constexpr size_t affine = sizeof(Affine<F>);
    static constexpr size_t xyzz = sizeof(XYZZ<F>);
};

// ---- vectorised loads/stores ------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
    return r;
}
template <class T>
__device__ __forceinline__ T ldg_vec(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = __ldg(s + i);
    return r;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
}

// ---- generic exclusive scan (uint32), in place -------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sh /*>= 33*/) {
    // exclusive scan of one value per thread across a 256/1024-thread block; returns the prefix, *total = block sum
    unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned nw = (blockDim.x + 31) >> 5;
        uint32_t s = lane < nw ? sh[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (unsigned)o) s += y;
        }
        if (lane < nw) sh[lane] = s;  // inclusive warp totals
        if (lane == nw - 1) sh[32] = s;
    }
    __syncthreads();
    uint32_t base = wid ? sh[wid - 1] : 0;
    *total = sh[32];
    uint32_t r = base + x - v;
    __syncthreads();
    return r;
}

__global__ void k_scan_tile_sums(const uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[33];
    size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        size_t i = base + (size_t)threadIdx.x * kScanItems + k;
        if (i < n) s += data[i];
    }
    uint32_t tot;
    block_exclusive_scan(s, &tot, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
__global__ void k_scan_single(uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sh[33];
    uint32_t carry = 0;
    for (size_t base = 0; base < n; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0;
        uint32_t tot;
        uint32_t p = block_exclusive_scan(v, &tot, sh);
        if (i < n) data[i] = carry + p;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void k_scan_apply(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[33];
    size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        size_t i = base + (size_t)threadIdx.x * kScanItems + k;
        v[k] = i < n ? data[i] : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t p = block_exclusive_scan(s, &tot, sh) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        size_t i = base + (size_t)threadIdx.x * kScanItems + k;
        if (i < n) data[i] = p;
        p += v[k];
    }
}
// data[0..n) -> exclusive prefix sums in place; tmp needs ceil(n / kScanTile) + 1 words; total (optional) device ptr
static int exclusive_scan(g16_ctx* ctx, uint32_t* data, size_t n, uint32_t* tmp, uint32_t* total_dev, cudaStream_t st) {
    if (n == 0) {
        if (total_dev) G16_CUDA(ctx, cudaMemsetAsync(total_dev, 0, 4, st));
        return G16_OK;
    }
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        G16_LAUNCH(ctx, k_scan_single, 1, 1024, 0, st, data, n, total_dev);
        return G16_OK;
    }
    G16_LAUNCH(ctx, k_scan_tile_sums, (unsigned)tiles, kScanThreads, 0, st, data, n, tmp);
    G16_LAUNCH(ctx, k_scan_single, 1, 1024, 0, st, tmp, tiles, total_dev);
    G16_LAUNCH(ctx, k_scan_apply, (unsigned)tiles, kScanThreads, 0, st, data, n, tmp);
    return G16_OK;
}

// ---- LSD radix sort of (key, value) pairs, 8..11 bits per pass, segmented ------------------------------------------------
// hist layout: [segment][bin][tile]  (flat exclusive scan => absolute output positions, segments stay separate
// because every segment holds exactly seg_len items).  Wide digits keep the pass count at 2 for 20-bit bucket keys.
constexpr unsigned kRsMaxBits = 11;

template <unsigned BITS>
__global__ void __launch_bounds__(kRsThreads)
    k_rs_hist(const uint32_t* __restrict__ keys, size_t seg_len, unsigned tiles_per_seg, unsigned shift,
              uint32_t* __restrict__ hist) {
    constexpr unsigned NB = 1u << BITS;
    __shared__ uint32_t sh[NB];
    unsigned seg = blockIdx.x / tiles_per_seg, tile = blockIdx.x % tiles_per_seg;
    for (unsigned b = threadIdx.x; b < NB; b += kRsThreads) sh[b] = 0;
    __syncthreads();
    size_t seg_base = (size_t)seg * seg_len;
    size_t lo = (size_t)tile * kRsTile;
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        size_t i = lo + (size_t)k * kRsThreads + threadIdx.x;
        if (i < seg_len) atomicAdd(&sh[(keys[seg_base + i] >> shift) & (NB - 1)], 1u);
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < NB; b += kRsThreads) hist[((size_t)seg * NB + b) * tiles_per_seg + tile] = sh[b];
}

template <unsigned BITS>
__global__ void __launch_bounds__(kRsThreads)
    k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ okeys,
                 uint32_t* __restrict__ ovals, size_t seg_len, unsigned tiles_per_seg, unsigned shift,
                 const uint32_t* __restrict__ hist) {
    constexpr unsigned NB = 1u << BITS;
    constexpr unsigned NW = kRsThreads / 32;
    extern __shared__ uint32_t wcnt[];  // [NW][NB]
    unsigned seg = blockIdx.x / tiles_per_seg, tile = blockIdx.x % tiles_per_seg;
    unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned i = threadIdx.x; i < NW * NB; i += kRsThreads) wcnt[i] = 0;
    __syncthreads();
    uint32_t* mine = wcnt + wid * NB;
    size_t seg_base = (size_t)seg * seg_len;
    // warp w owns items [w*512, (w+1)*512) of the tile, as 16 rows of 32: ranking order == index order (stable)
    size_t lo = (size_t)tile * kRsTile + (size_t)wid * (32 * kRsItems);
    uint32_t k[kRsItems], v[kRsItems];
    uint32_t packed[kRsItems];  // rank within the row's digit group | group size << 8 | leader << 16 | valid << 17
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * 32 + lane;
        bool valid = i < seg_len;
        k[r] = valid ? keys[seg_base + i] : 0xffffffffu;
        v[r] = valid ? vals[seg_base + i] : 0u;
        uint32_t d = valid ? ((k[r] >> shift) & (NB - 1)) : (NB + lane);
        uint32_t mask = __match_any_sync(0xffffffffu, d);
        uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        uint32_t cnt = __popc(mask);
        bool leader = rank == 0;
        packed[r] = rank | (cnt << 8) | ((leader && valid) ? 0x10000u : 0u) | (valid ? 0x20000u : 0u);
        if (leader && valid) mine[d] += cnt;
        __syncwarp();
    }
    __syncthreads();
    for (unsigned bin = threadIdx.x; bin < NB; bin += kRsThreads) {
        uint32_t run = hist[((size_t)seg * NB + bin) * tiles_per_seg + tile];
#pragma unroll
        for (unsigned w = 0; w < NW; w++) {
            uint32_t t = wcnt[w * NB + bin];
            wcnt[w * NB + bin] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        bool valid = (packed[r] & 0x20000u) != 0;
        uint32_t d = (k[r] >> shift) & (NB - 1);
        uint32_t pos = 0;
        if (valid) pos = mine[d] + (packed[r] & 0xffu);
        __syncwarp();
        if (packed[r] & 0x10000u) mine[d] += (packed[r] >> 8) & 0xffu;
        __syncwarp();
        if (valid) {
            okeys[pos] = k[r];
            ovals[pos] = v[r];
        }
    }
}

template <unsigned BITS>
static int radix_pass(g16_ctx* ctx, const uint32_t* keys, const uint32_t* vals, uint32_t* okeys, uint32_t* ovals, size_t seg_len,
                      unsigned nseg, unsigned tiles, unsigned shift, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t st) {
    constexpr unsigned NB = 1u << BITS;
    const size_t smem = (size_t)(kRsThreads / 32) * NB * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        G16_CUDA(ctx, cudaFuncSetAttribute(k_rs_scatter<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    G16_LAUNCH(ctx, k_rs_hist<BITS>, nseg * tiles, kRsThreads, 0, st, keys, seg_len, tiles, shift, hist);
    G16_TRY(exclusive_scan(ctx, hist, (size_t)nseg * NB * tiles, scan_tmp, nullptr, st));
    G16_LAUNCH(ctx, k_rs_scatter<BITS>, nseg * tiles, kRsThreads, smem, st, keys, vals, okeys, ovals, seg_len, tiles, shift, hist);
    return G16_OK;
}

static unsigned radix_digit_bits(unsigned bits) {
    unsigned passes = (bits + kRsMaxBits - 1) / kRsMaxBits;
    unsigned per = (bits + passes - 1) / passes;
    return per < 8 ? 8 : per;
}

// sorts nseg segments of seg_len pairs by the low `bits` key bits; result ends in (*keys, *vals) (buffers may swap)
static int radix_sort(g16_ctx* ctx, uint32_t** keys, uint32_t** vals, uint32_t** keys_alt, uint32_t** vals_alt,
                      size_t seg_len, unsigned nseg, unsigned bits, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t st) {
    if (seg_len == 0 || nseg == 0) return G16_OK;
    unsigned tiles = (unsigned)((seg_len + kRsTile - 1) / kRsTile);
    unsigned per = radix_digit_bits(bits);
    for (unsigned shift = 0; shift < bits; shift += per) {
        int rc;
        switch (per) {
            case 8: rc = radix_pass<8>(ctx, *keys, *vals, *keys_alt, *vals_alt, seg_len, nseg, tiles, shift, hist, scan_tmp, st); break;
            case 9: rc = radix_pass<9>(ctx, *keys, *vals, *keys_alt, *vals_alt, seg_len, nseg, tiles, shift, hist, scan_tmp, st); break;
            case 10: rc = radix_pass<10>(ctx, *keys, *vals, *keys_alt, *vals_alt, seg_len, nseg, tiles, shift, hist, scan_tmp, st); break;
            default: rc = radix_pass<11>(ctx, *keys, *vals, *keys_alt, *vals_alt, seg_len, nseg, tiles, shift, hist, scan_tmp, st); break;
        }
        G16_TRY(rc);
        std::swap(*keys, *keys_alt);
        std::swap(*vals, *vals_alt);
    }
    return G16_OK;
}

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

// ---- 4. tasks ----------------------------------------------------------------------------------------------------------------
__global__ void k_task_counts(const uint32_t* __restrict__ start, unsigned nseg, uint32_t nb, uint32_t* __restrict__ ntasks) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)nseg * nb;
    if (g >= total) return;
    unsigned seg = (unsigned)(g / nb);
    uint32_t b = (uint32_t)(g - (size_t)seg * nb);
    const uint32_t* st = start + (size_t)seg * (nb + 1);
    uint32_t cnt = st[b + 1] - st[b];
    ntasks[g] = (cnt + kTaskLen - 1) / kTaskLen;
}
__global__ void k_fill_u32(uint32_t* __restrict__ p, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
// task t of bucket g covers sorted positions [start + j*L, ...); sort key = kTaskLen - len (longest first)
__global__ void k_make_tasks(const uint32_t* __restrict__ start, const uint32_t* __restrict__ task_off, unsigned nseg,
                             uint32_t nb, uint32_t* __restrict__ tkeys, uint32_t* __restrict__ tvals,
                             uint32_t* __restrict__ t_start, uint32_t* __restrict__ t_len) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)nseg * nb;
    if (g >= total) return;
    unsigned seg = (unsigned)(g / nb);
    uint32_t b = (uint32_t)(g - (size_t)seg * nb);
    const uint32_t* st = start + (size_t)seg * (nb + 1);
    uint32_t s0 = st[b], cnt = st[b + 1] - s0;
    uint32_t t = task_off[g];
    for (uint32_t o = 0; o < cnt; o += kTaskLen, t++) {
        uint32_t len = cnt - o < kTaskLen ? cnt - o : kTaskLen;
        t_start[t] = s0 + o;
        t_len[t] = len;
        tkeys[t] = kTaskLen - len;
        tvals[t] = t;
    }
}

// ---- 5. bucket accumulation --------------------------------------------------------------------------------------------------
template <class F, int MINB>
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
        uint32_t ref = vals[s0];
        Affine<F> nxt = ldg_vec(bases + (ref & ~kNegBit));
        for (uint32_t k = 0; k < len; k++) {
            Affine<F> cur = nxt;
            uint32_t cref = ref;
            if (k + 1 < len) {
                ref = vals[s0 + k + 1];
                nxt = ldg_vec(bases + (ref & ~kNegBit));
            }
            if (cref & kNegBit) cur.y = cur.y.neg();
            acc.madd(cur);
        }
        st_vec(partial + t, acc);
    }
}

// ---- 6. per-bucket combine (one warp per bucket) ----------------------------------------------------------------------------
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
                     const uint32_t* __restrict__ total_tasks, size_t nbuckets, XYZZ<F>* __restrict__ bucket_sum) {
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned lane = threadIdx.x & 31;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b < nbuckets; b += nwarps) {
        uint32_t t0 = task_off[b];
        uint32_t t1 = (b + 1 < nbuckets) ? task_off[b + 1] : *total_tasks;
        uint32_t cnt = t1 - t0;
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

// ---- 7. bucket-set reduction:  S = Sigma_{b} (b + 1) * B_b ------------------------------------------------------------------
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
    for (uint32_t j = hi; j-- > lo;) {
        run.add(ld_vec(B + j));
        acc.add(run);
    }
    if (lo && !run.is_inf()) {
        XYZZ<F> t = XYZZ<F>::inf();
        for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
            t = t.dbl();
            if ((lo >> bit) & 1u) t.add(run);
        }
        acc.add(t);
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
    } else {
        c = (int)lg - 1;
        if (c > 20) c = 20;
    }
    if (c < 4) c = 4;
    return c;
}

static size_t task_capacity(size_t items, size_t nbuckets) { return items / kTaskLen + nbuckets + 1; }

template <class F>
static int msm_alloc_scratch(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc) {
    const size_t n = mb->n;
    const unsigned W = mb->windows;
    const uint32_t nb = 1u << (mb->c - 1);
    const unsigned nseg = mb->precomp ? 1 : W;
    size_t items = n * W;
    size_t nbuckets = (size_t)nseg * nb;
    size_t tcap = task_capacity(items, nbuckets);
    size_t sort_n = items > tcap ? items : tcap;
    size_t seg_len = mb->precomp ? items : n;
    size_t tiles = (seg_len + kRsTile - 1) / kRsTile;
    size_t hist_n = (size_t)nseg * (1u << kRsMaxBits) * tiles;
    size_t ttiles = (tcap + kRsTile - 1) / kRsTile;
    if ((size_t)(1u << kRsMaxBits) * ttiles > hist_n) hist_n = (size_t)(1u << kRsMaxBits) * ttiles;
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
    size_t scan_tmp = (hist_n > nbuckets ? hist_n : nbuckets) / kScanTile + 2;
    G16_TRY(dev_alloc(ctx, &sc->task_tmp, scan_tmp));
    G16_TRY(dev_alloc(ctx, &sc->tasks, 4 * tcap));  // t_start | t_len | task-sort alternates (keys, order)
    G16_TRY(dev_alloc(ctx, &sc->counters, 16));
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

template <class F>
static int msm_run_t(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, const Fr* scalars, size_t n, cudaStream_t st, cudaEvent_t ev0,
                     cudaEvent_t ev1) {
    XYZZ<F>* result = (XYZZ<F>*)sc->result;
    if (n > mb->n) return set_err(ctx, G16_ERR_BAD_ARG, "msm: %zu scalars for %zu bases", n, mb->n);
    if (n == 0 || mb->n == 0) {
        if (result) G16_CUDA(ctx, cudaMemsetAsync(result, 0, sizeof(XYZZ<F>), st));
        return G16_OK;
    }
    const unsigned W = mb->windows, c = mb->c;
    const uint32_t nb = 1u << (c - 1);
    const unsigned nseg = mb->precomp ? 1 : W;
    const size_t items = n * W;
    const size_t seg_len = mb->precomp ? items : n;
    const size_t nbuckets = (size_t)nseg * nb;
    const size_t tcap = task_capacity(items, nbuckets);
    unsigned key_bits = c;  // keys in [0, nb] need c bits (nb = 2^(c-1) is the EMPTY key)

    // 1. digits
    G16_LAUNCH(ctx, k_digits, (unsigned)((n + 127) / 128), 128, 0, st, scalars, n, mb->n, c, W, (int)mb->precomp, mb->skip,
               sc->keys_a, sc->vals_a);
    // 2. sort pairs by bucket
    uint32_t *keys = sc->keys_a, *vals = sc->vals_a, *keys2 = sc->keys_b, *vals2 = sc->vals_b;
    G16_TRY(radix_sort(ctx, &keys, &vals, &keys2, &vals2, seg_len, nseg, key_bits, sc->hist, sc->task_tmp, st));
    // 3. bucket boundaries
    G16_LAUNCH(ctx, k_bucket_bounds, (unsigned)((items + 255) / 256), 256, 0, st, keys, seg_len, nseg, nb, sc->bucket_start);
    // 4. tasks: counts -> offsets -> descriptors -> sort by length (descending)
    G16_LAUNCH(ctx, k_task_counts, (unsigned)((nbuckets + 255) / 256), 256, 0, st, sc->bucket_start, nseg, nb, sc->task_off);
    G16_TRY(exclusive_scan(ctx, sc->task_off, nbuckets, sc->task_tmp, sc->counters, st));
    uint32_t *tkeys = keys2, *tvals = vals2;  // the alternate sort buffers are free now
    uint32_t *tkeys2 = sc->keys_a == keys ? nullptr : nullptr;
    (void)tkeys2;
    G16_LAUNCH(ctx, k_fill_u32, kNumSMs * 4, 256, 0, st, tkeys, (uint32_t)kTaskLen, tcap);
    uint32_t* t_start = sc->tasks;
    uint32_t* t_len = sc->tasks + tcap;
    G16_LAUNCH(ctx, k_make_tasks, (unsigned)((nbuckets + 255) / 256), 256, 0, st, sc->bucket_start, sc->task_off, nseg, nb,
               tkeys, tvals, t_start, t_len);
    uint32_t* tk_alt = sc->tasks + 2 * tcap;  // dedicated: with one 9-bit pass the sorted order ENDS in the alternates,
    uint32_t* tv_alt = sc->tasks + 3 * tcap;  // which k_accumulate reads while it writes sc->partial
    {
        uint32_t *a = tkeys, *b = tvals, *a2 = tk_alt, *b2 = tv_alt;
        G16_TRY(radix_sort(ctx, &a, &b, &a2, &b2, tcap, 1, 9, sc->hist, sc->task_tmp, st));
        tkeys = a;
        tvals = b;
    }
    // 5. accumulate
    {
        size_t want = (tcap + 127) / 128;
        unsigned grid = (unsigned)(want < (size_t)kNumSMs * 8 ? want : (size_t)kNumSMs * 8);
        if (ev0) G16_CUDA(ctx, cudaEventRecord(ev0, st));
        const int variant = ctx->opt_acc_variant;
#define G16_ACC(MINB)                                                                                                        \
    G16_LAUNCH(ctx, (k_accumulate<F, MINB>), grid, 128, 0, st, (const Affine<F>*)mb->pts, vals, tkeys, tvals, t_start, t_len, \
               (XYZZ<F>*)sc->partial, tcap)
        if constexpr (sizeof(F) == sizeof(Fq)) {
            if (variant == 4) G16_ACC(4);
            else G16_ACC(3);  // 148 registers, no spills; occupancy beyond 3 blocks/SM buys nothing (measured)
        } else {
            if (variant == 2) G16_ACC(2);
            else G16_ACC(3);  // 168 registers with out-of-line Fq2 products
        }
#undef G16_ACC
        if (ev1) G16_CUDA(ctx, cudaEventRecord(ev1, st));
    }
    // 6. combine task partials per bucket
    {
        size_t want = (nbuckets * 32 + 127) / 128;
        unsigned grid = (unsigned)(want < (size_t)kNumSMs * 32 ? want : (size_t)kNumSMs * 32);
        G16_LAUNCH(ctx, k_bucket_combine<F>, grid, 128, 0, st, (const XYZZ<F>*)sc->partial, sc->task_off, sc->counters,
                   nbuckets, (XYZZ<F>*)sc->bucket_sum);
    }
    // 7. reduce every bucket set, then combine the windows
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
            cudaEvent_t ev1) {
    if (mb->group == 1) return msm_run_t<Fq>(ctx, mb, sc, scalars_dev, n, st, ev0, ev1);
    if (mb->group == 2) return msm_run_t<Fq2>(ctx, mb, sc, scalars_dev, n, st, ev0, ev1);
    return set_err(ctx, G16_ERR_BAD_ARG, "msm: bases not set");
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
