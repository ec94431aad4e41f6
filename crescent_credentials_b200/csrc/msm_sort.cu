// msm_sort.cu -- device-wide exclusive scan and the hand-written LSD radix sort that groups (bucket key, point reference)
// pairs by bucket for the Pippenger MSMs (K6).  arkworks has no counterpart: its msm_bigint walks the scalars once per
// window on one core (ark-ec 0.4 VariableBaseMSM, called at forks/groth16/src/prover.rs:66,74,266).
//
// Sort: 8..11 key bits per pass (two passes for the 20-bit keys of the c = 20 bucket set), tiles of 4096 pairs,
// per-tile histograms laid out [segment][bin][tile] so that one flat exclusive scan yields absolute output positions and
// segments (= windows, when every window has its own bucket set) never mix.  Ranking inside a tile is stable
// (warp match + per-warp counters), which is what makes the second LSD pass correct.
#include "msm_internal.cuh"

namespace g16 {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096
constexpr unsigned kRsMaxBits = 11;

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

// blockIdx.y selects the row of a batched scan
__global__ void k_scan_tile_sums(const uint32_t* __restrict__ data, size_t n, size_t stride, uint32_t* __restrict__ sums,
                                 size_t sums_stride) {
    __shared__ uint32_t sh[33];
    data += (size_t)blockIdx.y * stride;
    sums += (size_t)blockIdx.y * sums_stride;
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
__global__ void k_scan_single(uint32_t* __restrict__ data, size_t n, size_t stride, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sh[33];
    data += (size_t)blockIdx.y * stride;
    uint32_t carry = 0;
    for (size_t base = 0; base < n; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0;
        uint32_t tot;
        uint32_t p = block_exclusive_scan(v, &tot, sh);
        if (i < n) data[i] = carry + p;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) total_out[blockIdx.y] = carry;
}
__global__ void k_scan_apply(uint32_t* __restrict__ data, size_t n, size_t stride, const uint32_t* __restrict__ sums,
                             size_t sums_stride) {
    __shared__ uint32_t sh[33];
    data += (size_t)blockIdx.y * stride;
    sums += (size_t)blockIdx.y * sums_stride;
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

size_t scan_tmp_words(size_t n) { return (n + kScanTile - 1) / kScanTile + 2; }

int exclusive_scan_batched(g16_ctx* ctx, uint32_t* data, size_t n, size_t stride, unsigned batch, uint32_t* tmp, cudaStream_t st) {
    if (n == 0 || batch == 0) return G16_OK;
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        G16_LAUNCH(ctx, k_scan_single, dim3(1, batch), 1024, 0, st, data, n, stride, (uint32_t*)nullptr);
        return G16_OK;
    }
    size_t ts = scan_tmp_words(n);
    G16_LAUNCH(ctx, k_scan_tile_sums, dim3((unsigned)tiles, batch), kScanThreads, 0, st, data, n, stride, tmp, ts);
    G16_LAUNCH(ctx, k_scan_single, dim3(1, batch), 1024, 0, st, tmp, tiles, ts, (uint32_t*)nullptr);
    G16_LAUNCH(ctx, k_scan_apply, dim3((unsigned)tiles, batch), kScanThreads, 0, st, data, n, stride, tmp, ts);
    return G16_OK;
}

int exclusive_scan(g16_ctx* ctx, uint32_t* data, size_t n, uint32_t* tmp, uint32_t* total_dev, cudaStream_t st) {
    if (n == 0) {
        if (total_dev) G16_CUDA(ctx, cudaMemsetAsync(total_dev, 0, 4, st));
        return G16_OK;
    }
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        G16_LAUNCH(ctx, k_scan_single, 1, 1024, 0, st, data, n, (size_t)0, total_dev);
        return G16_OK;
    }
    G16_LAUNCH(ctx, k_scan_tile_sums, (unsigned)tiles, kScanThreads, 0, st, data, n, (size_t)0, tmp, (size_t)0);
    G16_LAUNCH(ctx, k_scan_single, 1, 1024, 0, st, tmp, tiles, (size_t)0, total_dev);
    G16_LAUNCH(ctx, k_scan_apply, (unsigned)tiles, kScanThreads, 0, st, data, n, (size_t)0, tmp, (size_t)0);
    return G16_OK;
}

// ---- LSD radix sort of (key, value) pairs, 8..11 bits per pass, segmented ------------------------------------------------
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
    uint32_t k[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * kRsThreads + threadIdx.x;
        k[r] = i < seg_len ? keys[seg_base + i] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * kRsThreads + threadIdx.x;
        if (i < seg_len) atomicAdd(&sh[(k[r] >> shift) & (NB - 1)], 1u);
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < NB; b += kRsThreads) hist[((size_t)seg * NB + b) * tiles_per_seg + tile] = sh[b];
}

// Scatter of one tile.  Warp w owns items [w*512, (w+1)*512) of the tile as 16 rows of 32, so ranking order == index
// order (stable).  The tile is first sorted by digit INTO SHARED MEMORY and then written out in that order: the items of
// a bin leave as one contiguous run, so a warp's store touches a handful of 32-byte sectors instead of 32 (the direct
// scatter was bound by L2 sector writes carrying 4 useful bytes each).
template <unsigned BITS>
__global__ void __launch_bounds__(kRsThreads, 3)
    k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ okeys,
                 uint32_t* __restrict__ ovals, size_t seg_len, unsigned tiles_per_seg, unsigned shift,
                 const uint32_t* __restrict__ hist) {
    constexpr unsigned NB = 1u << BITS;
    constexpr unsigned NW = kRsThreads / 32;
    constexpr unsigned BPT = NB / kRsThreads;  // bins per thread in the prefix step (NB >= 256)
    extern __shared__ uint32_t smem[];
    uint32_t* wcnt = smem;                 // [NW][NB] per-warp bin counts, then per-warp local offsets
    uint32_t* local_off = wcnt + NW * NB;  // [NB] first tile-local slot of every bin
    uint32_t* gbase = local_off + NB;      // [NB] first global slot of this tile's items of every bin
    uint32_t* skey = gbase + NB;           // [kRsTile] the tile sorted by digit
    uint32_t* sval = skey + kRsTile;
    __shared__ uint32_t scan_sh[33];
    unsigned seg = blockIdx.x / tiles_per_seg, tile = blockIdx.x % tiles_per_seg;
    unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned i = threadIdx.x; i < NW * NB; i += kRsThreads) wcnt[i] = 0;
    __syncthreads();
    uint32_t* mine = wcnt + wid * NB;
    size_t seg_base = (size_t)seg * seg_len;
    size_t tile_lo = (size_t)tile * kRsTile;
    size_t lo = tile_lo + (size_t)wid * (32 * kRsItems);
    uint32_t k[kRsItems];
    uint32_t packed[kRsItems];  // rank within the row's digit group | group size << 8 | leader << 16 | valid << 17
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * 32 + lane;
        k[r] = i < seg_len ? keys[seg_base + i] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * 32 + lane;
        bool valid = i < seg_len;
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
    // bins [BPT*t, BPT*(t+1)) belong to thread t: bin totals -> exclusive scan over the bins -> per-warp offsets
    {
        uint32_t tot[BPT];
        uint32_t tsum = 0;
#pragma unroll
        for (unsigned j = 0; j < BPT; j++) {
            unsigned bin = threadIdx.x * BPT + j;
            uint32_t t = 0;
#pragma unroll
            for (unsigned w = 0; w < NW; w++) t += wcnt[w * NB + bin];
            tot[j] = t;
            tsum += t;
        }
        uint32_t dummy;
        uint32_t run = block_exclusive_scan(tsum, &dummy, scan_sh);
#pragma unroll
        for (unsigned j = 0; j < BPT; j++) {
            unsigned bin = threadIdx.x * BPT + j;
            local_off[bin] = run;
            gbase[bin] = hist[((size_t)seg * NB + bin) * tiles_per_seg + tile];
            uint32_t o = run;
#pragma unroll
            for (unsigned w = 0; w < NW; w++) {
                uint32_t t = wcnt[w * NB + bin];
                wcnt[w * NB + bin] = o;
                o += t;
            }
            run += tot[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        size_t i = lo + (size_t)r * 32 + lane;
        bool valid = (packed[r] & 0x20000u) != 0;
        uint32_t v = valid ? vals[seg_base + i] : 0u;
        uint32_t d = (k[r] >> shift) & (NB - 1);
        uint32_t pos = 0;
        if (valid) pos = mine[d] + (packed[r] & 0xffu);
        __syncwarp();
        if (packed[r] & 0x10000u) mine[d] += (packed[r] >> 8) & 0xffu;
        __syncwarp();
        if (valid) {
            skey[pos] = k[r];
            sval[pos] = v;
        }
    }
    __syncthreads();
    size_t left = seg_len - tile_lo;
    unsigned count = left < (size_t)kRsTile ? (unsigned)left : (unsigned)kRsTile;
    for (unsigned i = threadIdx.x; i < count; i += kRsThreads) {
        uint32_t key = skey[i];
        uint32_t d = (key >> shift) & (NB - 1);
        uint32_t dst = gbase[d] + (i - local_off[d]);
        okeys[dst] = key;
        ovals[dst] = sval[i];
    }
}

template <unsigned BITS>
static int radix_pass(g16_ctx* ctx, const uint32_t* keys, const uint32_t* vals, uint32_t* okeys, uint32_t* ovals, size_t seg_len,
                      unsigned nseg, unsigned tiles, unsigned shift, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t st) {
    constexpr unsigned NB = 1u << BITS;
    const size_t smem = ((size_t)(kRsThreads / 32) * NB + 2 * NB + 2 * kRsTile) * sizeof(uint32_t);
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

size_t radix_hist_words(size_t seg_len, unsigned nseg) {
    size_t tiles = (seg_len + kRsTile - 1) / kRsTile;
    return (size_t)nseg * (1u << kRsMaxBits) * (tiles ? tiles : 1);
}

int radix_sort(g16_ctx* ctx, uint32_t** keys, uint32_t** vals, uint32_t** keys_alt, uint32_t** vals_alt, size_t seg_len,
               unsigned nseg, unsigned bits, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t st) {
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

}  // namespace g16
