// msm_internal.cuh -- declarations shared by the MSM translation units (msm.cu, msm_sort.cu, msm_affine.cu).
#pragma once
#include "internal.cuh"

namespace g16 {

constexpr unsigned kTaskLen = 256;  // max points per XYZZ accumulate task
constexpr unsigned kChunk = 8;      // buckets per reduction chunk (16 adds + a ~17-bit double-and-add per thread: the chain is pure latency)
constexpr uint32_t kNegBit = 0x80000000u;

// batched-affine bucket accumulation (msm_affine.cu)
constexpr int kBaKMax = 16;       // output slots per thread at full size (one shared-inversion chain per thread); see ba_slots_per_thread
constexpr int kBaGroup = 8;       // thread products per Fq inversion in k_ba_invert
constexpr int kBaMaxLevels = 8;   // pairwise levels run in affine form before the XYZZ tail
constexpr int kBaDefaultLevels = 5;
constexpr uint32_t kTailDirect = 32;  // after the affine levels: buckets up to this size are summed by one thread, no tasks

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

// ---- msm_sort.cu ----------------------------------------------------------------------------------------------------------
// data[0..n) -> exclusive prefix sums in place; tmp needs scan_tmp_words(n) words; total_dev (optional) receives the sum
size_t scan_tmp_words(size_t n);
int exclusive_scan(g16_ctx* ctx, uint32_t* data, size_t n, uint32_t* tmp, uint32_t* total_dev, cudaStream_t st);
// `batch` independent scans of n words each; row b starts at data + b*stride; tmp needs batch * scan_tmp_words(n) words
int exclusive_scan_batched(g16_ctx* ctx, uint32_t* data, size_t n, size_t stride, unsigned batch, uint32_t* tmp, cudaStream_t st);
// histogram words the sort needs for nseg segments of seg_len pairs
size_t radix_hist_words(size_t seg_len, unsigned nseg);
// sorts nseg segments of seg_len (key, value) pairs by the low `bits` key bits; the result ends in (*keys, *vals)
// (the buffer pairs may swap).  Equal keys keep no particular order.
int radix_sort(g16_ctx* ctx, uint32_t** keys, uint32_t** vals, uint32_t** keys_alt, uint32_t** vals_alt, size_t seg_len,
               unsigned nseg, unsigned bits, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t st);

// ---- msm_affine.cu ------------------------------------------------------------------------------------------------------
// upper bound of the number of points left after `level` pairwise levels
size_t ba_level_cap(size_t items, size_t nbuckets, int level);
int ba_slots_per_thread(size_t items, size_t nbuckets, int level);
int ba_alloc(g16_ctx* ctx, int group, MsmScratch* sc, size_t items, size_t nbuckets, int levels);
void ba_free(MsmScratch* sc);
// level tables of a digit set: per-bucket point counts and offsets after every pairwise level (point independent)
int ba_build_levels(g16_ctx* ctx, MsmScratch* dg, unsigned nseg, uint32_t nb, int levels, cudaStream_t st);
// runs `levels` batched-affine levels over the sorted references of `dg`; leaves the surviving points (affine, sign applied)
// in *out_pts, laid out by dg's level-`levels` offsets.  add0_ev0/1 (optional) bracket the first level's k_ba_add launch.
int ba_run_levels(g16_ctx* ctx, const MsmBases* mb, MsmScratch* sc, const MsmScratch* dg, size_t items, size_t nbuckets,
                  int levels, const void** out_pts, cudaStream_t st, cudaEvent_t add0_ev0 = nullptr, cudaEvent_t add0_ev1 = nullptr);

}  // namespace g16
