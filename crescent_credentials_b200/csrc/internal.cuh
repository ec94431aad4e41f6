// internal.cuh -- context object and cross-translation-unit declarations of libg16b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/g16_b200.h"
#include "ec.cuh"

namespace g16 {

constexpr int kSideStreams = 9;  // 0-3 MSM chains, 4 (r, s)-only scalar multiplications, 5-6 s*MSM_a / r*MSM_b1, 7-8 second MSM of a split chain
constexpr int kMsmSlots = 8;
constexpr int kNumSMs = 148;  // B200

// ---- NTT domain tables (one per log_n, built lazily on the device) ----------------------------------------------
struct NttTables {
    unsigned log_n = 0;
    Fr* tw = nullptr;         // omega^k, k < n/2
    Fr* tw_inv = nullptr;     // omega^-k, k < n/2
    Fr* coset = nullptr;      // g^i, i < n                    (forward coset pre-scale)
    Fr* coset_inv = nullptr;  // g^-i / n, i < n               (inverse coset post-scale)
    Fr n_inv;                 // 1/n (host copy, Montgomery)
    Fr zinv;                  // 1/(g^n - 1)
    bool zinv_ok = false;
    Fr* coset_scaled = nullptr;  // g^i / n                    (iNTT's 1/n folded into the following coset NTT)
    Fr* odd_scaled = nullptr;    // omega_2n^i / n, i < n      (CircomReduction coset: qap.rs:63-72, same folding)
    Fr* coset_inv_z = nullptr;   // g^-i / (n (g^n - 1))        (coset iNTT post-scale with the division by Z folded in)
    Fr zinv_n;                   // 1 / (n (g^n - 1))
};

// ---- MSM engine state for one set of bases -------------------------------------------------------------------------
struct MsmBases {
    int group = 0;          // 1 = G1 (64 B points), 2 = G2 (128 B points)
    size_t n = 0;           // number of (logical) points
    int c = 0;              // window bits
    int windows = 0;        // number of signed windows
    bool precomp = false;   // bases hold windows * n points: 2^(c*w) * P_i at [w*n + i]
    void* pts = nullptr;    // device: Affine<Fq> or Affine<Fq2>
    bool owns_pts = true;
    uint8_t* skip = nullptr;  // device: 1 if point i is infinity (b-queries hold many)
    int ba_levels = 0;      // batched-affine levels run before the XYZZ tail
};

struct MsmScratch {
    // -- digit stage: depends on the scalars only, so MSMs over the same scalars and skip pattern share it ----------
    size_t cap_items = 0;    // capacity in (window,point) pairs
    size_t cap_buckets = 0;  // capacity in buckets (sets * nb)
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
    uint32_t* hist = nullptr;       // radix histograms
    size_t hist_cap = 0;
    uint32_t* bucket_start = nullptr;  // sets*(nb+1)
    uint32_t* task_off = nullptr;      // per bucket: first task id (exclusive scan of tasks per bucket), +1
    uint32_t* task_tmp = nullptr;      // scan scratch
    uint32_t* tasks = nullptr;         // t_start | t_len | task-sort alternates
    uint32_t* counters = nullptr;      // [0] = number of tasks
    size_t cap_tasks = 0;
    // results of the last digit stage (the sort ping-pongs between buffers)
    uint32_t* s_vals = nullptr;        // point references sorted by bucket
    uint32_t *s_tkeys = nullptr, *s_tvals = nullptr;  // tasks sorted by length: (kTaskLen - len, task id)
    // batched-affine level tables (msm_affine.cu)
    int ba_levels = 0;
    size_t ba_stride = 0, ba_tstride = 0, ba_items = 0;
    uint32_t* ba_start0 = nullptr;     // first sorted position of every bucket
    uint32_t* ba_lvl = nullptr;        // [levels+1][nbuckets+1]: row 0 counts, rows 1.. offsets after each level
    uint32_t* ba_tb = nullptr;         // [levels][2][tstride]: first / last bucket of every level-kernel thread
    // -- point stage: per MSM ---------------------------------------------------------------------------------------------
    void *ba_buf_a = nullptr, *ba_buf_b = nullptr;  // affine points after odd / even levels
    void *ba_pre = nullptr, *ba_T = nullptr, *ba_Q = nullptr;  // Fq: prefix products per slot, thread totals, their inverses
    void* ba_den = nullptr;                                    // Fq: per-slot denominator norms (G2 only)
    void* partial = nullptr;           // XYZZ per task
    void* bucket_sum = nullptr;        // XYZZ per bucket
    void* chunk_sum = nullptr;         // XYZZ per chunk
    void* block_sum = nullptr;         // XYZZ per reduce block
    void* result = nullptr;            // XYZZ final (device)
    size_t point_bytes = 0;
};

struct CsrDev {
    uint64_t* row_ptr = nullptr;
    uint32_t* col = nullptr;
    Fr* val = nullptr;
    size_t nnz = 0;
    // sliced-ELL copy (witness.cu): rows sorted by length, 32 rows per slice, entries of a slice stored step-major
    uint32_t* sell_row = nullptr;   // [slices * 32] row evaluated by every lane (kSellNoRow = none)
    uint64_t* sell_ptr = nullptr;   // [slices + 1] first entry of every slice (multiples of 32)
    uint32_t* sell_col = nullptr;   // [sell_ptr[slices]] column | coefficient class << 30
    Fr* sell_val = nullptr;         // [sell_ptr[slices]] coefficient (read for class 0 only)
    size_t slices = 0;
};
constexpr uint32_t kSellNoRow = 0xFFFFFFFFu;

// ---- verifier state (verify.cu): the PreparedVerifyingKey of forks/groth16/src/data_structures.rs:62-72 on the device -----
struct VerifyKeyDev {
    bool have = false;
    size_t n_inputs = 0;           // gamma_abc_g1.len() - 1
    G1Affine* abc0 = nullptr;      // device: gamma_abc_g1[0]
    G1Affine* abc_tbl = nullptr;   // device: window tables of gamma_abc_g1[1..] (pairing.cuh: prepare_inputs_one)
    void* neg_gamma = nullptr;     // device: EllCoeff[kEllCoeffs] of -gamma_g2 (gamma_g2_neg_pc)
    void* neg_delta = nullptr;     // device: EllCoeff[kEllCoeffs] of -delta_g2 (delta_g2_neg_pc)
    void* alpha_beta = nullptr;    // device: Fq12 e(alpha_g1, beta_g2)
    uint64_t alpha_beta_host[48] = {};
    unsigned g2_inf = 0;           // bit 1 / 2: gamma_g2 / delta_g2 is the point at infinity (pair filtered out, as the reference does)
    // per-call scratch, grown on demand
    void* proofs = nullptr;        // g16_proof[cap]
    Fr* inputs = nullptr;          // [cap * n_inputs]
    G1Affine* prepared = nullptr;  // [cap]
    uint8_t* verdict = nullptr;    // [cap]
    size_t cap = 0;
};

// one captured launch sequence (api.cu: run_graphed)
struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    uint64_t key = ~0ull;     // what the sequence depends on (reduction, h source, option epoch)
    uint64_t launches = 0;    // kernel launches one replay stands for
    int seen = 0;             // runs with this key so far (first eager, second captured)
};

}  // namespace g16

struct g16_ctx {
    int device = 0;
    cudaStream_t main = nullptr;
    bool own_main = false;
    cudaStream_t side[g16::kSideStreams] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaStream_t wire = nullptr;  // root of the z-only MSM chains (forked from main, joined back at the end of a shard run)
    cudaEvent_t ev_wfork = nullptr, ev_wire_done = nullptr, ev_pre = nullptr;
    g16::GraphSlot graphs[8];     // FULL / WIRE / WM / H launch sequences, WM parts A / B / C / FINAL
    uint64_t graph_epoch = 1;     // bumped by everything that changes a sequence (options, key, R1CS)
    bool capturing = false;
    uint64_t graph_replays = 0, graph_captures = 0, graph_fallbacks = 0;
    int opt_graph = 1;            // replay the launch sequences as CUDA graphs (0: eager launches)
    void* h_proof = nullptr;      // page-locked landing zone of the proof read-back
    cudaEvent_t ev_scale[2] = {};
    cudaEvent_t ev_join[g16::kSideStreams] = {};
    cudaEvent_t ev_t[16] = {};
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;

    std::map<unsigned, g16::NttTables> ntt;

    // R1CS
    bool have_r1cs = false;
    uint64_t nc = 0, ni = 0, m = 0;
    unsigned log_n = 0;
    g16::CsrDev mat[3];
    g16::Fr *d_z = nullptr, *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;  // witness + three n-vectors
    uint64_t* h_pinned = nullptr;                                           // pinned staging for z
    size_t h_pinned_bytes = 0;
    bool witness_resident = false;  // all m elements of z are on the device
    bool witness_partial = false;   // at least this rank's wire-MSM slice of z is (g16_upload_witness_async(shard_only))
    bool l_on_a_space = false;      // l_query laid out on a_query[1..]'s index space (always, for a Groth16 key): l reads a's z slice

    // proving key
    bool have_pk = false;
    int shard_rank = 0, shard_count = 1;
    size_t pk_len[5] = {};  // full lengths of h, l, a, b_g1, b_g2 MSM ranges
    size_t sh_lo[5] = {}, sh_hi[5] = {};
    g16::MsmBases q[5];     // h, l, a(1..), b_g1(1..), b_g2(1..)
    g16::MsmScratch scratch[5];
    g16::G1Affine alpha_g1, beta_g1, delta_g1, a0, b1_0;  // host copies (Montgomery)
    g16::G2Affine beta_g2, delta_g2, b2_0;
    void* d_partial = nullptr;  // g16_partial on the device
    void* d_small = nullptr;    // small device scratch for assembly
    void* h_scalars = nullptr;  // page-locked staging of (r, s, rs) (assemble.cu), two alternating slots
    int h_scalars_slot = 0;
    void* d_asm_tables = nullptr;  // assemble.cu: per-key fixed-base tables of the (r, s)-only scalar multiplications
    int opt_asm_tables = 1;        // 0: the single-lane double-and-add k_assemble_pre
    g16_timings tm = {};
    int opt_serialize = 0, opt_kernel_events = 0, opt_window_bits = 0, opt_acc_variant = 0;
    int opt_ba_levels = -1;  // batched-affine levels (-1 = default)
    int opt_share_digits = 1;
    int opt_spmv_sell = 1;     // sliced-ELL SpMV (0: row-per-thread CSR kernel)
    int opt_ntt_batch = -1;    // witness map: a, b, c inverse transforms (and the a, b coset transforms) as one launch per pass;
                               // -1 = auto: only when nothing runs beside the transforms (like opt_ntt_radix4), 0 / 1 force
    int opt_ntt_radix4 = -1;   // k_ntt_pass4 (two butterfly levels per shared-memory round trip).  Alone it is faster (witness map
                               // 3.32 vs 3.67 ms) but beside the MSM chains it costs +1.1 ms per proof (profiles/r01_sched_sweep_*):
                               // -1 = auto: radix-4 when nothing runs beside the transforms (stand-alone calls, serialize, a
                               // shard rank without wire MSM work), radix-2 passes otherwise; 0 / 1 force one kernel
    bool wm_alone = true;      // state of the auto choice for the transforms being queued right now
    int opt_split_chains = 1;  // the MSM that reuses a digit stage runs beside the one that built it, not after it
    int opt_wm_priority = 0;   // witness map + h MSM on the internal high-priority stream
    int opt_wm_first = -1;     // the wire MSM chains start only when the witness map is done (it then runs alone);
                               // -1 = auto: for domains of 2^20 and more, 0 / 1 force
    cudaStream_t hi = nullptr;  // high-priority twin of main
    cudaEvent_t ev_dig[2] = {}, ev_hi = nullptr;
    bool share_al = false, share_b = false;  // l reuses a's digit stage / b_g2 reuses b_g1's
    cudaStream_t sh_wm_stream = nullptr;      // stream the witness map of the open shard run was queued on
    bool shard_open = false;
    bool tm_stale = false;  // events of a *_dev shard run not yet read into tm
    cudaEvent_t ev_acc[10] = {};
    bool pre_pending = false;  // k_assemble_pre already in flight for (pre_r, pre_s)
    uint64_t pre_r[4] = {}, pre_s[4] = {};

    g16::VerifyKeyDev vk;  // row f-4
    int opt_verify_occupancy = 8;  // resident warps per SM k_verify is compiled for (8 / 12 / 16)

    // generic MSM slots
    g16::MsmBases slot[g16::kMsmSlots];
    g16::MsmScratch slot_scratch[g16::kMsmSlots];
};

namespace g16 {

// error helpers -----------------------------------------------------------------------------------------------------
int set_err(g16_ctx* ctx, int code, const char* fmt, ...);
#define G16_CUDA(ctx, expr)                                                                            \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return g16::set_err(ctx, _e == cudaErrorMemoryAllocation ? G16_ERR_OOM : G16_ERR_CUDA,     \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define G16_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != G16_OK) return _rc; \
    } while (0)
#define G16_LAUNCH(ctx, kern, grid, block, smem, stream, ...)                       \
    do {                                                                            \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                   \
        (ctx)->launches++;                                                          \
        G16_CUDA(ctx, cudaPeekAtLastError());                                       \
    } while (0)

template <class T>
int dev_alloc(g16_ctx* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return G16_OK;
    G16_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
    return G16_OK;
}
inline void dev_free(void* p) {
    if (p) cudaFree(p);
}

// field_ops.cu ------------------------------------------------------------------------------------------------------
int field_op_dev(g16_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n, cudaStream_t st);
int bench_int_pipe(g16_ctx* ctx, int which, double* gops);
int convert_mont_dev(g16_ctx* ctx, int field /*FR|FQ*/, void* data, size_t n_elems, bool to_mont, cudaStream_t st);

// ntt.cu --------------------------------------------------------------------------------------------------------------
int ntt_get_tables(g16_ctx* ctx, unsigned log_n, NttTables** out);
// core transforms on device data (no bit reversal):  dif: natural -> bit-reversed;  dit: bit-reversed -> natural.
// pre (dif) / post (dit) are optional element-wise scale tables indexed by the natural index; post_scalar (dit) is an
// optional uniform Montgomery factor.
int ntt_dif(g16_ctx* ctx, Fr* data, NttTables* t, bool inverse_root, const Fr* pre, cudaStream_t st);
// pre_mul (dit): optional vector multiplied in element-wise on the first load (same physical order as data);
// post_sub (dit, with post): optional vector subtracted after the post scaling (natural order).
int ntt_dit(g16_ctx* ctx, Fr* data, NttTables* t, bool inverse_root, const Fr* post, const Fr* post_scalar,
            cudaStream_t st, const Fr* pre_mul = nullptr, const Fr* post_sub = nullptr);
// the same transforms for up to three vectors in one launch per pass (member m of scalar_mask gets post_scalar)
int ntt_dit_batch(g16_ctx* ctx, Fr* const* data, int count, NttTables* t, bool inverse_root, const Fr* post, const Fr* post_scalar,
                  unsigned scalar_mask, cudaStream_t st, const Fr* pre_mul = nullptr, const Fr* post_sub = nullptr);
int ntt_dif_batch(g16_ctx* ctx, Fr* const* data, int count, NttTables* t, bool inverse_root, const Fr* pre, cudaStream_t st);
int bitrev_permute(g16_ctx* ctx, Fr* data, unsigned log_n, cudaStream_t st);
int pow_table_dev(g16_ctx* ctx, Fr* out, size_t n, Fr base, Fr scale, cudaStream_t st);
int ntt_api(g16_ctx* ctx, Fr* data_dev, unsigned log_n, int inverse, int coset, cudaStream_t st);

// witness.cu ------------------------------------------------------------------------------------------------------------
int witness_map_dev(g16_ctx* ctx, int reduction, cudaStream_t st);  // z in ctx->d_z; h left in ctx->d_a (natural order)
int witness_map_part_dev(g16_ctx* ctx, int parts, cudaStream_t st);  // LibsnarkReduction in parts (G16_WM_PART_*)
int r1cs_eval_dev(g16_ctx* ctx, Fr* az, Fr* bz, Fr* cz, bool bitrev, cudaStream_t st);
// builds the sliced-ELL copy of matrix k from its CSR arrays (host row_ptr for the row order, device col / val for the entries)
int sell_build(g16_ctx* ctx, int k, const uint64_t* row_ptr_host, cudaStream_t st);
void sell_free(CsrDev* m);

// msm.cu ----------------------------------------------------------------------------------------------------------------
int msm_pick_window(size_t n, int group, bool precomp);
int msm_set_bases(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, int group, const void* pts_dev, size_t n, int c,
                  bool precomp, cudaStream_t st);
void msm_free(MsmBases* mb, MsmScratch* sc);
// runs Pippenger over the first n bases with device scalars (Montgomery Fr); leaves the XYZZ result in sc->result
// `digits` (optional): scratch of another MSM that already ran over the SAME scalars on this stream, with the same
// n / window / table geometry and skip pattern (msm_can_share): its sorted references and task lists are reused.
int msm_run(g16_ctx* ctx, MsmBases* mb, MsmScratch* sc, const Fr* scalars_dev, size_t n, cudaStream_t st,
            cudaEvent_t ev_acc0 = nullptr, cudaEvent_t ev_acc1 = nullptr, const MsmScratch* digits = nullptr,
            cudaEvent_t ev_digits_done = nullptr);  // recorded on st once the digit stage this call built is complete
// true when MSMs over `b` may reuse the digit stage of `a` (same geometry; every point `a` skips is infinity in `b` too
// and `b` has at most `max_extra_inf` further points at infinity).  Synchronises the stream.
int msm_can_share(g16_ctx* ctx, const MsmBases* a, const MsmBases* b, size_t max_extra_inf, bool* ok, cudaStream_t st);
int fixed_base_dev(g16_ctx* ctx, int group, const Fr* scalars_dev, size_t n, void* out_pts_dev, cudaStream_t st);

// assemble.cu -----------------------------------------------------------------------------------------------------------
int assemble_set_scalars(g16_ctx* ctx, const uint64_t* r, const uint64_t* s, cudaStream_t st);  // (r, s, rs) -> device, stream-ordered
int assemble_pre(g16_ctx* ctx, cudaStream_t st);
int assemble_build_tables(g16_ctx* ctx, cudaStream_t st);
G1Affine g1_generator();
G2Affine g2_generator();
// out = k * in for one G1 XYZZ point on the device; k = r (which = 0) or s (which = 1) of the scalars set by assemble_set_scalars
int scale_point_dev(g16_ctx* ctx, const void* in_xyzz, int which, void* out_xyzz, cudaStream_t st);
// k_assemble_post + the read-back into the context's page-locked block (stream-ordered); _read unpacks it after a synchronise
int assemble_proof_queue(g16_ctx* ctx, const void* partials_dev, int count, cudaStream_t st);
void assemble_proof_read(g16_ctx* ctx, g16_proof* out);
int xyzz_to_affine_host(g16_ctx* ctx, int group, const void* xyzz_dev, uint64_t* out, int* out_inf, cudaStream_t st);
// sum of `count` XYZZ points -> one affine point on the host (the combine step of a point-range sharded MSM)
int sum_partials_to_affine_host(g16_ctx* ctx, int group, const void* xyzz_dev, int count, uint64_t* out, int* out_inf, cudaStream_t st);

// verify.cu ---------------------------------------------------------------------------------------------------------------
void verify_free(g16_ctx* ctx);

}  // namespace g16
