// api.cu -- extern "C" entry points of libg16b200.so (declared in include/g16_b200.h) and the prove orchestration:
// stream fork/join of the witness map, the five MSMs and the assembly, mirroring
// Groth16::create_proof_with_reduction_and_matrices (forks/groth16/src/prover.rs:26-51).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "internal.cuh"

using namespace g16;

static thread_local std::string g_create_err;

namespace g16 {
int set_err(g16_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_err = buf;
    return code;
}
}  // namespace g16

// MSM order used throughout: 0 = h, 1 = l, 2 = a, 3 = b_g1, 4 = b_g2
enum { Q_H = 0, Q_L = 1, Q_A = 2, Q_B1 = 3, Q_B2 = 4 };

struct Guard {
    g16_ctx* c;
    explicit Guard(g16_ctx* ctx) : c(ctx) {
        c->mu.lock();
        cudaSetDevice(c->device);
    }
    ~Guard() { c->mu.unlock(); }
};

extern "C" {

const char* g16_version(void) { return "g16-b200 0.1 (sm_100a)"; }

int g16_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* g16_last_error(const g16_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int g16_ctx_create(g16_ctx** out, int device, void* main_stream) {
    if (!out) return set_err(nullptr, G16_ERR_BAD_ARG, "g16_ctx_create: out is NULL");
    *out = nullptr;
    int n = g16_device_count();
    if (n <= 0) return set_err(nullptr, G16_ERR_NO_DEVICE, "no CUDA device visible: libg16b200 has no CPU fallback");
    if (device < 0 || device >= n) return set_err(nullptr, G16_ERR_BAD_ARG, "device %d out of range (%d visible)", device, n);
    g16_ctx* ctx = new g16_ctx();
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        delete ctx;
        return set_err(nullptr, G16_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    }
    // All streams share one priority: giving the main stream (witness map -> h MSM -> assembly) a higher one was measured
    // 2 % slower on one GPU -- the side chains' latency-bound tails then pile up at the end of the proof.
    if (main_stream) {
        ctx->main = (cudaStream_t)main_stream;
        ctx->own_main = false;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->main, cudaStreamNonBlocking);
        ctx->own_main = true;
    }
    for (int i = 0; i < kSideStreams && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;  // numerically lower = higher priority
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->hi, cudaStreamNonBlocking, hi);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_dig[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_hi, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->wire, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_wfork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_wire_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pre, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_proof, 512, cudaHostAllocDefault);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_scale[i], cudaEventDisableTiming);
    for (int i = 0; i < 16 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev_t[i]);
    for (int i = 0; i < 10 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev_acc[i]);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_small, 4096);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_partial, sizeof(g16_partial));
    if (e != cudaSuccess) {
        int rc = set_err(nullptr, e == cudaErrorMemoryAllocation ? G16_ERR_OOM : G16_ERR_CUDA, "context setup: %s",
                         cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return G16_OK;
}

static void graphs_drop(g16_ctx* ctx);

static void free_r1cs(g16_ctx* ctx) {
    ctx->graph_epoch++;
    for (int k = 0; k < 3; k++) {
        dev_free(ctx->mat[k].row_ptr);
        dev_free(ctx->mat[k].col);
        dev_free(ctx->mat[k].val);
        sell_free(&ctx->mat[k]);
        ctx->mat[k] = CsrDev();
    }
    dev_free(ctx->d_z);
    dev_free(ctx->d_a);
    dev_free(ctx->d_b);
    dev_free(ctx->d_c);
    ctx->d_z = ctx->d_a = ctx->d_b = ctx->d_c = nullptr;
    ctx->have_r1cs = false;
    ctx->witness_resident = false;
    ctx->witness_partial = false;
}

void g16_ctx_destroy(g16_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_r1cs(ctx);
    for (int k = 0; k < 5; k++) msm_free(&ctx->q[k], &ctx->scratch[k]);
    for (int k = 0; k < kMsmSlots; k++) msm_free(&ctx->slot[k], &ctx->slot_scratch[k]);
    for (auto& kv : ctx->ntt) {
        dev_free(kv.second.tw);
        dev_free(kv.second.tw_inv);
        dev_free(kv.second.coset);
        dev_free(kv.second.coset_inv);
        dev_free(kv.second.coset_scaled);
        dev_free(kv.second.odd_scaled);
        dev_free(kv.second.coset_inv_z);
    }
    verify_free(ctx);
    dev_free(ctx->d_small);
    dev_free(ctx->d_asm_tables);
    dev_free(ctx->d_partial);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int i = 0; i < kSideStreams; i++) {
        if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    graphs_drop(ctx);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->wire) cudaStreamDestroy(ctx->wire);
    if (ctx->ev_wfork) cudaEventDestroy(ctx->ev_wfork);
    if (ctx->ev_wire_done) cudaEventDestroy(ctx->ev_wire_done);
    if (ctx->ev_pre) cudaEventDestroy(ctx->ev_pre);
    if (ctx->h_proof) cudaFreeHost(ctx->h_proof);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->hi) cudaStreamDestroy(ctx->hi);
    if (ctx->ev_hi) cudaEventDestroy(ctx->ev_hi);
    for (int i = 0; i < 2; i++)
        if (ctx->ev_dig[i]) cudaEventDestroy(ctx->ev_dig[i]);
    for (int i = 0; i < 2; i++)
        if (ctx->ev_scale[i]) cudaEventDestroy(ctx->ev_scale[i]);
    for (int i = 0; i < 16; i++)
        if (ctx->ev_t[i]) cudaEventDestroy(ctx->ev_t[i]);
    for (int i = 0; i < 10; i++)
        if (ctx->ev_acc[i]) cudaEventDestroy(ctx->ev_acc[i]);
    if (ctx->own_main && ctx->main) cudaStreamDestroy(ctx->main);
    delete ctx;
}

uint64_t g16_launch_count(const g16_ctx* ctx) { return ctx ? ctx->launches : 0; }

int g16_graph_stats(const g16_ctx* ctx, uint64_t out[3]) {
    if (!ctx || !out) return G16_ERR_BAD_ARG;
    out[0] = ctx->graph_replays;    // cudaGraphLaunch calls so far
    out[1] = ctx->graph_captures;   // launch sequences captured
    out[2] = ctx->graph_fallbacks;  // captures that failed (the context then queues eagerly)
    return G16_OK;
}

int g16_sync(g16_ctx* ctx) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}

// ---- device memory ----------------------------------------------------------------------------------------------------
int g16_host_alloc(g16_ctx* ctx, size_t bytes, void** p) {
    if (!ctx || !p) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return G16_OK;
}
int g16_host_free(g16_ctx* ctx, void* p) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (p) G16_CUDA(ctx, cudaFreeHost(p));
    return G16_OK;
}
int g16_dev_alloc(g16_ctx* ctx, size_t bytes, void** p) {
    if (!ctx || !p) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaMalloc(p, bytes ? bytes : 1));
    return G16_OK;
}
int g16_dev_free(g16_ctx* ctx, void* p) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaFree(p));
    return G16_OK;
}
int g16_dev_upload(g16_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}
int g16_dev_download(g16_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}

// ---- building blocks ----------------------------------------------------------------------------------------------------
int g16_field_op(g16_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    if (!ctx || !a || !out) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "field_op: null pointer") : G16_ERR_BAD_ARG;
    bool bcast = (op == G16_OP_MUL_BCAST || op == G16_OP_ADD_BCAST);
    bool binary = (op == G16_OP_MUL || op == G16_OP_ADD || op == G16_OP_SUB || bcast);
    if (binary && !b) return set_err(ctx, G16_ERR_BAD_ARG, "field_op: binary op needs b");
    if (n == 0) return G16_OK;
    Guard g(ctx);
    size_t esz = field == G16_FIELD_FQ2 ? 64 : 32;
    void *da = nullptr, *db = nullptr, *dc = nullptr;
    G16_CUDA(ctx, cudaMalloc(&da, n * esz));
    G16_CUDA(ctx, cudaMalloc(&db, n * esz));
    G16_CUDA(ctx, cudaMalloc(&dc, n * esz));
    int rc = G16_OK;
    cudaMemcpyAsync(da, a, n * esz, cudaMemcpyHostToDevice, ctx->main);
    if (binary) cudaMemcpyAsync(db, b, (bcast ? 1 : n) * esz, cudaMemcpyHostToDevice, ctx->main);
    rc = field_op_dev(ctx, field, op, da, db, dc, n, ctx->main);
    if (rc == G16_OK) {
        cudaMemcpyAsync(out, dc, n * esz, cudaMemcpyDeviceToHost, ctx->main);
        cudaError_t e = cudaStreamSynchronize(ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "field_op: %s", cudaGetErrorString(e));
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dc);
    return rc;
}

int g16_bench_int_pipe(g16_ctx* ctx, int which, double* gops) {
    if (!ctx || !gops) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return bench_int_pipe(ctx, which, gops);
}

int g16_ntt_dev(g16_ctx* ctx, void* data_dev, unsigned log_n, int inverse, int coset) {
    if (!ctx || !data_dev) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return ntt_api(ctx, (Fr*)data_dev, log_n, inverse, coset, ctx->main);
}

int g16_ntt(g16_ctx* ctx, uint64_t* data, unsigned log_n, int inverse, int coset) {
    if (!ctx || !data) return G16_ERR_BAD_ARG;
    if (log_n > 28) return set_err(ctx, G16_ERR_DEGREE_TOO_LARGE, "domain 2^%u exceeds Fr two-adicity 28", log_n);
    Guard g(ctx);
    size_t bytes = ((size_t)32) << log_n;
    Fr* d;
    G16_CUDA(ctx, cudaMalloc((void**)&d, bytes));
    cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->main);
    int rc = ntt_api(ctx, d, log_n, inverse, coset, ctx->main);
    if (rc == G16_OK) {
        cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->main);
        cudaError_t e = cudaStreamSynchronize(ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "ntt: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

static int msm_host(g16_ctx* ctx, int group, const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out,
                    int* out_inf) {
    if (!ctx || !out || (n && (!points || !scalars))) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm: null pointer") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    size_t pb = group == 1 ? 64 : 128;
    if (n == 0) {
        memset(out, 0, pb);
        if (out_inf) *out_inf = 1;
        return G16_OK;
    }
    void* dp = nullptr;
    Fr* ds = nullptr;
    G16_CUDA(ctx, cudaMalloc(&dp, n * pb));
    G16_CUDA(ctx, cudaMalloc((void**)&ds, n * 32));
    cudaMemcpyAsync(dp, points, n * pb, cudaMemcpyHostToDevice, ctx->main);
    cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->main);
    MsmBases mb;
    MsmScratch sc;
    int rc = msm_set_bases(ctx, &mb, &sc, group, dp, n, ctx->opt_window_bits, false, ctx->main);
    if (rc == G16_OK) rc = msm_run(ctx, &mb, &sc, ds, n, ctx->main);
    if (rc == G16_OK) rc = xyzz_to_affine_host(ctx, group, sc.result, out, out_inf, ctx->main);
    cudaStreamSynchronize(ctx->main);
    msm_free(&mb, &sc);
    cudaFree(dp);
    cudaFree(ds);
    return rc;
}
int g16_msm_g1(g16_ctx* ctx, const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t out[8], int* out_inf) {
    return msm_host(ctx, 1, points, scalars, n, out, out_inf);
}
int g16_msm_g2(g16_ctx* ctx, const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t out[16], int* out_inf) {
    return msm_host(ctx, 2, points, scalars, n, out, out_inf);
}

int g16_msm_set_bases_dev(g16_ctx* ctx, int slot, int group, const void* points_dev, size_t n, int window_bits, int precompute) {
    if (!ctx || slot < 0 || slot >= kMsmSlots) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm slot out of range") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_TRY(msm_set_bases(ctx, &ctx->slot[slot], &ctx->slot_scratch[slot], group, points_dev, n, window_bits, precompute != 0,
                          ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}
int g16_msm_set_bases(g16_ctx* ctx, int slot, int group, const uint64_t* points, size_t n, int window_bits, int precompute) {
    if (!ctx || (n && !points)) return G16_ERR_BAD_ARG;
    if (group != 1 && group != 2) return set_err(ctx, G16_ERR_BAD_ARG, "group must be 1 or 2");
    size_t pb = group == 1 ? 64 : 128;
    void* dp = nullptr;
    {
        Guard g(ctx);
        G16_CUDA(ctx, cudaMalloc(&dp, n ? n * pb : 1));
        G16_CUDA(ctx, cudaMemcpy(dp, points, n * pb, cudaMemcpyHostToDevice));
    }
    int rc = g16_msm_set_bases_dev(ctx, slot, group, dp, n, window_bits, precompute);
    cudaFree(dp);
    return rc;
}
int g16_msm_run_dev(g16_ctx* ctx, int slot, const void* scalars_dev, size_t n, uint64_t* out, int* out_inf) {
    if (!ctx || slot < 0 || slot >= kMsmSlots) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm slot out of range") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    MsmBases* mb = &ctx->slot[slot];
    if (mb->group == 0) return set_err(ctx, G16_ERR_BAD_ARG, "msm slot %d has no bases", slot);
    G16_TRY(msm_run(ctx, mb, &ctx->slot_scratch[slot], (const Fr*)scalars_dev, n, ctx->main,
                    ctx->opt_kernel_events ? ctx->ev_acc[0] : nullptr, ctx->opt_kernel_events ? ctx->ev_acc[1] : nullptr));
    if (ctx->opt_kernel_events) {
        G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
        float ms = -1;
        if (cudaEventElapsedTime(&ms, ctx->ev_acc[0], ctx->ev_acc[1]) != cudaSuccess) cudaGetLastError();
        ctx->tm.acc_ms[0] = ms;
    }
    if (out) return xyzz_to_affine_host(ctx, mb->group, ctx->slot_scratch[slot].result, out, out_inf, ctx->main);
    return G16_OK;
}

int g16_msm_window_bits(size_t n, int precompute) { return msm_pick_window(n, 1, precompute != 0); }

int g16_msm_copy_result_dev(g16_ctx* ctx, int slot, void* dst_dev) {
    if (!ctx || !dst_dev || slot < 0 || slot >= kMsmSlots) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm_copy_result: bad argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    const MsmBases* mb = &ctx->slot[slot];
    if (mb->group == 0) return set_err(ctx, G16_ERR_BAD_ARG, "msm slot %d has no bases", slot);
    const size_t bytes = mb->group == 1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ);
    if (ctx->slot_scratch[slot].result) G16_CUDA(ctx, cudaMemcpyAsync(dst_dev, ctx->slot_scratch[slot].result, bytes, cudaMemcpyDeviceToDevice, ctx->main));
    else G16_CUDA(ctx, cudaMemsetAsync(dst_dev, 0, bytes, ctx->main));  // an empty share: the point at infinity (zz = 0)
    return G16_OK;
}
int g16_msm_combine_dev(g16_ctx* ctx, int group, const void* partials_dev, int count, uint64_t* out, int* out_inf) {
    if (!ctx || !partials_dev || !out || count < 1 || (group != 1 && group != 2))
        return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm_combine: bad argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    return sum_partials_to_affine_host(ctx, group, partials_dev, count, out, out_inf, ctx->main);
}

static int fixed_base_host(g16_ctx* ctx, int group, const uint64_t* scalars, size_t n, uint64_t* out) {
    if (!ctx || (n && (!scalars || !out))) return G16_ERR_BAD_ARG;
    if (n == 0) return G16_OK;
    Guard g(ctx);
    size_t pb = group == 1 ? 64 : 128;
    Fr* ds;
    void* dp;
    G16_CUDA(ctx, cudaMalloc((void**)&ds, n * 32));
    G16_CUDA(ctx, cudaMalloc(&dp, n * pb));
    cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->main);
    int rc = fixed_base_dev(ctx, group, ds, n, dp, ctx->main);
    if (rc == G16_OK) {
        cudaMemcpyAsync(out, dp, n * pb, cudaMemcpyDeviceToHost, ctx->main);
        cudaError_t e = cudaStreamSynchronize(ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "fixed_base: %s", cudaGetErrorString(e));
    }
    cudaFree(ds);
    cudaFree(dp);
    return rc;
}
int g16_fixed_base_g1(g16_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out) { return fixed_base_host(ctx, 1, scalars, n, out); }
int g16_fixed_base_g2(g16_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out) { return fixed_base_host(ctx, 2, scalars, n, out); }
int g16_fixed_base_g1_dev(g16_ctx* ctx, const void* s, size_t n, void* out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return fixed_base_dev(ctx, 1, (const Fr*)s, n, out, ctx->main);
}
int g16_fixed_base_g2_dev(g16_ctx* ctx, const void* s, size_t n, void* out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return fixed_base_dev(ctx, 2, (const Fr*)s, n, out, ctx->main);
}

// ---- R1CS -------------------------------------------------------------------------------------------------------------------
int g16_ctx_load_r1cs(g16_ctx* ctx, const g16_r1cs_view* v) {
    if (!ctx || !v) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (v->num_instance == 0 || v->num_instance > v->num_wires)
        return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: num_instance %llu / num_wires %llu inconsistent",
                       (unsigned long long)v->num_instance, (unsigned long long)v->num_wires);
    if (v->num_wires >= ((uint64_t)1 << 32)) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: more than 2^32 wires");
    // domain: smallest power of two >= num_constraints + num_instance (r1cs_to_qap.rs:156-158)
    uint64_t need = v->num_constraints + v->num_instance;
    unsigned log_n = 0;
    while (((uint64_t)1 << log_n) < need) log_n++;
    if (log_n > 28) return set_err(ctx, G16_ERR_DEGREE_TOO_LARGE, "domain size %llu exceeds 2^28", (unsigned long long)need);
    for (int k = 0; k < 3; k++) {
        if (v->num_constraints && (!v->row_ptr[k])) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: null row_ptr");
        uint64_t nnz = v->num_constraints ? v->row_ptr[k][v->num_constraints] : 0;
        if (nnz && (!v->col[k] || !v->val[k])) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: null col/val");
        if (v->num_constraints && v->row_ptr[k][0] != 0) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: row_ptr[0] != 0");
        for (uint64_t i = 0; i < v->num_constraints; i++)
            if (v->row_ptr[k][i + 1] < v->row_ptr[k][i]) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: row_ptr not monotone");
        for (uint64_t j = 0; j < nnz; j++)
            if (v->col[k][j] >= v->num_wires) return set_err(ctx, G16_ERR_BAD_ARG, "r1cs: column index out of range");
    }
    free_r1cs(ctx);
    ctx->nc = v->num_constraints;
    ctx->ni = v->num_instance;
    ctx->m = v->num_wires;
    ctx->log_n = log_n;
    size_t n = (size_t)1 << log_n;
    for (int k = 0; k < 3; k++) {
        uint64_t nnz = ctx->nc ? v->row_ptr[k][ctx->nc] : 0;
        ctx->mat[k].nnz = nnz;
        G16_TRY(dev_alloc(ctx, &ctx->mat[k].row_ptr, ctx->nc + 1));
        G16_TRY(dev_alloc(ctx, &ctx->mat[k].col, nnz));
        G16_TRY(dev_alloc(ctx, &ctx->mat[k].val, nnz));
        if (ctx->nc) G16_CUDA(ctx, cudaMemcpyAsync(ctx->mat[k].row_ptr, v->row_ptr[k], (ctx->nc + 1) * 8, cudaMemcpyHostToDevice, ctx->main));
        else G16_CUDA(ctx, cudaMemsetAsync(ctx->mat[k].row_ptr, 0, 8, ctx->main));
        if (nnz) {
            G16_CUDA(ctx, cudaMemcpyAsync(ctx->mat[k].col, v->col[k], nnz * 4, cudaMemcpyHostToDevice, ctx->main));
            G16_CUDA(ctx, cudaMemcpyAsync(ctx->mat[k].val, v->val[k], nnz * 32, cudaMemcpyHostToDevice, ctx->main));
            if (v->encoding == G16_ENC_CANONICAL) G16_TRY(convert_mont_dev(ctx, G16_FIELD_FR, ctx->mat[k].val, nnz, true, ctx->main));
        }
        if (ctx->nc && ctx->m < ((uint64_t)1 << 30)) G16_TRY(sell_build(ctx, k, v->row_ptr[k], ctx->main));
    }
    G16_TRY(dev_alloc(ctx, &ctx->d_z, ctx->m));
    G16_TRY(dev_alloc(ctx, &ctx->d_a, n));
    G16_TRY(dev_alloc(ctx, &ctx->d_b, n));
    G16_TRY(dev_alloc(ctx, &ctx->d_c, n));
    NttTables* t;
    G16_TRY(ntt_get_tables(ctx, log_n, &t));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    ctx->have_r1cs = true;
    return G16_OK;
}

int g16_domain_size(g16_ctx* ctx, size_t* n_out) {
    if (!ctx || !n_out) return G16_ERR_BAD_ARG;
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "no R1CS loaded");
    *n_out = (size_t)1 << ctx->log_n;
    return G16_OK;
}

static int upload_witness(g16_ctx* ctx, const uint64_t* z, cudaStream_t st) {
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "no R1CS loaded");
    if (!z) return set_err(ctx, G16_ERR_BAD_ARG, "witness pointer is NULL");
    G16_CUDA(ctx, cudaMemcpyAsync(ctx->d_z, z, ctx->m * 32, cudaMemcpyHostToDevice, st));
    ctx->witness_resident = true;
    ctx->witness_partial = true;
    return G16_OK;
}

int g16_upload_witness(g16_ctx* ctx, const uint64_t* z) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_TRY(upload_witness(ctx, z, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}

int g16_upload_witness_async(g16_ctx* ctx, const uint64_t* z, int shard_only) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "no R1CS loaded");
    if (!z) return set_err(ctx, G16_ERR_BAD_ARG, "witness pointer is NULL");
    if (!shard_only) return upload_witness(ctx, z, ctx->main);
    if (!ctx->have_pk) return set_err(ctx, G16_ERR_BAD_ARG, "upload_witness_async(shard_only): no proving key loaded");
    // the wire MSMs of this rank read z[1 + lo, 1 + hi) (a / b_g1 / b_g2, prover.rs:84-89) and, when l_query is not laid out on
    // a's index space, z[ni + l_lo, ni + l_hi) (l, prover.rs:70-74): nothing else of the 32 * m bytes has to cross PCIe
    ctx->witness_resident = false;
    const size_t lo = ctx->sh_lo[Q_A], hi = ctx->sh_hi[Q_A];
    if (hi > lo)
        G16_CUDA(ctx, cudaMemcpyAsync(ctx->d_z + 1 + lo, z + 4 * (1 + lo), (hi - lo) * 32, cudaMemcpyHostToDevice, ctx->main));
    if (!ctx->l_on_a_space && ctx->sh_hi[Q_L] > ctx->sh_lo[Q_L])
        G16_CUDA(ctx, cudaMemcpyAsync(ctx->d_z + ctx->ni + ctx->sh_lo[Q_L], z + 4 * (ctx->ni + ctx->sh_lo[Q_L]),
                                      (ctx->sh_hi[Q_L] - ctx->sh_lo[Q_L]) * 32, cudaMemcpyHostToDevice, ctx->main));
    ctx->witness_partial = true;
    return G16_OK;
}

int g16_memcpy_h2d_async(void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    if (!dst_dev || !src_host) return G16_ERR_BAD_ARG;
    return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream) == cudaSuccess ? G16_OK : G16_ERR_CUDA;
}

int g16_upload_witness_dev(g16_ctx* ctx, const void* z_dev) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "no R1CS loaded");
    if (!z_dev) return set_err(ctx, G16_ERR_BAD_ARG, "witness pointer is NULL");
    G16_CUDA(ctx, cudaMemcpyAsync(ctx->d_z, z_dev, ctx->m * 32, cudaMemcpyDeviceToDevice, ctx->main));
    ctx->witness_resident = true;
    ctx->witness_partial = true;
    return G16_OK;
}

int g16_r1cs_eval(g16_ctx* ctx, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_TRY(upload_witness(ctx, z, ctx->main));
    size_t n = (size_t)1 << ctx->log_n;
    G16_CUDA(ctx, cudaMemsetAsync(ctx->d_a, 0, n * 32, ctx->main));
    G16_CUDA(ctx, cudaMemsetAsync(ctx->d_b, 0, n * 32, ctx->main));
    G16_CUDA(ctx, cudaMemsetAsync(ctx->d_c, 0, n * 32, ctx->main));
    G16_TRY(r1cs_eval_dev(ctx, az ? ctx->d_a : nullptr, bz ? ctx->d_b : nullptr, cz ? ctx->d_c : nullptr, false, ctx->main));
    if (az) G16_CUDA(ctx, cudaMemcpyAsync(az, ctx->d_a, ctx->nc * 32, cudaMemcpyDeviceToHost, ctx->main));
    if (bz) G16_CUDA(ctx, cudaMemcpyAsync(bz, ctx->d_b, ctx->nc * 32, cudaMemcpyDeviceToHost, ctx->main));
    if (cz) G16_CUDA(ctx, cudaMemcpyAsync(cz, ctx->d_c, ctx->nc * 32, cudaMemcpyDeviceToHost, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}

int g16_witness_map(g16_ctx* ctx, const uint64_t* z, int reduction, uint64_t* h_out, size_t h_capacity, size_t* n_out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "no R1CS loaded");
    size_t n = (size_t)1 << ctx->log_n;
    if (n_out) *n_out = n;
    if (!h_out || h_capacity < n) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map: h buffer holds %zu < %zu elements", h_capacity, n);
    G16_TRY(upload_witness(ctx, z, ctx->main));
    G16_TRY(witness_map_dev(ctx, reduction, ctx->main));
    G16_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_a, n * 32, cudaMemcpyDeviceToHost, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return G16_OK;
}

// ---- proving key ---------------------------------------------------------------------------------------------------------------
static void load_fq(Fq* dst, const uint64_t* src, int count, int enc) {
    for (int k = 0; k < count; k++) {
        Fq x;
        for (int i = 0; i < 4; i++) {
            x.v[2 * i] = (uint32_t)src[4 * k + i];
            x.v[2 * i + 1] = (uint32_t)(src[4 * k + i] >> 32);
        }
        dst[k] = enc == G16_ENC_CANONICAL ? x.to_mont() : x;
    }
}

// uploads points [lo, hi) of a host query (skipping `skip_first` leading points) into an MSM base set
static int load_query(g16_ctx* ctx, int qi, int group, const uint64_t* pts, size_t total, size_t lo, size_t hi, int enc,
                      int precompute) {
    size_t pb = group == 1 ? 64 : 128;
    size_t cnt = hi - lo;
    ctx->pk_len[qi] = total;
    ctx->sh_lo[qi] = lo;
    ctx->sh_hi[qi] = hi;
    void* stage = nullptr;
    G16_CUDA(ctx, cudaMalloc(&stage, cnt ? cnt * pb : 1));
    int rc = G16_OK;
    if (cnt) {
        cudaError_t e = cudaMemcpyAsync(stage, (const char*)pts + lo * pb, cnt * pb, cudaMemcpyHostToDevice, ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "pk upload: %s", cudaGetErrorString(e));
        if (rc == G16_OK && enc == G16_ENC_CANONICAL) rc = convert_mont_dev(ctx, G16_FIELD_FQ, stage, cnt * (pb / 32), true, ctx->main);
    }
    if (rc == G16_OK) rc = msm_set_bases(ctx, &ctx->q[qi], &ctx->scratch[qi], group, stage, cnt, ctx->opt_window_bits, precompute != 0, ctx->main);
    cudaStreamSynchronize(ctx->main);
    cudaFree(stage);
    return rc;
}

// h_range / z_range (optional): this rank's point range of h_query and of the a / b_g1 / b_g2 queries (index space of
// query[1..]); NULL = the uniform split [rank*N/G, (rank+1)*N/G).
static int load_pk_ranges(g16_ctx* ctx, const g16_pk_view* pk, int shard_rank, int shard_count, const uint64_t* h_range,
                          const uint64_t* z_range, int precompute) {
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
        return set_err(ctx, G16_ERR_BAD_ARG, "shard %d of %d is not a valid rank", shard_rank, shard_count);
    if (!pk->alpha_g1 || !pk->beta_g1 || !pk->delta_g1 || !pk->beta_g2 || !pk->delta_g2)
        return set_err(ctx, G16_ERR_BAD_ARG, "pk: a single-point pointer is NULL");
    if (pk->a_len == 0 || pk->b_g1_len != pk->a_len || pk->b_g2_len != pk->a_len)
        return set_err(ctx, G16_ERR_BAD_ARG, "pk: a/b_g1/b_g2 query lengths %zu/%zu/%zu must be equal and non-zero", pk->a_len,
                       pk->b_g1_len, pk->b_g2_len);
    if (!pk->a_query || !pk->b_g1_query || !pk->b_g2_query || (pk->h_len && !pk->h_query) || (pk->l_len && !pk->l_query))
        return set_err(ctx, G16_ERR_BAD_ARG, "pk: a query pointer is NULL");
    ctx->have_pk = false;
    ctx->graph_epoch++;
    ctx->shard_rank = shard_rank;
    ctx->shard_count = shard_count;
    int enc = pk->encoding;
    load_fq(&ctx->alpha_g1.x, pk->alpha_g1, 2, enc);
    load_fq(&ctx->beta_g1.x, pk->beta_g1, 2, enc);
    load_fq(&ctx->delta_g1.x, pk->delta_g1, 2, enc);
    load_fq(&ctx->beta_g2.x.c0, pk->beta_g2, 4, enc);
    load_fq(&ctx->delta_g2.x.c0, pk->delta_g2, 4, enc);
    load_fq(&ctx->a0.x, pk->a_query, 2, enc);          // query[0]: the constant-1 wire (prover.rs:265)
    load_fq(&ctx->b1_0.x, pk->b_g1_query, 2, enc);
    load_fq(&ctx->b2_0.x.c0, pk->b_g2_query, 4, enc);
    G16_TRY(assemble_build_tables(ctx, ctx->main));
    auto range = [&](size_t total, size_t* lo, size_t* hi) {
        *lo = total * (size_t)shard_rank / (size_t)shard_count;
        *hi = total * (size_t)(shard_rank + 1) / (size_t)shard_count;
    };
    size_t lo, hi;
    size_t m1 = pk->a_len - 1;  // MSM over query[1..] (prover.rs:266)
    if (h_range && (h_range[0] > h_range[1] || h_range[1] > pk->h_len))
        return set_err(ctx, G16_ERR_BAD_ARG, "pk: h range [%llu, %llu) outside h_query (%zu points)", (unsigned long long)h_range[0],
                       (unsigned long long)h_range[1], pk->h_len);
    if (z_range && (z_range[0] > z_range[1] || z_range[1] > m1))
        return set_err(ctx, G16_ERR_BAD_ARG, "pk: wire range [%llu, %llu) outside a_query[1..] (%zu points)", (unsigned long long)z_range[0],
                       (unsigned long long)z_range[1], m1);
    if (h_range) lo = (size_t)h_range[0], hi = (size_t)h_range[1];
    else range(pk->h_len, &lo, &hi);
    G16_TRY(load_query(ctx, Q_H, 1, pk->h_query, pk->h_len, lo, hi, enc, precompute));
    if (z_range) lo = (size_t)z_range[0], hi = (size_t)z_range[1];
    else range(m1, &lo, &hi);
    G16_TRY(load_query(ctx, Q_A, 1, pk->a_query + 8, m1, lo, hi, enc, precompute));
    G16_TRY(load_query(ctx, Q_B1, 1, pk->b_g1_query + 8, m1, lo, hi, enc, precompute));
    G16_TRY(load_query(ctx, Q_B2, 2, pk->b_g2_query + 16, m1, lo, hi, enc, precompute));
    // b_g1 and b_g2 run over the same scalars z[1..] and are at infinity for the same wires (those absent from B): one
    // digit stage (scalar recoding + bucket sort) serves both.
    ctx->share_b = false;
    if (ctx->opt_share_digits) G16_TRY(msm_can_share(ctx, &ctx->q[Q_B1], &ctx->q[Q_B2], (hi - lo) / 64, &ctx->share_b, ctx->main));
    // l runs over aux = z[num_instance..], a over z[1..]: laying l_query out on a's index space (num_instance - 1 leading
    // points at infinity) lets l reuse a's digit stage.  Only worth it when a_query has (almost) no infinity points of its
    // own, because a's skipped wires would otherwise be idle lanes in l's additions -- decided here, per key.
    ctx->share_al = false;
    // l's shard always follows a's (the ranks' l ranges must partition l_query whatever each rank decides about sharing)
    size_t l_lo, l_hi;
    const bool aligned = pk->l_len <= m1;
    ctx->l_on_a_space = aligned;
    const size_t shift = aligned ? m1 - pk->l_len : 0;  // = num_instance - 1
    if (aligned) {
        l_lo = (lo > shift ? lo : shift) - shift;
        l_hi = (hi > shift ? hi : shift) - shift;
    } else if (z_range) {  // l_query longer than a_query[1..] never happens for a Groth16 key; keep the ranges a partition anyway
        l_lo = m1 ? pk->l_len * lo / m1 : 0;
        l_hi = m1 ? pk->l_len * hi / m1 : 0;
    } else {
        range(pk->l_len, &l_lo, &l_hi);
    }
    if (ctx->opt_share_digits && aligned && hi > lo) {
        size_t cnt = hi - lo;
        void* stage = nullptr;
        G16_CUDA(ctx, cudaMalloc(&stage, cnt * 64));
        int rc = G16_OK;
        cudaError_t e = cudaMemsetAsync(stage, 0, cnt * 64, ctx->main);
        if (e == cudaSuccess && l_hi > l_lo)
            e = cudaMemcpyAsync((char*)stage + (l_lo + shift - lo) * 64, (const char*)pk->l_query + l_lo * 64, (l_hi - l_lo) * 64,
                                cudaMemcpyHostToDevice, ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "pk upload: %s", cudaGetErrorString(e));
        if (rc == G16_OK && enc == G16_ENC_CANONICAL) rc = convert_mont_dev(ctx, G16_FIELD_FQ, stage, cnt * 2, true, ctx->main);
        if (rc == G16_OK)
            rc = msm_set_bases(ctx, &ctx->q[Q_L], &ctx->scratch[Q_L], 1, stage, cnt, ctx->opt_window_bits, precompute != 0, ctx->main);
        if (rc == G16_OK) rc = msm_can_share(ctx, &ctx->q[Q_A], &ctx->q[Q_L], shift + cnt / 64, &ctx->share_al, ctx->main);
        cudaStreamSynchronize(ctx->main);
        cudaFree(stage);
        G16_TRY(rc);
        ctx->pk_len[Q_L] = pk->l_len;
        ctx->sh_lo[Q_L] = l_lo;
        ctx->sh_hi[Q_L] = l_hi;
    }
    if (!ctx->share_al) G16_TRY(load_query(ctx, Q_L, 1, pk->l_query, pk->l_len, l_lo, l_hi, enc, precompute));
    ctx->have_pk = true;
    return G16_OK;
}

int g16_ctx_load_pk(g16_ctx* ctx, const g16_pk_view* pk, int shard_rank, int shard_count, int precompute) {
    if (!ctx || !pk) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return load_pk_ranges(ctx, pk, shard_rank, shard_count, nullptr, nullptr, precompute);
}
int g16_ctx_load_pk_ranges(g16_ctx* ctx, const g16_pk_view* pk, int shard_rank, int shard_count, const uint64_t h_range[2],
                           const uint64_t z_range[2], int precompute) {
    if (!ctx || !pk || !h_range || !z_range) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "load_pk_ranges: null argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    return load_pk_ranges(ctx, pk, shard_rank, shard_count, h_range, z_range, precompute);
}

// ---- prove ------------------------------------------------------------------------------------------------------------------------
static int check_ready(g16_ctx* ctx) {
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "prove: no R1CS loaded");
    if (!ctx->have_pk) return set_err(ctx, G16_ERR_BAD_ARG, "prove: no proving key loaded");
    size_t n = (size_t)1 << ctx->log_n;
    if (ctx->pk_len[Q_A] + 1 != ctx->m)
        return set_err(ctx, G16_ERR_BAD_ARG, "prove: a_query has %zu points for %llu wires", ctx->pk_len[Q_A] + 1, (unsigned long long)ctx->m);
    if (ctx->pk_len[Q_L] != ctx->m - ctx->ni)
        return set_err(ctx, G16_ERR_BAD_ARG, "prove: l_query has %zu points for %llu witness wires", ctx->pk_len[Q_L],
                       (unsigned long long)(ctx->m - ctx->ni));
    if (ctx->pk_len[Q_H] > n) return set_err(ctx, G16_ERR_BAD_ARG, "prove: h_query longer than the domain");
    return G16_OK;
}

}  // extern "C"

struct PartialLayout {
    G1XYZZ h, l, a, sa, rb1;
    G2XYZZ b2;
};

// ---- CUDA-graph replay of the launch sequences --------------------------------------------------------------------------------
// A proof is ~240 kernel launches on nine streams; everything data dependent is read from device memory by the kernels (sizes,
// offsets, scalars), so the sequence itself is identical from proof to proof.  The first run of a sequence is eager (lazy
// initialisations, loud errors), the second is captured, every later one is ONE cudaGraphLaunch: the host no longer feeds ~240
// launches + ~60 event operations per proof, which is what bounded small circuits and the small shards of an 8-GPU run
// (VERDICT r01 item 4).  Segments: FULL = a whole single-GPU proof; in the sharded path, where NCCL calls of the host glue sit
// between the pieces, WIRE (the z-only MSM chains, on their own root stream), WM (witness map) and H (the h MSM).
// Timing events inside a captured sequence are recorded as external event nodes, so the stage timings keep working.
enum { GR_FULL = 0, GR_WIRE = 1, GR_WM = 2, GR_H = 3, GR_PART = 4 /* .. 7: witness-map parts A, B, C, FINAL */ };

static int rec_t(g16_ctx* ctx, cudaEvent_t ev, cudaStream_t st) {
    if (ctx->capturing) G16_CUDA(ctx, cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal));
    else G16_CUDA(ctx, cudaEventRecord(ev, st));
    return G16_OK;
}

static void graphs_drop(g16_ctx* ctx) {
    for (auto& g : ctx->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = g16::GraphSlot();
    }
}

template <class Body>
static int run_graphed(g16_ctx* ctx, int kind, uint64_t key, cudaStream_t origin, Body body) {
    const bool want = ctx->opt_graph && !ctx->capturing && !ctx->opt_serialize && !ctx->opt_kernel_events && !ctx->opt_wm_priority &&
                      !getenv("G16_DEBUG_MSM");
    if (!want) return body();
    g16::GraphSlot& g = ctx->graphs[kind];
    key ^= ctx->graph_epoch * 0x9E3779B97F4A7C15ull;
    if (g.exec && g.key == key) {
        G16_CUDA(ctx, cudaGraphLaunch(g.exec, origin));
        ctx->launches += g.launches;
        ctx->graph_replays++;
        return G16_OK;
    }
    if (g.key != key || g.exec) {  // another sequence (other reduction / h source / options): start over
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = g16::GraphSlot();
        g.key = key;
    }
    if (g.seen++ == 0) return body();  // first run eager
    const uint64_t l0 = ctx->launches;
    G16_CUDA(ctx, cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    int rc = body();
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(origin, &graph);
    if (rc == G16_OK && e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != G16_OK || e != cudaSuccess) {
        // nothing has run yet (a capture only records): give up on graphs for this context and queue the sequence eagerly -- a
        // genuine error (not one caused by the capture itself) shows up again there and is reported
        cudaGetLastError();
        g = g16::GraphSlot();
        ctx->launches = l0;
        ctx->opt_graph = 0;
        ctx->graph_fallbacks++;
        return body();
    }
    g.launches = ctx->launches - l0;
    ctx->graph_captures++;
    ctx->graph_replays++;
    G16_CUDA(ctx, cudaGraphLaunch(g.exec, origin));
    return G16_OK;
}

// The z-only MSMs: l (aux = z[ni..]), a / b_g1 / b_g2 (assignment = z[1..])   (prover.rs:70-74,89-117), forked from `root`
// onto the side streams and joined back into `root`.  MSMs that share a digit stage are one chain.  Leaves l, a, s*a, r*b_g1,
// b_g2 in ctx->d_partial.  The scalars (r, s) must be on the device (assemble_set_scalars).
static int queue_wire_chains(g16_ctx* ctx, cudaStream_t root) {
    PartialLayout* part = (PartialLayout*)ctx->d_partial;
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_wfork, root));
    struct Job {
        int qi, from;
    };
    Job chains[4][2];
    int chain_len[4] = {0, 0, 0, 0};
    int nchains = 0;
    if (ctx->share_al) {
        chains[nchains][0] = {Q_A, -1};
        chains[nchains][1] = {Q_L, Q_A};
        chain_len[nchains++] = 2;
    } else {
        chains[nchains][0] = {Q_L, -1};
        chain_len[nchains++] = 1;
        chains[nchains][0] = {Q_A, -1};
        chain_len[nchains++] = 1;
    }
    if (ctx->share_b) {
        chains[nchains][0] = {Q_B1, -1};
        chains[nchains][1] = {Q_B2, Q_B1};
        chain_len[nchains++] = 2;
    } else {
        chains[nchains][0] = {Q_B1, -1};
        chain_len[nchains++] = 1;
        chains[nchains][0] = {Q_B2, -1};
        chain_len[nchains++] = 1;
    }
    uint32_t join_mask = 0;
    int nsplit = 0;
    for (int k = 0; k < nchains; k++) {
        cudaStream_t st0 = ctx->opt_serialize ? root : ctx->side[k];
        if (!ctx->opt_serialize) G16_CUDA(ctx, cudaStreamWaitEvent(st0, ctx->ev_wfork, 0));
        // split chain: the MSM that reuses the digit stage starts on its own stream as soon as that stage exists, instead of
        // queueing behind the first MSM's point stage (a serial chain left the G2 MSM alone at the end of the proof)
        const bool split = chain_len[k] == 2 && ctx->opt_split_chains && !ctx->opt_serialize;
        const int split_slot = split ? nsplit++ : -1;
        for (int j = 0; j < chain_len[k]; j++) {
            int qi = chains[k][j].qi, from = chains[k][j].from;
            cudaStream_t st = st0;
            int join_idx = k;
            if (split && j == 1) {
                st = ctx->side[7 + split_slot];
                join_idx = 7 + split_slot;
                G16_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_dig[split_slot], 0));
            }
            cudaEvent_t ea0 = ctx->opt_kernel_events ? ctx->ev_acc[2 * qi] : nullptr;
            cudaEvent_t ea1 = ctx->opt_kernel_events ? ctx->ev_acc[2 * qi + 1] : nullptr;
            const Fr* sc;
            size_t cnt;
            if (from >= 0) {  // same scalars and index space as the MSM whose digit stage is reused
                sc = nullptr;
                cnt = ctx->sh_hi[from] - ctx->sh_lo[from];
            } else {
                sc = ctx->d_z + (qi == Q_L ? ctx->ni : 1) + ctx->sh_lo[qi];
                cnt = ctx->sh_hi[qi] - ctx->sh_lo[qi];
            }
            G16_TRY(rec_t(ctx, ctx->ev_t[2 + 2 * qi], st));
            if (split && j == 0 && cnt == 0) G16_CUDA(ctx, cudaEventRecord(ctx->ev_dig[split_slot], st));  // nothing to build
            G16_TRY(msm_run(ctx, &ctx->q[qi], &ctx->scratch[qi], sc, cnt, st, ea0, ea1, from >= 0 ? &ctx->scratch[from] : nullptr,
                            split && j == 0 ? ctx->ev_dig[split_slot] : nullptr));
            const bool have = cnt && ctx->scratch[qi].result;
            // s * MSM_a and r * MSM_b1 are ~1.6 ms single-lane chains: they get their own streams so that neither the next
            // MSM of this chain nor anything else waits for them; they finish beside the MSMs still in flight
            auto scale_aside = [&](int slot, int which, void* dst) -> int {
                if (!have) {
                    G16_CUDA(ctx, cudaMemsetAsync(dst, 0, sizeof(G1XYZZ), st));
                    return G16_OK;
                }
                cudaStream_t ss = ctx->opt_serialize ? st : ctx->side[5 + slot];
                if (!ctx->opt_serialize) {
                    G16_CUDA(ctx, cudaEventRecord(ctx->ev_scale[slot], st));
                    G16_CUDA(ctx, cudaStreamWaitEvent(ss, ctx->ev_scale[slot], 0));
                }
                G16_TRY(scale_point_dev(ctx, ctx->scratch[qi].result, which, dst, ss));
                if (!ctx->opt_serialize) {
                    G16_CUDA(ctx, cudaEventRecord(ctx->ev_join[5 + slot], ss));
                    join_mask |= 1u << (5 + slot);
                }
                return G16_OK;
            };
            if (qi == Q_B1) {
                G16_TRY(scale_aside(1, 0, &part->rb1));  // only r * MSM_b1 is ever needed (prover.rs:118)
            } else {
                void* dst = qi == Q_L ? (void*)&part->l : qi == Q_A ? (void*)&part->a : (void*)&part->b2;
                size_t bytes = qi == Q_B2 ? sizeof(G2XYZZ) : sizeof(G1XYZZ);
                if (have) G16_CUDA(ctx, cudaMemcpyAsync(dst, ctx->scratch[qi].result, bytes, cudaMemcpyDeviceToDevice, st));
                else G16_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, st));
                if (qi == Q_A) G16_TRY(scale_aside(0, 1, &part->sa));  // s * MSM_a for s * g_a (prover.rs:98)
            }
            G16_TRY(rec_t(ctx, ctx->ev_t[3 + 2 * qi], st));
            if (!ctx->opt_serialize && (j + 1 == chain_len[k] || split)) {
                G16_CUDA(ctx, cudaEventRecord(ctx->ev_join[join_idx], st));
                join_mask |= 1u << join_idx;
            }
        }
    }
    for (int k = 0; k < kSideStreams; k++)
        if (join_mask & (1u << k)) G16_CUDA(ctx, cudaStreamWaitEvent(root, ctx->ev_join[k], 0));
    return G16_OK;
}

// Runs witness map + the five (sharded) MSMs; leaves this rank's partial sums in ctx->d_partial.  Witness and scalars must be
// on the device (ordered on main).  Split in two so that a rank which does not run the witness map itself can receive h
// between the halves:
//   shard_begin  starts the z-only MSM chains on the wire root stream (forked from main) and (run_wm) the witness map on main;
//   shard_finish runs the h MSM on main over h_src (NULL = the witness map's own output) and joins the wire root.
static int shard_begin(g16_ctx* ctx, int reduction, bool run_wm, bool allow_hi = false) {
    cudaStream_t main = ctx->main;
    bool busy = false;
    for (int qi : {Q_L, Q_A, Q_B1, Q_B2}) busy = busy || ctx->sh_hi[qi] > ctx->sh_lo[qi];
    cudaStream_t wire = ctx->opt_serialize ? main : ctx->wire;
    // option wm_first: the witness map runs first and alone, the wire chains start when it is done, so that all five MSM
    // chains run (and end) together instead of the h MSM finishing alone behind the others
    // (measured: S-rs256 29.0 -> 28.2 ms; S-2^12 1.86 -> 2.16 ms and S-2^16 3.06 -> 3.43 ms, where the map is a fraction of a
    // millisecond and delaying the chains only adds latency -- hence automatic from n = 2^20 on: profiles/r02_ab_sched2_glv.log)
    // (letting only the POINT stages wait for the map, with the digit stages beside it, was measured too: sort and transforms
    // slow each other down, 28.3 -> 28.8-30.7 ms: profiles/r02_ab_gate.log)
    const bool want_first = ctx->opt_wm_first < 0 ? ctx->log_n >= 20 : ctx->opt_wm_first != 0;
    const bool wm_first = run_wm && want_first && !ctx->opt_serialize && busy;
    auto start_wire = [&]() -> int {
        if (wire != main) {
            G16_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main));
            G16_CUDA(ctx, cudaStreamWaitEvent(wire, ctx->ev_fork, 0));
        }
        G16_TRY(run_graphed(ctx, GR_WIRE, 0x57, wire, [&] { return queue_wire_chains(ctx, wire); }));
        if (wire != main) G16_CUDA(ctx, cudaEventRecord(ctx->ev_wire_done, wire));
        return G16_OK;
    };
    if (!wm_first) G16_TRY(start_wire());
    // main: witness map (r1cs_to_qap.rs:150-213)
    // (option wm_priority: on the high-priority twin of main, so that the h MSM -- which can only start afterwards -- is not
    // pushed to the end of the proof by the four z-only MSMs already in flight)
    cudaStream_t ws = (allow_hi && run_wm && ctx->opt_wm_priority && !ctx->opt_serialize && !wm_first && ctx->hi) ? ctx->hi : main;
    ctx->sh_wm_stream = ws;
    if (ws != main) {
        G16_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main));
        G16_CUDA(ctx, cudaStreamWaitEvent(ws, ctx->ev_fork, 0));
    }
    if (run_wm) {
        // transforms beside MSM chains use the radix-2 passes, a lone witness map the radix-4 ones (see opt_ntt_radix4)
        const bool alone = ctx->opt_serialize || !busy || wm_first;
        G16_TRY(run_graphed(ctx, GR_WM, 0x100 + (uint64_t)reduction * 2 + alone, ws, [&] {
            G16_TRY(rec_t(ctx, ctx->ev_t[0], ws));
            ctx->wm_alone = alone;
            int rc = witness_map_dev(ctx, reduction, ws);
            ctx->wm_alone = true;
            G16_TRY(rc);
            return rec_t(ctx, ctx->ev_t[1], ws);
        }));
    } else {
        G16_TRY(rec_t(ctx, ctx->ev_t[0], ws));
        G16_TRY(rec_t(ctx, ctx->ev_t[1], ws));
    }
    if (wm_first) G16_TRY(start_wire());
    return G16_OK;
}

static int shard_finish(g16_ctx* ctx, const Fr* h_src) {
    cudaStream_t main = ctx->main;
    // wm_priority = 1: the h MSM stays on the witness map's (high-priority) stream; = 2: only the witness map ran there
    cudaStream_t hs = (!h_src && ctx->sh_wm_stream && ctx->opt_wm_priority != 2) ? ctx->sh_wm_stream : main;
    if (hs == main && ctx->sh_wm_stream && ctx->sh_wm_stream != main) {
        G16_CUDA(ctx, cudaEventRecord(ctx->ev_hi, ctx->sh_wm_stream));
        G16_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_hi, 0));
    }
    PartialLayout* part = (PartialLayout*)ctx->d_partial;
    // the h MSM over h[lo..hi)  (prover.rs:63-66; the zip drops h[n-1], generator.rs:178)
    G16_TRY(run_graphed(ctx, GR_H, 0x200 ^ (uint64_t)(uintptr_t)h_src, hs, [&] {
        size_t cnt = ctx->sh_hi[Q_H] - ctx->sh_lo[Q_H];
        G16_TRY(rec_t(ctx, ctx->ev_t[2 + 2 * Q_H], hs));
        G16_TRY(msm_run(ctx, &ctx->q[Q_H], &ctx->scratch[Q_H], (h_src ? h_src : ctx->d_a) + ctx->sh_lo[Q_H], cnt, hs,
                        ctx->opt_kernel_events ? ctx->ev_acc[0] : nullptr, ctx->opt_kernel_events ? ctx->ev_acc[1] : nullptr));
        if (cnt && ctx->scratch[Q_H].result)
            G16_CUDA(ctx, cudaMemcpyAsync(&part->h, ctx->scratch[Q_H].result, sizeof(G1XYZZ), cudaMemcpyDeviceToDevice, hs));
        else
            G16_CUDA(ctx, cudaMemsetAsync(&part->h, 0, sizeof(G1XYZZ), hs));
        return rec_t(ctx, ctx->ev_t[3 + 2 * Q_H], hs);
    }));
    if (hs != main) {
        G16_CUDA(ctx, cudaEventRecord(ctx->ev_hi, hs));
        G16_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_hi, 0));
    }
    ctx->sh_wm_stream = nullptr;
    if (!ctx->opt_serialize) G16_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_wire_done, 0));
    return G16_OK;
}

static void collect_timings(g16_ctx* ctx, bool with_asm) {
    auto el = [&](int a, int b) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_t[a], ctx->ev_t[b]) != cudaSuccess) {
            cudaGetLastError();
            ms = -1;
        }
        return ms;
    };
    ctx->tm.witness_map_ms = el(0, 1);
    ctx->tm.msm_h_ms = el(2, 3);
    ctx->tm.msm_l_ms = el(4, 5);
    ctx->tm.msm_a_ms = el(6, 7);
    ctx->tm.msm_b_g1_ms = el(8, 9);
    ctx->tm.msm_b_g2_ms = el(10, 11);
    ctx->tm.h2d_ms = el(14, 0);
    ctx->tm.h_wait_ms = el(1, 2);
    ctx->tm.h_start_ms = el(14, 2);
    for (int k = 0; k < 5; k++) {
        float ms = -1;
        if (ctx->opt_kernel_events && cudaEventElapsedTime(&ms, ctx->ev_acc[2 * k], ctx->ev_acc[2 * k + 1]) != cudaSuccess) {
            cudaGetLastError();
            ms = -1;
        }
        ctx->tm.acc_ms[k] = ms;
    }
    if (with_asm) {
        ctx->tm.assemble_ms = el(12, 13);
        ctx->tm.total_ms = el(14, 13);
    } else {
        ctx->tm.assemble_ms = 0;
        ctx->tm.total_ms = el(14, 12);
    }
}

static int prove_full(g16_ctx* ctx, const uint64_t* z, const uint64_t* r, const uint64_t* s, int reduction, g16_proof* out) {
    if (!r || !s || !out) return set_err(ctx, G16_ERR_BAD_ARG, "prove: null r/s/out");
    G16_TRY(check_ready(ctx));
    if (ctx->shard_count != 1) return set_err(ctx, G16_ERR_BAD_ARG, "prove: context holds shard %d/%d; use prove_shard + prove_combine", ctx->shard_rank, ctx->shard_count);
    cudaStream_t main = ctx->main;
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[14], main));
    if (z) G16_TRY(upload_witness(ctx, z, main));
    else if (!ctx->witness_resident) return set_err(ctx, G16_ERR_BAD_ARG, "prove_resident: no witness uploaded");
    ctx->pre_pending = false;  // the AsmPre slot is about to hold THIS call's (r, s): a g16_prove_prepare before it is void
    G16_TRY(assemble_set_scalars(ctx, r, s, main));
    G16_TRY(run_graphed(ctx, GR_FULL, 0x300 + (uint64_t)reduction, main, [&] {
        // (r, s, pk)-only scalar multiplications overlap everything else
        G16_CUDA(ctx, cudaEventRecord(ctx->ev_pre, main));
        G16_CUDA(ctx, cudaStreamWaitEvent(ctx->side[4], ctx->ev_pre, 0));
        G16_TRY(assemble_pre(ctx, ctx->side[4]));
        G16_CUDA(ctx, cudaEventRecord(ctx->ev_join[4], ctx->side[4]));
        G16_TRY(shard_begin(ctx, reduction, true, true));
        G16_TRY(shard_finish(ctx, nullptr));
        G16_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_join[4], 0));
        G16_TRY(rec_t(ctx, ctx->ev_t[12], main));
        G16_TRY(assemble_proof_queue(ctx, ctx->d_partial, 1, main));
        return rec_t(ctx, ctx->ev_t[13], main);
    }));
    G16_CUDA(ctx, cudaStreamSynchronize(main));
    collect_timings(ctx, true);
    {
        float ms = -1;
        if (cudaEventElapsedTime(&ms, ctx->ev_t[12], ctx->ev_t[15]) != cudaSuccess) cudaGetLastError();
        ctx->tm.assemble_kernel_ms = ms;
    }
    assemble_proof_read(ctx, out);
    return G16_OK;
}

extern "C" {

int g16_prove(g16_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], int reduction, g16_proof* out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    if (!z) return set_err(ctx, G16_ERR_BAD_ARG, "prove: witness pointer is NULL");
    Guard g(ctx);
    return prove_full(ctx, z, r, s, reduction, out);
}
int g16_prove_resident(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction, g16_proof* out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    return prove_full(ctx, nullptr, r, s, reduction, out);
}

// the (r, s) the per-rank scalings read: already on the device when g16_prove_prepare queued them for the same pair
static int shard_scalars(g16_ctx* ctx, const uint64_t* r, const uint64_t* s) {
    if (ctx->pre_pending && !memcmp(ctx->pre_r, r, 32) && !memcmp(ctx->pre_s, s, 32)) return G16_OK;
    ctx->pre_pending = false;
    return assemble_set_scalars(ctx, r, s, ctx->main);
}

int g16_prove_shard_dev(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction) {
    if (!ctx) return G16_ERR_BAD_ARG;
    if (!r || !s) return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard: null r/s");
    Guard g(ctx);
    G16_TRY(check_ready(ctx));
    if (!ctx->witness_resident) return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard_dev: no witness uploaded");
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[14], ctx->main));
    G16_TRY(shard_scalars(ctx, r, s));
    G16_TRY(shard_begin(ctx, reduction, true, true));
    G16_TRY(shard_finish(ctx, nullptr));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[12], ctx->main));
    ctx->tm_stale = true;
    return G16_OK;
}

int g16_prove_shard_begin_dev(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction, int run_witness_map) {
    if (!ctx) return G16_ERR_BAD_ARG;
    if (!r || !s) return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard_begin: null r/s");
    Guard g(ctx);
    G16_TRY(check_ready(ctx));
    if (!(ctx->witness_resident || (ctx->witness_partial && !run_witness_map)))
        return set_err(ctx, G16_ERR_BAD_ARG, run_witness_map ? "prove_shard_begin_dev: the witness map needs the whole witness on the device"
                                                              : "prove_shard_begin_dev: no witness uploaded");
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[14], ctx->main));
    G16_TRY(shard_scalars(ctx, r, s));
    G16_TRY(shard_begin(ctx, reduction, run_witness_map != 0));
    ctx->shard_open = true;
    return G16_OK;
}

int g16_prove_shard_finish_dev(g16_ctx* ctx, const void* h_dev, size_t h_first, size_t h_count) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->shard_open) return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard_finish_dev without prove_shard_begin_dev");
    ctx->shard_open = false;
    if (h_dev && (h_first > ctx->sh_lo[Q_H] || h_first + h_count < ctx->sh_hi[Q_H])) {
        // the wire chains are already in flight: join them so that the context is reusable, then report
        if (!ctx->opt_serialize) cudaStreamWaitEvent(ctx->main, ctx->ev_wire_done, 0);
        return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard_finish: h[%zu, %zu) does not cover this rank's range [%zu, %zu)", h_first,
                       h_first + h_count, ctx->sh_lo[Q_H], ctx->sh_hi[Q_H]);
    }
    // shard_finish indexes its source by the absolute coefficient index
    G16_TRY(shard_finish(ctx, h_dev ? (const Fr*)h_dev - h_first : nullptr));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[12], ctx->main));
    ctx->tm_stale = true;
    return G16_OK;
}

int g16_witness_map_part_dev(g16_ctx* ctx, int parts) {
    if (!ctx) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map_part: no R1CS loaded");
    if (!ctx->witness_resident) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map_part: the whole witness must be on the device");
    if (parts <= 0 || parts > 15) return set_err(ctx, G16_ERR_BAD_ARG, "witness_map_part: parts mask %d", parts);
    for (int k = 0; k < 4; k++) {
        const int bit = 1 << k;
        if (!(parts & bit)) continue;
        const bool alone = ctx->wm_alone;
        bool busy = false;  // wire MSMs beside the transforms? (then the modest radix-2, unbatched launches: see opt_ntt_radix4)
        if (ctx->have_pk)
            for (int qi : {Q_L, Q_A, Q_B1, Q_B2}) busy = busy || ctx->sh_hi[qi] > ctx->sh_lo[qi];
        G16_TRY(run_graphed(ctx, GR_PART + k, 0x400 + (uint64_t)bit * 2 + busy, ctx->main, [&] {
            if (bit == G16_WM_PART_A) G16_TRY(rec_t(ctx, ctx->ev_t[0], ctx->main));
            ctx->wm_alone = !busy;
            int rc = witness_map_part_dev(ctx, bit, ctx->main);
            ctx->wm_alone = alone;
            G16_TRY(rc);
            if (bit == G16_WM_PART_FINAL) G16_TRY(rec_t(ctx, ctx->ev_t[1], ctx->main));
            return G16_OK;
        }));
    }
    return G16_OK;
}

int g16_wm_vector_copy_dev(g16_ctx* ctx, int which, void* ext_dev, size_t capacity_elems, int to_ctx) {
    if (!ctx || !ext_dev || which < 0 || which > 2) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "wm_vector_copy: bad argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "wm_vector_copy: no R1CS loaded");
    const size_t n = (size_t)1 << ctx->log_n;
    if (capacity_elems < n) return set_err(ctx, G16_ERR_BAD_ARG, "wm_vector_copy: buffer holds %zu < %zu elements", capacity_elems, n);
    Fr* v = which == 0 ? ctx->d_a : which == 1 ? ctx->d_b : ctx->d_c;
    if (to_ctx) G16_CUDA(ctx, cudaMemcpyAsync(v, ext_dev, n * 32, cudaMemcpyDeviceToDevice, ctx->main));
    else G16_CUDA(ctx, cudaMemcpyAsync(ext_dev, v, n * 32, cudaMemcpyDeviceToDevice, ctx->main));
    return G16_OK;
}

int g16_copy_h_dev(g16_ctx* ctx, void* dst_dev, size_t capacity_elems) {
    if (!ctx || !dst_dev) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_r1cs) return set_err(ctx, G16_ERR_BAD_ARG, "copy_h: no R1CS loaded");
    size_t n = (size_t)1 << ctx->log_n;
    if (capacity_elems < n) return set_err(ctx, G16_ERR_BAD_ARG, "copy_h: buffer holds %zu < %zu elements", capacity_elems, n);
    G16_CUDA(ctx, cudaMemcpyAsync(dst_dev, ctx->d_a, n * 32, cudaMemcpyDeviceToDevice, ctx->main));
    return G16_OK;
}

int g16_prove_shard(g16_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], int reduction, g16_partial* out) {
    if (!ctx || !out) return G16_ERR_BAD_ARG;
    if (!r || !s) return set_err(ctx, G16_ERR_BAD_ARG, "prove_shard: null r/s");
    Guard g(ctx);
    G16_TRY(check_ready(ctx));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[14], ctx->main));
    G16_TRY(upload_witness(ctx, z, ctx->main));
    G16_TRY(shard_scalars(ctx, r, s));
    G16_TRY(shard_begin(ctx, reduction, true, true));
    G16_TRY(shard_finish(ctx, nullptr));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[12], ctx->main));
    G16_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_partial, sizeof(g16_partial), cudaMemcpyDeviceToHost, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    collect_timings(ctx, false);
    return G16_OK;
}

int g16_partial_dev(g16_ctx* ctx, void** dev_ptr, size_t* bytes) {
    if (!ctx || !dev_ptr || !bytes) return G16_ERR_BAD_ARG;
    *dev_ptr = ctx->d_partial;
    *bytes = sizeof(g16_partial);
    return G16_OK;
}

int g16_copy_partial_dev(g16_ctx* ctx, void* dst_dev) {
    if (!ctx || !dst_dev) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    G16_CUDA(ctx, cudaMemcpyAsync(dst_dev, ctx->d_partial, sizeof(g16_partial), cudaMemcpyDeviceToDevice, ctx->main));
    return G16_OK;
}

int g16_prove_combine_dev(g16_ctx* ctx, const void* dev_partials, int count, const uint64_t r[4], const uint64_t s[4],
                          g16_proof* out) {
    if (!ctx || !dev_partials || !r || !s || !out || count < 1) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "prove_combine: bad argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_pk) return set_err(ctx, G16_ERR_BAD_ARG, "prove_combine: no proving key loaded");
    if (ctx->pre_pending && !memcmp(ctx->pre_r, r, 32) && !memcmp(ctx->pre_s, s, 32)) {
        // g16_prove_prepare already ran the (r, s)-only scalar multiplications on a side stream: just join it
        G16_CUDA(ctx, cudaStreamWaitEvent(ctx->main, ctx->ev_join[4], 0));
    } else {
        ctx->pre_pending = false;
        G16_TRY(assemble_set_scalars(ctx, r, s, ctx->main));
        G16_TRY(assemble_pre(ctx, ctx->main));
    }
    ctx->pre_pending = false;
    G16_TRY(assemble_proof_queue(ctx, dev_partials, count, ctx->main));
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    assemble_proof_read(ctx, out);
    return G16_OK;
}

int g16_prove_prepare(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4]) {
    if (!ctx || !r || !s) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_pk) return set_err(ctx, G16_ERR_BAD_ARG, "prove_prepare: no proving key loaded");
    G16_TRY(assemble_set_scalars(ctx, r, s, ctx->main));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_pre, ctx->main));
    G16_CUDA(ctx, cudaStreamWaitEvent(ctx->side[4], ctx->ev_pre, 0));
    G16_TRY(assemble_pre(ctx, ctx->side[4]));
    G16_CUDA(ctx, cudaEventRecord(ctx->ev_join[4], ctx->side[4]));
    memcpy(ctx->pre_r, r, 32);
    memcpy(ctx->pre_s, s, 32);
    ctx->pre_pending = true;
    return G16_OK;
}

int g16_prove_combine(g16_ctx* ctx, const g16_partial* partials, int count, const uint64_t r[4], const uint64_t s[4],
                      g16_proof* out) {
    if (!ctx || !partials || count < 1) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "prove_combine: bad argument") : G16_ERR_BAD_ARG;
    void* d = nullptr;
    {
        Guard g(ctx);
        G16_CUDA(ctx, cudaMalloc(&d, sizeof(g16_partial) * (size_t)count));
        G16_CUDA(ctx, cudaMemcpy(d, partials, sizeof(g16_partial) * (size_t)count, cudaMemcpyHostToDevice));
    }
    int rc = g16_prove_combine_dev(ctx, d, count, r, s, out);
    cudaFree(d);
    return rc;
}

int g16_set_option(g16_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    ctx->graph_epoch++;  // captured launch sequences bake the options in
    if (!strcmp(key, "serialize")) ctx->opt_serialize = value;
    else if (!strcmp(key, "graph")) ctx->opt_graph = value;
    else if (!strcmp(key, "kernel_events")) ctx->opt_kernel_events = value;
    else if (!strcmp(key, "window_bits")) ctx->opt_window_bits = value;
    else if (!strcmp(key, "acc_variant")) ctx->opt_acc_variant = value;
    else if (!strcmp(key, "ba_levels")) ctx->opt_ba_levels = value;
    else if (!strcmp(key, "share_digits")) ctx->opt_share_digits = value;
    else if (!strcmp(key, "split_chains")) ctx->opt_split_chains = value;
    else if (!strcmp(key, "ntt_radix4")) ctx->opt_ntt_radix4 = value;
    else if (!strcmp(key, "spmv_sell")) ctx->opt_spmv_sell = value;
    else if (!strcmp(key, "ntt_batch")) ctx->opt_ntt_batch = value;
    else if (!strcmp(key, "wm_priority")) ctx->opt_wm_priority = value;
    else if (!strcmp(key, "wm_first")) ctx->opt_wm_first = value;
    else if (!strcmp(key, "verify_occupancy")) ctx->opt_verify_occupancy = value;
    else if (!strcmp(key, "asm_tables")) ctx->opt_asm_tables = value;
    else return set_err(ctx, G16_ERR_BAD_ARG, "unknown option '%s'", key);
    return G16_OK;
}

int g16_pow_table(g16_ctx* ctx, const uint64_t base[4], const uint64_t scale[4], size_t n, uint64_t* out) {
    if (!ctx || !base || !scale || (n && !out)) return G16_ERR_BAD_ARG;
    if (n == 0) return G16_OK;
    if (n >= ((size_t)1 << 32)) return set_err(ctx, G16_ERR_BAD_ARG, "pow_table: n too large");
    Guard g(ctx);
    Fr b, s;
    memcpy(&b, base, 32);
    memcpy(&s, scale, 32);
    Fr* d;
    G16_CUDA(ctx, cudaMalloc((void**)&d, n * 32));
    int rc = pow_table_dev(ctx, d, n, b, s, ctx->main);
    if (rc == G16_OK) {
        cudaMemcpyAsync(out, d, n * 32, cudaMemcpyDeviceToHost, ctx->main);
        cudaError_t e = cudaStreamSynchronize(ctx->main);
        if (e != cudaSuccess) rc = set_err(ctx, G16_ERR_CUDA, "pow_table: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

int g16_get_msm_stats(g16_ctx* ctx, int which, uint64_t* out, int capacity) {
    if (!ctx || !out || which < 0 || which > 4 || capacity < 16) return ctx ? set_err(ctx, G16_ERR_BAD_ARG, "msm_stats: bad argument") : G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (!ctx->have_pk) return set_err(ctx, G16_ERR_BAD_ARG, "msm_stats: no proving key loaded");
    int owner = which;
    if (which == Q_L && ctx->share_al) owner = Q_A;
    if (which == Q_B2 && ctx->share_b) owner = Q_B1;
    const MsmBases& mb = ctx->q[which];
    const MsmScratch& dg = ctx->scratch[owner];
    G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    memset(out, 0, sizeof(uint64_t) * (size_t)capacity);
    out[0] = mb.n;
    out[1] = (uint64_t)mb.c;
    out[2] = (uint64_t)mb.windows;
    out[3] = (uint64_t)mb.ba_levels;
    out[4] = dg.cap_buckets;
    out[5] = owner != which;  // digit stage shared with another MSM
    if (mb.n == 0) return G16_OK;
    uint32_t v = 0;
    for (int k = 1; k <= mb.ba_levels && 7 + k < capacity; k++) {
        G16_CUDA(ctx, cudaMemcpy(&v, dg.ba_lvl + (size_t)k * dg.ba_stride + dg.cap_buckets, 4, cudaMemcpyDeviceToHost));
        out[7 + k] = v;  // points left after level k == output slots of level k
    }
    G16_CUDA(ctx, cudaMemcpy(&v, dg.counters, 4, cudaMemcpyDeviceToHost));
    out[6] = v;  // XYZZ tasks of the tail
    return G16_OK;
}

int g16_get_timings(g16_ctx* ctx, g16_timings* out) {
    if (!ctx || !out) return G16_ERR_BAD_ARG;
    Guard g(ctx);
    if (ctx->tm_stale) {  // the *_dev shard entry points do not synchronise: read their events now
        G16_CUDA(ctx, cudaStreamSynchronize(ctx->main));
        collect_timings(ctx, false);
        ctx->tm_stale = false;
    }
    *out = ctx->tm;
    return G16_OK;
}

}  // extern "C"
