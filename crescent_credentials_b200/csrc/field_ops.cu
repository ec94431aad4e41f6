// field_ops.cu -- element-wise field kernels (K1 parity hook), Montgomery conversion, integer-pipe probes.
#include "internal.cuh"

namespace g16 {

template <class F>
__global__ void k_field_op(int op, const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        F x = a[i];
        F r;
        switch (op) {
            case G16_OP_MUL: r = x * b[i]; break;
            case G16_OP_ADD: r = x + b[i]; break;
            case G16_OP_SUB: r = x - b[i]; break;
            case G16_OP_NEG: r = x.neg(); break;
            case G16_OP_INV: r = x.inverse(); break;
            case G16_OP_SQR: r = x.sqr(); break;
            case G16_OP_MUL_BCAST: r = x * b[0]; break;
            case G16_OP_ADD_BCAST: r = x + b[0]; break;
            default: r = x; break;
        }
        out[i] = r;
    }
}

template <class F>
__global__ void k_mont_conv(F* __restrict__ data, size_t n, bool to_mont) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        F x = data[i];
        data[i] = to_mont ? x.to_mont() : x.from_mont();
    }
}

static inline int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    size_t cap = (size_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int convert_mont_dev(g16_ctx* ctx, int field, void* data, size_t n, bool to_mont, cudaStream_t st) {
    if (n == 0) return G16_OK;
    if (field == G16_FIELD_FR)
        G16_LAUNCH(ctx, k_mont_conv<Fr>, grid_for(n, 256), 256, 0, st, (Fr*)data, n, to_mont);
    else
        G16_LAUNCH(ctx, k_mont_conv<Fq>, grid_for(n, 256), 256, 0, st, (Fq*)data, n, to_mont);
    return G16_OK;
}

int field_op_dev(g16_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return G16_OK;
    if (op == G16_OP_TO_MONT || op == G16_OP_FROM_MONT) {
        if (field == G16_FIELD_FQ2) {
            field = G16_FIELD_FQ;
            n *= 2;
        }
        if (out != a) G16_CUDA(ctx, cudaMemcpyAsync(out, a, n * 32, cudaMemcpyDeviceToDevice, st));
        return convert_mont_dev(ctx, field, out, n, op == G16_OP_TO_MONT, st);
    }
    int g = grid_for(n, 128);
    switch (field) {
        case G16_FIELD_FR:
            G16_LAUNCH(ctx, k_field_op<Fr>, g, 128, 0, st, op, (const Fr*)a, (const Fr*)b, (Fr*)out, n);
            break;
        case G16_FIELD_FQ:
            G16_LAUNCH(ctx, k_field_op<Fq>, g, 128, 0, st, op, (const Fq*)a, (const Fq*)b, (Fq*)out, n);
            break;
        case G16_FIELD_FQ2:
            G16_LAUNCH(ctx, k_field_op<Fq2>, g, 128, 0, st, op, (const Fq2*)a, (const Fq2*)b, (Fq2*)out, n);
            break;
        default:
            return set_err(ctx, G16_ERR_BAD_ARG, "unknown field id %d", field);
    }
    return G16_OK;
}

// ---- integer-pipe probes ----------------------------------------------------------------------------------------------
// which = 0: independent 32-bit IMAD chains; 1: IMAD.WIDE.U32 chains; 2/3: dependent Fr/Fq Montgomery products with
// 4 independent streams per thread.  All are pure register kernels sized to fill every SM.
__global__ void k_probe_imad(uint32_t* out, int iters) {
    uint32_t a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11, a6 = a0 * 13,
             a7 = a0 * 17;
    uint32_t m = blockIdx.x * 2 + 1;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = a0 * m + a1;
            a1 = a1 * m + a2;
            a2 = a2 * m + a3;
            a3 = a3 * m + a4;
            a4 = a4 * m + a5;
            a5 = a5 * m + a6;
            a6 = a6 * m + a7;
            a7 = a7 * m + a0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

__global__ void k_probe_imad_wide(uint64_t* out, int iters) {
    uint64_t a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11, a6 = a0 * 13,
             a7 = a0 * 17;
    uint32_t m = blockIdx.x * 2 + 1;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            // 32 x 32 + 64 -> 64: one IMAD.WIDE.U32 each
            a0 = (uint64_t)(uint32_t)a1 * m + a0;
            a1 = (uint64_t)(uint32_t)a2 * m + a1;
            a2 = (uint64_t)(uint32_t)a3 * m + a2;
            a3 = (uint64_t)(uint32_t)a4 * m + a3;
            a4 = (uint64_t)(uint32_t)a5 * m + a4;
            a5 = (uint64_t)(uint32_t)a6 * m + a5;
            a6 = (uint64_t)(uint32_t)a7 * m + a6;
            a7 = (uint64_t)(uint32_t)a0 * m + a7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

template <class F>
__global__ void k_probe_mul(F* out, int iters) {
    F a = F::one(), b = F::r2(), c = F::one() + F::one(), d = b + a;
    a.v[0] ^= threadIdx.x;
    b.v[1] ^= blockIdx.x;
    a = a + F::zero();
    for (int i = 0; i < iters; i++) {
        a = a * b;
        b = b * c;
        c = c * d;
        d = d * a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}

int bench_int_pipe(g16_ctx* ctx, int which, double* gops) {
    const int block = 256;
    const int grid = kNumSMs * 8;
    void* buf = nullptr;
    G16_CUDA(ctx, cudaMalloc(&buf, (size_t)grid * block * 32));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = (which <= 1) ? 2000 : 400;
    double ops_per_thread = (which <= 1) ? (double)iters * 16 * 8 : (double)iters * 4;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, ctx->main);
        switch (which) {
            case 0: k_probe_imad<<<grid, block, 0, ctx->main>>>((uint32_t*)buf, iters); break;
            case 1: k_probe_imad_wide<<<grid, block, 0, ctx->main>>>((uint64_t*)buf, iters); break;
            case 2: k_probe_mul<Fr><<<grid, block, 0, ctx->main>>>((Fr*)buf, iters); break;
            default: k_probe_mul<Fq><<<grid, block, 0, ctx->main>>>((Fq*)buf, iters); break;
        }
        ctx->launches++;
        cudaEventRecord(e1, ctx->main);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(buf);
            return set_err(ctx, G16_ERR_CUDA, "probe kernel failed: %s", cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *gops = ops_per_thread * grid * block / (best * 1e-3) / 1e9;
    return G16_OK;
}

}  // namespace g16
