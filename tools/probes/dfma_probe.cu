// dfma_probe.cu -- is the FP64 pipe a second multiplier for Fq?  Stand-alone probe (not part of libg16b200.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o dfma_probe dfma_probe.cu
// Prints: DFMA peak, IMAD.WIDE peak, both co-issued from one thread, and dependent Fq Montgomery products per second for
// the integer product (fp.cuh), the FP64 product (fp_dfma.cuh) and warp-interleaved mixes of the two; every variant's
// result is compared with the integer product's.
#include <cstdio>
#include <cstdlib>
#include "fp_dfma.cuh"
using namespace g16;

__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11, a6 = a0 * 13, a7 = a0 * 17;
    double m = 1.0 + 1e-9 * blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = fma(a0, m, a1); a1 = fma(a1, m, a2); a2 = fma(a2, m, a3); a3 = fma(a3, m, a4);
            a4 = fma(a4, m, a5); a5 = fma(a5, m, a6); a6 = fma(a6, m, a7); a7 = fma(a7, m, a0);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_wide(uint64_t* out, int iters) {
    uint64_t a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11, a6 = a0 * 13, a7 = a0 * 17;
    uint32_t m = blockIdx.x * 2 + 1;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = (uint64_t)(uint32_t)a1 * m + a0; a1 = (uint64_t)(uint32_t)a2 * m + a1; a2 = (uint64_t)(uint32_t)a3 * m + a2;
            a3 = (uint64_t)(uint32_t)a4 * m + a3; a4 = (uint64_t)(uint32_t)a5 * m + a4; a5 = (uint64_t)(uint32_t)a6 * m + a5;
            a6 = (uint64_t)(uint32_t)a7 * m + a6; a7 = (uint64_t)(uint32_t)a0 * m + a7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
// RATIO DFMA per IMAD.WIDE in one instruction stream
template <int ND, int NW>
__global__ void k_mix(uint64_t* out, int iters) {
    uint64_t a[8];
    double d[8];
    for (int k = 0; k < 8; k++) { a[k] = threadIdx.x + 1 + k; d[k] = threadIdx.x * 0.5 + k; }
    uint32_t m = blockIdx.x * 2 + 1;
    double md = 1.0 + 1e-9 * blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int x = 0; x < NW; x++) a[(k + x) & 7] = (uint64_t)(uint32_t)a[(k + x + 1) & 7] * m + a[(k + x) & 7];
#pragma unroll
            for (int x = 0; x < ND; x++) d[(k + x) & 7] = fma(d[(k + x) & 7], md, d[(k + x + 1) & 7]);
        }
    }
    uint64_t r = 0;
    for (int k = 0; k < 8; k++) r ^= a[k] ^ (uint64_t)__double_as_longlong(d[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// mode 0: integer product; 1: FP64 product; 2..: warp w uses the FP64 product iff (w % den) < num
template <class F, class PR>
__global__ void __launch_bounds__(256) k_mul(F* out, int iters, int num, int den) {
    F a = F::one(), b = F::r2(), c = F::one() + F::one(), d = b + a;
    a.v[0] ^= threadIdx.x;
    b.v[1] ^= blockIdx.x;
    a = a + F::zero();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if ((warp % den) < num) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            a = mul_dfma<PR>(a, b); b = mul_dfma<PR>(b, c); c = mul_dfma<PR>(c, d); d = mul_dfma<PR>(d, a);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            a = a * b; b = b * c; c = c * d; d = d * a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <class L>
static float time_it(L launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int block = 256, grid = sms * 8;
    void* buf; CK(cudaMalloc(&buf, (size_t)grid * block * 32));
    void* ref; CK(cudaMalloc(&ref, (size_t)grid * block * 32));
    const size_t threads = (size_t)grid * block;
    {
        int iters = 2000;
        float ms = time_it([&] { k_dfma<<<grid, block>>>((double*)buf, iters); });
        printf("{\"probe\": \"dfma\", \"gops\": %.1f}\n", (double)iters * 128 * threads / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_wide<<<grid, block>>>((uint64_t*)buf, iters); });
        printf("{\"probe\": \"imad_wide\", \"gops\": %.1f}\n", (double)iters * 128 * threads / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_mix<1, 1><<<grid, block>>>((uint64_t*)buf, iters); });
        printf("{\"probe\": \"mix 1 dfma : 1 wide\", \"dfma_gops\": %.1f, \"wide_gops\": %.1f}\n", (double)iters * 8 * threads / (ms * 1e-3) / 1e9, (double)iters * 8 * threads / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_mix<2, 1><<<grid, block>>>((uint64_t*)buf, iters); });
        printf("{\"probe\": \"mix 2 dfma : 1 wide\", \"dfma_gops\": %.1f, \"wide_gops\": %.1f}\n", (double)iters * 16 * threads / (ms * 1e-3) / 1e9, (double)iters * 8 * threads / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_mix<3, 1><<<grid, block>>>((uint64_t*)buf, iters); });
        printf("{\"probe\": \"mix 3 dfma : 1 wide\", \"dfma_gops\": %.1f, \"wide_gops\": %.1f}\n", (double)iters * 24 * threads / (ms * 1e-3) / 1e9, (double)iters * 8 * threads / (ms * 1e-3) / 1e9);
    }
    {
        int iters = 400;
        const int mixes[][2] = {{0, 1}, {1, 1}, {1, 2}, {2, 3}, {3, 4}, {1, 3}, {3, 5}};
        for (auto& mx : mixes) {
            void* dst = (mx[0] == 0) ? ref : buf;
            float ms = time_it([&] { k_mul<Fq, FqParams><<<grid, block>>>((Fq*)dst, iters, mx[0], mx[1]); });
            int same = -1;
            if (mx[0] != 0) {
                static uint32_t *h0 = nullptr, *h1 = nullptr;
                size_t bytes = threads * 32;
                if (!h0) { h0 = (uint32_t*)malloc(bytes); h1 = (uint32_t*)malloc(bytes); }
                CK(cudaMemcpy(h0, ref, bytes, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(h1, buf, bytes, cudaMemcpyDeviceToHost));
                same = memcmp(h0, h1, bytes) == 0;
            }
            printf("{\"probe\": \"fq_mul\", \"dfma_warps\": \"%d/%d\", \"gmul_per_s\": %.2f, \"matches_integer_product\": %d}\n", mx[0], mx[1],
                   (double)iters * 4 * threads / (ms * 1e-3) / 1e9, same);
        }
        float ms = time_it([&] { k_mul<Fr, FrParams><<<grid, block>>>((Fr*)ref, iters, 0, 1); });
        float ms2 = time_it([&] { k_mul<Fr, FrParams><<<grid, block>>>((Fr*)buf, iters, 1, 1); });
        printf("{\"probe\": \"fr_mul\", \"int_gmul_per_s\": %.2f, \"dfma_gmul_per_s\": %.2f}\n", (double)iters * 4 * threads / (ms * 1e-3) / 1e9,
               (double)iters * 4 * threads / (ms2 * 1e-3) / 1e9);
    }
    return 0;
}
