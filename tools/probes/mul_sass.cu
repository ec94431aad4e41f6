// mul_sass.cu -- one Montgomery product per thread (Fp::mul_cios of csrc/fp.cuh), compiled only to read its SASS:
//   bash tools/sass_mul.sh   ->  profiles/r02_sass_mul_cios.txt (mnemonic histogram + the instruction stream)
#include "../../crescent_credentials_b200/csrc/fp.cuh"
using namespace g16;
extern "C" __global__ void k_one_mul_fq(const Fq* __restrict__ a, const Fq* __restrict__ b, Fq* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    out[i] = a[i] * b[i];
}
extern "C" __global__ void k_one_fq2_mul(const Fq2* __restrict__ a, const Fq2* __restrict__ b, Fq2* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    out[i] = a[i] * b[i];
}
