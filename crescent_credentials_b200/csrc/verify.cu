// verify.cu -- Groth16 verification on the device ("next" row f-4 of the scope table).
//
// Reference lines: forks/groth16/src/verifier.rs:13-20 (prepare_verifying_key), :25-39 (prepare_inputs), :44-65
// (verify_proof_with_prepared_inputs), :69-76 (verify_proof).  The reference verifies one proof per call on the CPU; a
// verifier that checks many show proofs against one key is embarrassingly parallel over proofs, so the
// device form is: one thread per proof for prepare_inputs (fixed-base window tables of gamma_abc_g1, built once per key),
// one thread per proof for the three-pair Miller loop + final exponentiation (the lines of -gamma and -delta come from
// tables prepared once per key -- G2Prepared -- and are read by every thread at the same index, i.e. as broadcasts; the
// lines of B are computed on the fly).  Every verdict is the reference's verdict for that proof.
//
// Integer-pipe work per proof (Fq products): Miller loop 64 x (Fq12 squaring 36 + 3 sparse products 39 + B doubling
// step ~30) + 27 x (3 x 39 + addition step ~40) ~ 17.6 k; final exponentiation ~ 3 x (62 x 18 + 27 x 54) + ~1 k ~ 8.7 k;
// prepare_inputs 32 x 10 per public input.  No HBM traffic to speak of: the kernels are bound by the multiplier pipe and,
// at small batch sizes, by the latency of one thread's dependent chain.
#ifdef G16_VERIFY_MUL_CALL
#define G16_MUL_AS_CALL  // fp.cuh: the Montgomery product as one shared function for this translation unit
#endif
#include "internal.cuh"
#include "pairing.cuh"

namespace g16 {

static_assert(sizeof(g16_proof) == 272, "g16_proof layout");
static_assert(sizeof(Fq12) == 384 && sizeof(EllCoeff) == 192, "tower layout");

struct ProofIn {  // same bytes as g16_proof
    G1Affine a;
    G2Affine b;
    G1Affine c;
    int32_t a_inf, b_inf, c_inf, pad;
};
static_assert(sizeof(ProofIn) == sizeof(g16_proof), "ProofIn mirrors g16_proof");

// tbl[(i * 32 + w) * 255 + d - 1] = d * 2^(8w) * abc[i]: one thread per entry (a scalar multiplication by a one-byte scalar
// shifted into place, then one inversion); runs once per verifying key
__global__ void __launch_bounds__(128) k_abc_table(const G1Affine* __restrict__ abc, size_t n_points, G1Affine* __restrict__ tbl) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_point = (size_t)kAbcWindows * kAbcDigits;
    if (g >= n_points * per_point) return;
    size_t i = g / per_point;
    unsigned rem = (unsigned)(g - i * per_point);
    unsigned w = rem / kAbcDigits, d = rem % kAbcDigits + 1;
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    k[w >> 2] = d << (8 * (w & 3));
    tbl[g] = scalar_mul(G1XYZZ::from_affine(abc[i]), k).to_affine();
}

// G2Prepared of the points q[0..n): EllCoeff[kEllCoeffs] each (a point at infinity leaves its table untouched)
__global__ void k_g2_prepare(const G2Affine* __restrict__ q, unsigned n, bool negate, EllCoeff* __restrict__ out) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    G2Affine p = negate ? q[t].neg() : q[t];
    if (!p.is_inf()) g2_prepare(p, out + (size_t)t * kEllCoeffs);
}

// gt[i] = e(p[i], q[i])
__global__ void __launch_bounds__(32) k_pairing(const G1Affine* __restrict__ p, const G2Affine* __restrict__ q, size_t n,
                                                Fq12* __restrict__ gt) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    G1Affine p3[3] = {p[t], G1Affine::inf(), G1Affine::inf()};
    G2Affine q0 = q[t];
    bool act[3] = {!p3[0].is_inf() && !q0.is_inf(), false, false};
    Fq12 f = miller_loop3(p3, act, q0, nullptr, nullptr);
    bool ok;
    gt[t] = final_exponentiation(f, ok);
}

__global__ void __launch_bounds__(64) k_prepare_inputs(const G1Affine* __restrict__ abc0, const G1Affine* __restrict__ tbl,
                                                       const Fr* __restrict__ inputs, size_t n_inputs, size_t n,
                                                       G1Affine* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = prepare_inputs_one(*abc0, tbl, inputs + t * n_inputs, n_inputs);
}

// MINB = resident 32-thread blocks per SM the register allocation must allow: 8 leaves ptxas all 255 registers, 16 caps it at
// 128 (more warps to hide the dependent product chains, more spills of the Fq12 state) -- option "verify_occupancy"
template <int MINB>
__global__ void __launch_bounds__(32, MINB) k_verify(const ProofIn* __restrict__ proofs, const G1Affine* __restrict__ prepared, size_t n,
                                               const EllCoeff* __restrict__ neg_gamma, const EllCoeff* __restrict__ neg_delta,
                                               const Fq12* __restrict__ alpha_beta, unsigned g2_inf, uint8_t* __restrict__ verdict) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const ProofIn& pr = proofs[t];
    G1Affine a = pr.a_inf ? G1Affine::inf() : pr.a;
    G2Affine b = pr.b_inf ? G2Affine::inf() : pr.b;
    G1Affine c = pr.c_inf ? G1Affine::inf() : pr.c;
    verdict[t] = (uint8_t)verify_one(a, b, c, prepared[t], neg_gamma, neg_delta, *alpha_beta, g2_inf);
}

void verify_free(g16_ctx* ctx) {
    VerifyKeyDev& v = ctx->vk;
    dev_free(v.abc0);
    dev_free(v.abc_tbl);
    dev_free(v.neg_gamma);  // neg_delta points into the same allocation
    dev_free(v.alpha_beta);
    dev_free(v.proofs);
    dev_free(v.inputs);
    dev_free(v.prepared);
    dev_free(v.verdict);
    v = VerifyKeyDev();
}

static int verify_reserve(g16_ctx* ctx, size_t n) {
    VerifyKeyDev& v = ctx->vk;
    if (n <= v.cap) return G16_OK;
    dev_free(v.proofs);
    dev_free(v.inputs);
    dev_free(v.prepared);
    dev_free(v.verdict);
    v.proofs = nullptr;
    v.inputs = nullptr;
    v.prepared = nullptr;
    v.verdict = nullptr;
    v.cap = 0;
    G16_CUDA(ctx, cudaMalloc(&v.proofs, n * sizeof(g16_proof)));
    G16_TRY(dev_alloc(ctx, &v.inputs, n * (v.n_inputs ? v.n_inputs : 1)));
    G16_TRY(dev_alloc(ctx, &v.prepared, n));
    G16_TRY(dev_alloc(ctx, &v.verdict, n));
    v.cap = n;
    return G16_OK;
}

static int prepare_inputs_dev(g16_ctx* ctx, const Fr* inputs_dev, size_t n, G1Affine* out_dev, cudaStream_t st) {
    const VerifyKeyDev& v = ctx->vk;
    G16_LAUNCH(ctx, k_prepare_inputs, (unsigned)((n + 63) / 64), 64, 0, st, (const G1Affine*)v.abc0, (const G1Affine*)v.abc_tbl,
               inputs_dev, v.n_inputs, n, out_dev);
    return G16_OK;
}

static int launch_verify(g16_ctx* ctx, const void* proofs_dev, size_t n, uint8_t* verdict_dev, cudaStream_t st) {
    VerifyKeyDev& v = ctx->vk;
    const unsigned grid = (unsigned)((n + 31) / 32);
    if (ctx->opt_verify_occupancy == 16)
        G16_LAUNCH(ctx, k_verify<16>, grid, 32, 0, st, (const ProofIn*)proofs_dev, (const G1Affine*)v.prepared, n,
                   (const EllCoeff*)v.neg_gamma, (const EllCoeff*)v.neg_delta, (const Fq12*)v.alpha_beta, v.g2_inf, verdict_dev);
    else if (ctx->opt_verify_occupancy == 12)
        G16_LAUNCH(ctx, k_verify<12>, grid, 32, 0, st, (const ProofIn*)proofs_dev, (const G1Affine*)v.prepared, n,
                   (const EllCoeff*)v.neg_gamma, (const EllCoeff*)v.neg_delta, (const Fq12*)v.alpha_beta, v.g2_inf, verdict_dev);
    else
        G16_LAUNCH(ctx, k_verify<8>, grid, 32, 0, st, (const ProofIn*)proofs_dev, (const G1Affine*)v.prepared, n,
                   (const EllCoeff*)v.neg_gamma, (const EllCoeff*)v.neg_delta, (const Fq12*)v.alpha_beta, v.g2_inf, verdict_dev);
    return G16_OK;
}

static int verify_dev(g16_ctx* ctx, const void* proofs_dev, const Fr* inputs_dev, size_t n, uint8_t* verdict_dev, cudaStream_t st) {
    VerifyKeyDev& v = ctx->vk;
    G16_TRY(prepare_inputs_dev(ctx, inputs_dev, n, v.prepared, st));
    return launch_verify(ctx, proofs_dev, n, verdict_dev, st);
}

}  // namespace g16

using namespace g16;

namespace {
struct VGuard {
    g16_ctx* c;
    explicit VGuard(g16_ctx* ctx) : c(ctx) {
        c->mu.lock();
        cudaSetDevice(c->device);
    }
    ~VGuard() { c->mu.unlock(); }
};
}  // namespace

extern "C" {

int g16_ctx_load_vk(g16_ctx* ctx, const g16_vk_view* vk) {
    if (!ctx || !vk) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    if (!vk->alpha_g1 || !vk->beta_g2 || !vk->gamma_g2 || !vk->delta_g2 || !vk->gamma_abc_g1 || vk->gamma_abc_len == 0)
        return set_err(ctx, G16_ERR_BAD_ARG, "vk: null pointer or empty gamma_abc_g1");
    if (vk->encoding != G16_ENC_MONTGOMERY && vk->encoding != G16_ENC_CANONICAL)
        return set_err(ctx, G16_ERR_BAD_ARG, "vk: unknown encoding %d", vk->encoding);
    verify_free(ctx);
    VerifyKeyDev& v = ctx->vk;
    cudaStream_t st = ctx->main;
    v.n_inputs = vk->gamma_abc_len - 1;
    auto all_zero = [](const uint64_t* p, int words) {
        uint64_t o = 0;
        for (int i = 0; i < words; i++) o |= p[i];
        return o == 0;
    };
    v.g2_inf = (all_zero(vk->gamma_g2, 16) ? 2u : 0u) | (all_zero(vk->delta_g2, 16) ? 4u : 0u);
    // staging: alpha_g1 (64 B) | beta_g2 | gamma_g2 | delta_g2 (128 B each) | gamma_abc_g1
    const size_t head = 64 + 3 * 128;
    const size_t bytes = head + vk->gamma_abc_len * 64;
    char* stage = nullptr;
    G16_CUDA(ctx, cudaMalloc(&stage, bytes));
    int rc = G16_OK;
    auto up = [&](size_t off, const void* src, size_t len) {
        if (rc == G16_OK && cudaMemcpyAsync(stage + off, src, len, cudaMemcpyHostToDevice, st) != cudaSuccess)
            rc = set_err(ctx, G16_ERR_CUDA, "vk upload failed");
    };
    up(0, vk->alpha_g1, 64);
    up(64, vk->beta_g2, 128);
    up(192, vk->gamma_g2, 128);
    up(320, vk->delta_g2, 128);
    up(head, vk->gamma_abc_g1, vk->gamma_abc_len * 64);
    if (rc == G16_OK && vk->encoding == G16_ENC_CANONICAL) rc = convert_mont_dev(ctx, G16_FIELD_FQ, stage, bytes / 32, true, st);
    auto body = [&]() -> int {
        G16_TRY(dev_alloc(ctx, &v.abc0, 1));
        G16_CUDA(ctx, cudaMemcpyAsync(v.abc0, stage + head, 64, cudaMemcpyDeviceToDevice, st));
        if (v.n_inputs) {
            const size_t entries = v.n_inputs * (size_t)kAbcWindows * kAbcDigits;
            G16_TRY(dev_alloc(ctx, &v.abc_tbl, entries));
            G16_LAUNCH(ctx, k_abc_table, (unsigned)((entries + 127) / 128), 128, 0, st, (const G1Affine*)(stage + head + 64),
                       v.n_inputs, v.abc_tbl);
        }
        // gamma_g2_neg_pc, delta_g2_neg_pc: gamma and delta are adjacent in the staging buffer, one launch prepares both tables
        G16_CUDA(ctx, cudaMalloc(&v.neg_gamma, 2 * kEllCoeffs * sizeof(EllCoeff)));
        v.neg_delta = (EllCoeff*)v.neg_gamma + kEllCoeffs;
        G16_CUDA(ctx, cudaMemsetAsync(v.neg_gamma, 0, 2 * kEllCoeffs * sizeof(EllCoeff), st));
        G16_LAUNCH(ctx, k_g2_prepare, 1, 32, 0, st, (const G2Affine*)(stage + 192), 2u, true, (EllCoeff*)v.neg_gamma);
        G16_CUDA(ctx, cudaMalloc(&v.alpha_beta, sizeof(Fq12)));
        G16_LAUNCH(ctx, k_pairing, 1, 32, 0, st, (const G1Affine*)stage, (const G2Affine*)(stage + 64), (size_t)1, (Fq12*)v.alpha_beta);
        G16_CUDA(ctx, cudaMemcpyAsync(v.alpha_beta_host, v.alpha_beta, sizeof(Fq12), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        return G16_OK;
    };
    if (rc == G16_OK) rc = body();
    cudaStreamSynchronize(st);
    cudaFree(stage);
    if (rc != G16_OK) {
        std::string keep = ctx->err;
        verify_free(ctx);
        ctx->err = keep;
        return rc;
    }
    v.have = true;
    return G16_OK;
}

int g16_vk_alpha_beta(g16_ctx* ctx, uint64_t out[48]) {
    if (!ctx || !out) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    if (!ctx->vk.have) return set_err(ctx, G16_ERR_BAD_ARG, "no verifying key loaded");
    for (int i = 0; i < 48; i++) out[i] = ctx->vk.alpha_beta_host[i];
    return G16_OK;
}

int g16_prepare_inputs(g16_ctx* ctx, const uint64_t* public_inputs, size_t n, uint64_t* out_points) {
    if (!ctx || !out_points) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    VerifyKeyDev& v = ctx->vk;
    if (!v.have) return set_err(ctx, G16_ERR_BAD_ARG, "no verifying key loaded");
    if (n == 0) return G16_OK;
    if (v.n_inputs && !public_inputs) return set_err(ctx, G16_ERR_BAD_ARG, "public_inputs is NULL");
    G16_TRY(verify_reserve(ctx, n));
    cudaStream_t st = ctx->main;
    if (v.n_inputs) G16_CUDA(ctx, cudaMemcpyAsync(v.inputs, public_inputs, n * v.n_inputs * 32, cudaMemcpyHostToDevice, st));
    G16_TRY(prepare_inputs_dev(ctx, v.inputs, n, v.prepared, st));
    G16_CUDA(ctx, cudaMemcpyAsync(out_points, v.prepared, n * 64, cudaMemcpyDeviceToHost, st));
    G16_CUDA(ctx, cudaStreamSynchronize(st));
    return G16_OK;
}

int g16_verify_batch(g16_ctx* ctx, const g16_proof* proofs, const uint64_t* public_inputs, size_t n, uint8_t* verdict) {
    if (!ctx) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    VerifyKeyDev& v = ctx->vk;
    if (!v.have) return set_err(ctx, G16_ERR_BAD_ARG, "no verifying key loaded");
    if (n == 0) return G16_OK;
    if (!proofs || !verdict || (v.n_inputs && !public_inputs)) return set_err(ctx, G16_ERR_BAD_ARG, "verify: null pointer");
    G16_TRY(verify_reserve(ctx, n));
    cudaStream_t st = ctx->main;
    G16_CUDA(ctx, cudaMemcpyAsync(v.proofs, proofs, n * sizeof(g16_proof), cudaMemcpyHostToDevice, st));
    if (v.n_inputs) G16_CUDA(ctx, cudaMemcpyAsync(v.inputs, public_inputs, n * v.n_inputs * 32, cudaMemcpyHostToDevice, st));
    G16_TRY(verify_dev(ctx, v.proofs, v.inputs, n, v.verdict, st));
    G16_CUDA(ctx, cudaMemcpyAsync(verdict, v.verdict, n, cudaMemcpyDeviceToHost, st));
    G16_CUDA(ctx, cudaStreamSynchronize(st));
    return G16_OK;
}

int g16_verify_batch_prepared(g16_ctx* ctx, const g16_proof* proofs, const uint64_t* prepared_inputs, size_t n, uint8_t* verdict) {
    if (!ctx) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    VerifyKeyDev& v = ctx->vk;
    if (!v.have) return set_err(ctx, G16_ERR_BAD_ARG, "no verifying key loaded");
    if (n == 0) return G16_OK;
    if (!proofs || !verdict || !prepared_inputs) return set_err(ctx, G16_ERR_BAD_ARG, "verify: null pointer");
    G16_TRY(verify_reserve(ctx, n));
    cudaStream_t st = ctx->main;
    G16_CUDA(ctx, cudaMemcpyAsync(v.proofs, proofs, n * sizeof(g16_proof), cudaMemcpyHostToDevice, st));
    G16_CUDA(ctx, cudaMemcpyAsync(v.prepared, prepared_inputs, n * 64, cudaMemcpyHostToDevice, st));
    G16_TRY(launch_verify(ctx, v.proofs, n, v.verdict, st));
    G16_CUDA(ctx, cudaMemcpyAsync(verdict, v.verdict, n, cudaMemcpyDeviceToHost, st));
    G16_CUDA(ctx, cudaStreamSynchronize(st));
    return G16_OK;
}

int g16_verify_batch_dev(g16_ctx* ctx, const void* proofs_dev, const void* public_inputs_dev, size_t n, void* verdict_dev) {
    if (!ctx) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    VerifyKeyDev& v = ctx->vk;
    if (!v.have) return set_err(ctx, G16_ERR_BAD_ARG, "no verifying key loaded");
    if (n == 0) return G16_OK;
    if (!proofs_dev || !verdict_dev || (v.n_inputs && !public_inputs_dev)) return set_err(ctx, G16_ERR_BAD_ARG, "verify: null pointer");
    G16_TRY(verify_reserve(ctx, n));
    return verify_dev(ctx, proofs_dev, (const Fr*)public_inputs_dev, n, (uint8_t*)verdict_dev, ctx->main);
}

int g16_pairing(g16_ctx* ctx, const uint64_t* g1_points, const uint64_t* g2_points, size_t n, uint64_t* gt_out) {
    if (!ctx) return G16_ERR_BAD_ARG;
    VGuard g(ctx);
    if (n == 0) return G16_OK;
    if (!g1_points || !g2_points || !gt_out) return set_err(ctx, G16_ERR_BAD_ARG, "pairing: null pointer");
    cudaStream_t st = ctx->main;
    char* buf = nullptr;
    G16_CUDA(ctx, cudaMalloc(&buf, n * (64 + 128 + sizeof(Fq12))));
    G1Affine* p = (G1Affine*)buf;
    G2Affine* q = (G2Affine*)(buf + n * 64);
    Fq12* gt = (Fq12*)(buf + n * (64 + 128));
    auto body = [&]() -> int {
        G16_CUDA(ctx, cudaMemcpyAsync(p, g1_points, n * 64, cudaMemcpyHostToDevice, st));
        G16_CUDA(ctx, cudaMemcpyAsync(q, g2_points, n * 128, cudaMemcpyHostToDevice, st));
        G16_LAUNCH(ctx, k_pairing, (unsigned)((n + 31) / 32), 32, 0, st, (const G1Affine*)p, (const G2Affine*)q, n, gt);
        G16_CUDA(ctx, cudaMemcpyAsync(gt_out, gt, n * sizeof(Fq12), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        return G16_OK;
    };
    int rc = body();
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return rc;
}

}  // extern "C"
