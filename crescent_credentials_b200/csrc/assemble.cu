// assemble.cu -- proof assembly (K9): the handful of scalar multiplications, additions and normalisations of
// create_proof_with_assignment / calculate_coeff (forks/groth16/src/prover.rs:76-135, 256-274).
//
//   g_a  = r*delta_g1 + a_query[0] + MSM_a + alpha_g1                         (:94-96, :256-274)
//   g1_b = s*delta_g1 + b_g1_query[0] + MSM_b1 + beta_g1     (zero when r == 0, :102-112)
//   g2_b = s*delta_g2 + b_g2_query[0] + MSM_b2 + beta_g2                      (:116-117)
//   g_c  = s*g_a + r*g1_b - (r*s)*delta_g1 + MSM_l + MSM_h                    (:98, :118, :124-128)
//   Proof{ a: g_a, b: g2_b, c: g_c } in affine form                          (:131-135)
//
// The four scalar multiplications that depend only on (r, s, pk) run in k_assemble_pre on a side stream while the
// witness map and MSMs are in flight; k_assemble_post runs after the MSM results (summed over `count` shard partials).
#include "internal.cuh"

namespace g16 {

struct AsmConsts {
    G1Affine alpha_g1, beta_g1, delta_g1, a0, b1_0;
    G2Affine beta_g2, delta_g2, b2_0;
};
struct AsmPre {
    G1XYZZ r_d1, s_d1, rs_d1;
    G2XYZZ s_d2;
};
struct ProofDev {
    G1Affine a;
    G2Affine b;
    G1Affine c;
};
struct Scalar256 {
    uint32_t w[8];
};

__global__ void k_assemble_pre(AsmConsts k, Scalar256 r, Scalar256 s, Scalar256 rs, AsmPre* out) {
    unsigned wid = threadIdx.x >> 5;
    if (threadIdx.x & 31) return;  // no barrier in this kernel
    if (wid == 0) out->r_d1 = scalar_mul(G1XYZZ::from_affine(k.delta_g1), r.w);
    if (wid == 1) out->s_d1 = scalar_mul(G1XYZZ::from_affine(k.delta_g1), s.w);
    if (wid == 2) out->rs_d1 = scalar_mul(G1XYZZ::from_affine(k.delta_g1), rs.w);
    if (wid == 3) out->s_d2 = scalar_mul(G2XYZZ::from_affine(k.delta_g2), s.w);
}

// partial layout (u64 words): h[16] l[16] a[16] b_g1[16] b_g2[32]  == 4 x G1XYZZ + 1 x G2XYZZ
struct PartialDev {
    G1XYZZ h, l, a, b1;
    G2XYZZ b2;
};
static_assert(sizeof(PartialDev) == sizeof(g16_partial), "partial layout");

__global__ void k_assemble_post(AsmConsts k, const AsmPre* pre, const PartialDev* parts, int count, Scalar256 r,
                                Scalar256 s, int r_is_zero, ProofDev* out) {
    __shared__ G1XYZZ sh_sa, sh_rb;
    const unsigned wid = (threadIdx.x & 31) ? 99u : (threadIdx.x >> 5);  // one lane per warp runs a chain
    if (wid == 0) {
        G1XYZZ ga = pre->r_d1;
        ga.madd(k.a0);
        for (int i = 0; i < count; i++) ga.add(parts[i].a);
        ga.madd(k.alpha_g1);
        out->a = ga.to_affine();
        sh_sa = scalar_mul(ga, s.w);
    } else if (wid == 1) {
        G1XYZZ gb = G1XYZZ::inf();
        if (!r_is_zero) {
            gb = pre->s_d1;
            gb.madd(k.b1_0);
            for (int i = 0; i < count; i++) gb.add(parts[i].b1);
            gb.madd(k.beta_g1);
            gb = scalar_mul(gb, r.w);
        }
        sh_rb = gb;
    } else if (wid == 2) {
        G2XYZZ g2 = pre->s_d2;
        g2.madd(k.b2_0);
        for (int i = 0; i < count; i++) g2.add(parts[i].b2);
        g2.madd(k.beta_g2);
        out->b = g2.to_affine();
    }
    // warps 0..2 finished their chains; warp 3 waits and finishes C
    __syncthreads();
    if (wid == 3) {
        G1XYZZ gc = sh_sa;
        gc.add(sh_rb);
        gc.add(pre->rs_d1.neg());
        for (int i = 0; i < count; i++) gc.add(parts[i].l);
        for (int i = 0; i < count; i++) gc.add(parts[i].h);
        out->c = gc.to_affine();
    }
}

template <class F>
__global__ void k_xyzz_to_affine(const XYZZ<F>* in, Affine<F>* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = in->to_affine();
}

static AsmConsts consts_of(const g16_ctx* ctx) {
    AsmConsts k;
    k.alpha_g1 = ctx->alpha_g1;
    k.beta_g1 = ctx->beta_g1;
    k.delta_g1 = ctx->delta_g1;
    k.a0 = ctx->a0;
    k.b1_0 = ctx->b1_0;
    k.beta_g2 = ctx->beta_g2;
    k.delta_g2 = ctx->delta_g2;
    k.b2_0 = ctx->b2_0;
    return k;
}
static Fr fr_load(const uint64_t* p) {
    Fr x;
    for (int i = 0; i < 4; i++) {
        x.v[2 * i] = (uint32_t)p[i];
        x.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
    }
    return x;
}
static Scalar256 canon(const Fr& m) {
    Fr c = m.from_mont();
    Scalar256 s;
    for (int i = 0; i < 8; i++) s.w[i] = c.v[i];
    return s;
}

// d_small layout: [AsmPre][ProofDev][generic affine scratch]
static AsmPre* pre_ptr(g16_ctx* ctx) { return (AsmPre*)ctx->d_small; }
static ProofDev* proof_ptr(g16_ctx* ctx) { return (ProofDev*)((char*)ctx->d_small + 1024); }
static void* affine_ptr(g16_ctx* ctx) { return (char*)ctx->d_small + 2048; }

int assemble_pre(g16_ctx* ctx, const uint64_t* r, const uint64_t* s, cudaStream_t st) {
    Fr rm = fr_load(r), sm = fr_load(s);
    Fr rs = rm * sm;
    G16_LAUNCH(ctx, k_assemble_pre, 1, 128, 0, st, consts_of(ctx), canon(rm), canon(sm), canon(rs), pre_ptr(ctx));
    return G16_OK;
}

int assemble_proof(g16_ctx* ctx, const void* partials_dev, int count, const uint64_t* r, const uint64_t* s, g16_proof* out,
                   cudaStream_t st) {
    Fr rm = fr_load(r), sm = fr_load(s);
    G16_LAUNCH(ctx, k_assemble_post, 1, 128, 0, st, consts_of(ctx), (const AsmPre*)pre_ptr(ctx),
               (const PartialDev*)partials_dev, count, canon(rm), canon(sm), (int)rm.is_zero(), proof_ptr(ctx));
    ProofDev host;
    G16_CUDA(ctx, cudaMemcpyAsync(&host, proof_ptr(ctx), sizeof(ProofDev), cudaMemcpyDeviceToHost, st));
    G16_CUDA(ctx, cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    memcpy(out->a, &host.a, 64);
    memcpy(out->b, &host.b, 128);
    memcpy(out->c, &host.c, 64);
    out->a_inf = host.a.is_inf();
    out->b_inf = host.b.is_inf();
    out->c_inf = host.c.is_inf();
    return G16_OK;
}

int xyzz_to_affine_host(g16_ctx* ctx, int group, const void* xyzz_dev, uint64_t* out, int* out_inf, cudaStream_t st) {
    if (group == 1) {
        G16_LAUNCH(ctx, k_xyzz_to_affine<Fq>, 1, 32, 0, st, (const G1XYZZ*)xyzz_dev, (G1Affine*)affine_ptr(ctx));
        G1Affine h;
        G16_CUDA(ctx, cudaMemcpyAsync(&h, affine_ptr(ctx), sizeof(h), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(out, &h, sizeof(h));
        if (out_inf) *out_inf = h.is_inf();
    } else {
        G16_LAUNCH(ctx, k_xyzz_to_affine<Fq2>, 1, 32, 0, st, (const G2XYZZ*)xyzz_dev, (G2Affine*)affine_ptr(ctx));
        G2Affine h;
        G16_CUDA(ctx, cudaMemcpyAsync(&h, affine_ptr(ctx), sizeof(h), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(out, &h, sizeof(h));
        if (out_inf) *out_inf = h.is_inf();
    }
    return G16_OK;
}

}  // namespace g16
