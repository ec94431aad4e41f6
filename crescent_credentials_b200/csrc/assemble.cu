// assemble.cu -- proof assembly (K9): the handful of scalar multiplications, additions and normalisations of
// create_proof_with_assignment / calculate_coeff (forks/groth16/src/prover.rs:76-135, 256-274).
//
//   g_a  = r*delta_g1 + a_query[0] + MSM_a + alpha_g1                         (:94-96, :256-274)
//   g1_b = s*delta_g1 + b_g1_query[0] + MSM_b1 + beta_g1     (zero when r == 0, :102-112)
//   g2_b = s*delta_g2 + b_g2_query[0] + MSM_b2 + beta_g2                      (:116-117)
//   g_c  = s*g_a + r*g1_b - (r*s)*delta_g1 + MSM_l + MSM_h                    (:98, :118, :124-128)
//   Proof{ a: g_a, b: g2_b, c: g_c } in affine form                          (:131-135)
//
// A 254-bit scalar multiplication of a fresh point is ~4000 dependent field products: ~1.6 ms for a lone lane, which
// would sit at the very end of the proof.  Instead the sums are split by linearity into
//   T_a = r*delta_g1 + a_query[0] + alpha_g1,   T_b = s*delta_g1 + b_g1_query[0] + beta_g1       ((r, s, pk) only)
//   g_a = T_a + MSM_a,        s*g_a  = s*T_a + s*MSM_a
//   g1_b = T_b + MSM_b1,      r*g1_b = r*T_b + r*MSM_b1
// k_assemble_pre computes T_a, s*T_a, r*T_b, rs*delta_g1 and T_b2 on a side stream at the start of the proof;
// k_scale_point computes s*MSM_a / r*MSM_b1 (per rank, on the MSM's own stream) as soon as that MSM is done, while the
// other MSMs are still running; k_assemble_post only adds (summing over the `count` shard partials) and normalises.
#include "internal.cuh"
#include "glv.cuh"

namespace g16 {

struct AsmConsts {
    G1Affine alpha_g1, beta_g1, delta_g1, a0, b1_0;
    G2Affine beta_g2, delta_g2, b2_0;
};
struct AsmPre {
    G1XYZZ t_a, s_ta, r_tb, rs_d1;
    G2XYZZ t_b2;
};
struct ProofDev {
    G1Affine a;
    G2Affine b;
    G1Affine c;
};
struct Scalar256 {
    uint32_t w[8];
};
// (r, s, r*s) in canonical form + the r == 0 flag (prover.rs:102), written per proof into ctx->d_small by assemble_set_scalars.
// The kernels read them from device memory rather than taking them by value: the launch sequence of a proof is then the same
// from proof to proof, which is what lets api.cu replay it as a CUDA graph.
struct AsmScalars {
    Scalar256 r, s, rs;
    int r_is_zero;
    int pad[7];
    GlvScalar glv[2];  // GLV splits of r (0) and s (1) for k_scale_point
};

__global__ void k_assemble_pre(AsmConsts k, const AsmScalars* __restrict__ sc, AsmPre* out) {
    unsigned wid = threadIdx.x >> 5;
    if (threadIdx.x & 31) return;  // no barrier in this kernel
    const Scalar256 r = sc->r, s = sc->s, rs = sc->rs;
    const int r_is_zero = sc->r_is_zero;
    if (wid == 0) {
        G1XYZZ t = scalar_mul(G1XYZZ::from_affine(k.delta_g1), r.w);
        t.madd(k.a0);
        t.madd(k.alpha_g1);
        out->t_a = t;
        out->s_ta = scalar_mul(t, s.w);
    }
    if (wid == 1) {
        G1XYZZ t = G1XYZZ::inf();
        if (!r_is_zero) {  // prover.rs:102: g1_b is only computed when r != 0
            t = scalar_mul(G1XYZZ::from_affine(k.delta_g1), s.w);
            t.madd(k.b1_0);
            t.madd(k.beta_g1);
            t = scalar_mul(t, r.w);
        }
        out->r_tb = t;
    }
    if (wid == 2) out->rs_d1 = scalar_mul(G1XYZZ::from_affine(k.delta_g1), rs.w);
    if (wid == 3) {
        G2XYZZ t = scalar_mul(G2XYZZ::from_affine(k.delta_g2), s.w);
        t.madd(k.b2_0);
        t.madd(k.beta_g2);
        out->t_b2 = t;
    }
}

// ---- the same (r, s, pk)-only points from per-key fixed-base tables -------------------------------------------------------
// Every scalar multiplication of k_assemble_pre has a base that depends on the key alone once the products are expanded:
//   T_a = r*delta_g1 + U,  s*T_a = (rs)*delta_g1 + s*U,  r*T_b = (rs)*delta_g1 + r*V,  T_b2 = s*delta_g2 + W
// with U = a_query[0] + alpha_g1, V = b_g1_query[0] + beta_g1, W = b_g2_query[0] + beta_g2.  With 2^(8j) * base, j < 32, stored
// per key (10 KB), one warp computes a 254-bit multiple with every lane working on one byte of the scalar (8 doublings and at
// most 8 mixed additions) and a 5-step tree sum: ~250 dependent field products instead of ~4000.  Large circuits never saw
// this kernel (it hides under the MSMs); for the ~1 k-constraint circuits the reference also proves (creds/src/rangeproof.rs:
// 490-511, creds/benches/proof_benchmark.rs:85-96) it was the floor of the proof latency: 4.4 ms.
struct AsmTables {
    G1Affine d1[32], u[32], v[32];  // 2^(8j) * {delta_g1, U, V}
    G2Affine d2[32];                // 2^(8j) * delta_g2
    G2Affine w;                     // W
};

__global__ void k_asm_tables(AsmConsts k, AsmTables* out) {
    const unsigned t = threadIdx.x, j = t & 31, which = t >> 5;
    uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e[j >> 2] = 1u << (8 * (j & 3));
    if (which < 3) {
        G1XYZZ b = G1XYZZ::from_affine(which == 0 ? k.delta_g1 : which == 1 ? k.a0 : k.b1_0);
        if (which == 1) b.madd(k.alpha_g1);
        if (which == 2) b.madd(k.beta_g1);
        G1Affine p = scalar_mul(b, e).to_affine();
        (which == 0 ? out->d1 : which == 1 ? out->u : out->v)[j] = p;
    } else {
        out->d2[j] = scalar_mul(G2XYZZ::from_affine(k.delta_g2), e).to_affine();
        if (j == 0) {
            G2XYZZ w = G2XYZZ::from_affine(k.b2_0);
            w.madd(k.beta_g2);
            out->w = w.to_affine();
        }
    }
}

// byte j of the scalar times tbl[j] (= 2^(8j) * base)
template <class F>
__device__ __forceinline__ XYZZ<F> byte_mul(const Affine<F>& p, uint32_t byte) {
    XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
    for (int bit = 7; bit >= 0; bit--) {
        acc = acc.dbl();
        if ((byte >> bit) & 1u) acc.madd(p);
    }
    return acc;
}
template <class F>
__device__ __forceinline__ void warp_tree_sum(XYZZ<F>* sh, unsigned lane) {  // sh[0] = sum of the 32 entries
#pragma unroll 1
    for (unsigned off = 16; off >= 1; off >>= 1) {
        __syncwarp();
        if (lane < off) {
            XYZZ<F> a = sh[lane];
            a.add(sh[lane + off]);
            sh[lane] = a;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(160) k_assemble_pre_tables(const AsmTables* __restrict__ T, const AsmScalars* __restrict__ sc,
                                                             AsmPre* out) {
    __shared__ G1XYZZ sh1[4][32];  // r*delta_g1, rs*delta_g1, s*U, r*V
    __shared__ G2XYZZ sh2[32];     // s*delta_g2
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const Scalar256 r = sc->r, s = sc->s, rs = sc->rs;
    const int r_is_zero = sc->r_is_zero;
    auto byte_of = [&](const Scalar256& k) { return (k.w[lane >> 2] >> (8 * (lane & 3))) & 0xffu; };
    if (wid < 4) {
        const Scalar256& k = wid == 0 ? r : wid == 1 ? rs : wid == 2 ? s : r;
        const G1Affine* tbl = wid <= 1 ? T->d1 : wid == 2 ? T->u : T->v;
        sh1[wid][lane] = byte_mul(tbl[lane], byte_of(k));
        warp_tree_sum(sh1[wid], lane);
    } else {
        sh2[lane] = byte_mul(T->d2[lane], byte_of(s));
        warp_tree_sum(sh2, lane);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        G1XYZZ t = sh1[0][0];
        t.madd(T->u[0]);  // U
        out->t_a = t;
    } else if (threadIdx.x == 32) {
        G1XYZZ t = sh1[1][0];
        out->rs_d1 = t;
        t.add(sh1[2][0]);
        out->s_ta = t;
    } else if (threadIdx.x == 64) {
        G1XYZZ t = G1XYZZ::inf();
        if (!r_is_zero) {  // prover.rs:102
            t = sh1[1][0];
            t.add(sh1[3][0]);
        }
        out->r_tb = t;
    } else if (threadIdx.x == 128) {
        G2XYZZ t = sh2[0];
        t.madd(T->w);
        out->t_b2 = t;
    }
}

// out = k * in for one G1 point (latency-bound, runs beside the remaining MSMs; exposed at the very end of a proof when the a /
// b_g1 MSMs are the last to finish).  Two warps, one lane each: k = k1 + k2 * lambda, warp 0 computes k1 * P, warp 1 k2 * phi(P)
// (glv.cuh), lane 0 adds.  `which` = 0: k = r, 1: k = s.
__global__ void __launch_bounds__(64) k_scale_point(const G1XYZZ* __restrict__ in, const AsmScalars* __restrict__ sc, int which,
                                                    G1XYZZ* __restrict__ out) {
    __shared__ G1XYZZ half[2];
    const unsigned w = threadIdx.x >> 5;
    const GlvScalar& g = sc->glv[which];
    if ((threadIdx.x & 31) == 0) {
        if (g.ok) {
            half[w] = glv_half_mul(*in, g.h[w], (int)w);
        } else if (w == 0) {  // magnitudes out of range (not observed): plain windowed multiplication
            const Scalar256 kk = which == 0 ? sc->r : sc->s;
            half[0] = scalar_mul_window(*in, kk.w);
            half[1] = G1XYZZ::inf();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        G1XYZZ r = half[0];
        r.add_inl(half[1]);
        *out = r;
    }
}

// partial layout (u64 words): h[16] l[16] a[16] s*a[16] r*b_g1[16] b_g2[32]  == 5 x G1XYZZ + 1 x G2XYZZ
struct PartialDev {
    G1XYZZ h, l, a, sa, rb1;
    G2XYZZ b2;
};
static_assert(sizeof(PartialDev) == sizeof(g16_partial), "partial layout");

__global__ void k_assemble_post(const AsmPre* pre, const PartialDev* parts, int count, ProofDev* out) {
    const unsigned wid = (threadIdx.x & 31) ? 99u : (threadIdx.x >> 5);  // one lane per warp runs a chain
    if (wid == 0) {
        G1XYZZ ga = pre->t_a;
        for (int i = 0; i < count; i++) ga.add(parts[i].a);
        out->a = ga.to_affine();
    } else if (wid == 1) {
        G2XYZZ g2 = pre->t_b2;
        for (int i = 0; i < count; i++) g2.add(parts[i].b2);
        out->b = g2.to_affine();
    } else if (wid == 2) {
        G1XYZZ gc = pre->s_ta;
        gc.add(pre->r_tb);
        gc.add(pre->rs_d1.neg());
        for (int i = 0; i < count; i++) {
            gc.add(parts[i].sa);
            gc.add(parts[i].rb1);
            gc.add(parts[i].l);
            gc.add(parts[i].h);
        }
        out->c = gc.to_affine();
    }
}

template <class F>
__global__ void k_xyzz_to_affine(const XYZZ<F>* in, Affine<F>* out, int count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = in[0];
    for (int i = 1; i < count; i++) acc.add(in[i]);
    *out = acc.to_affine();
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

// d_small layout: [AsmPre @0][ProofDev @1024][generic affine scratch @2048][AsmScalars @2560][msm_can_share counters @3072]
static AsmPre* pre_ptr(g16_ctx* ctx) { return (AsmPre*)ctx->d_small; }
static ProofDev* proof_ptr(g16_ctx* ctx) { return (ProofDev*)((char*)ctx->d_small + 1024); }
static void* affine_ptr(g16_ctx* ctx) { return (char*)ctx->d_small + 2048; }
static AsmScalars* scalars_ptr(g16_ctx* ctx) { return (AsmScalars*)((char*)ctx->d_small + 2560); }
static_assert(sizeof(AsmPre) <= 1024 && sizeof(ProofDev) <= 1024 && sizeof(G2Affine) <= 512 && sizeof(AsmScalars) <= 512, "d_small layout");

// (r, s) of the proof being queued -> device (stream-ordered; staged through the context's page-locked block so that the copy
// is a true asynchronous DMA and the staging memory outlives the call)
int assemble_set_scalars(g16_ctx* ctx, const uint64_t* r, const uint64_t* s, cudaStream_t st) {
    Fr rm = fr_load(r), sm = fr_load(s);
    Fr rs = rm * sm;
    if (!ctx->h_scalars) G16_CUDA(ctx, cudaHostAlloc(&ctx->h_scalars, 2 * sizeof(AsmScalars), cudaHostAllocDefault));
    // two slots, alternating: the previous proof's copy may still be in flight when the next one is queued (the *_dev entry
    // points do not synchronise)
    AsmScalars* h = (AsmScalars*)ctx->h_scalars + (ctx->h_scalars_slot ^= 1);
    memset(h, 0, sizeof(*h));
    h->r = canon(rm);
    h->s = canon(sm);
    h->rs = canon(rs);
    h->r_is_zero = (int)rm.is_zero();
    h->glv[0] = glv_decompose(h->r.w);
    h->glv[1] = glv_decompose(h->s.w);
    G16_CUDA(ctx, cudaMemcpyAsync(scalars_ptr(ctx), h, sizeof(AsmScalars), cudaMemcpyHostToDevice, st));
    return G16_OK;
}

// per-key tables for k_assemble_pre_tables; called by g16_ctx_load_pk once the single points are in the context
int assemble_build_tables(g16_ctx* ctx, cudaStream_t st) {
    if (!ctx->d_asm_tables) G16_CUDA(ctx, cudaMalloc(&ctx->d_asm_tables, sizeof(AsmTables)));
    G16_LAUNCH(ctx, k_asm_tables, 1, 128, 0, st, consts_of(ctx), (AsmTables*)ctx->d_asm_tables);
    return G16_OK;
}

// the scalars must already be on the device (assemble_set_scalars, ordered before `st`)
int assemble_pre(g16_ctx* ctx, cudaStream_t st) {
    if (ctx->opt_asm_tables && ctx->d_asm_tables) {
        G16_LAUNCH(ctx, k_assemble_pre_tables, 1, 160, 0, st, (const AsmTables*)ctx->d_asm_tables, (const AsmScalars*)scalars_ptr(ctx),
                   pre_ptr(ctx));
        return G16_OK;
    }
    G16_LAUNCH(ctx, k_assemble_pre, 1, 128, 0, st, consts_of(ctx), (const AsmScalars*)scalars_ptr(ctx), pre_ptr(ctx));
    return G16_OK;
}

// which = 0: k = r, 1: k = s (of the scalars assemble_set_scalars put on the device)
int scale_point_dev(g16_ctx* ctx, const void* in_xyzz, int which, void* out_xyzz, cudaStream_t st) {
    G16_LAUNCH(ctx, k_scale_point, 1, 64, 0, st, (const G1XYZZ*)in_xyzz, (const AsmScalars*)scalars_ptr(ctx), which, (G1XYZZ*)out_xyzz);
    return G16_OK;
}

int assemble_proof_queue(g16_ctx* ctx, const void* partials_dev, int count, cudaStream_t st) {
    G16_LAUNCH(ctx, k_assemble_post, 1, 128, 0, st, (const AsmPre*)pre_ptr(ctx), (const PartialDev*)partials_dev, count,
               proof_ptr(ctx));
    if (ctx->capturing) G16_CUDA(ctx, cudaEventRecordWithFlags(ctx->ev_t[15], st, cudaEventRecordExternal));
    else G16_CUDA(ctx, cudaEventRecord(ctx->ev_t[15], st));
    G16_CUDA(ctx, cudaMemcpyAsync(ctx->h_proof, proof_ptr(ctx), sizeof(ProofDev), cudaMemcpyDeviceToHost, st));
    return G16_OK;
}

void assemble_proof_read(g16_ctx* ctx, g16_proof* out) {
    ProofDev host;
    memcpy(&host, ctx->h_proof, sizeof(host));
    memset(out, 0, sizeof(*out));
    memcpy(out->a, &host.a, 64);
    memcpy(out->b, &host.b, 128);
    memcpy(out->c, &host.c, 64);
    out->a_inf = host.a.is_inf();
    out->b_inf = host.b.is_inf();
    out->c_inf = host.c.is_inf();
}

int xyzz_to_affine_host(g16_ctx* ctx, int group, const void* xyzz_dev, uint64_t* out, int* out_inf, cudaStream_t st) {
    return sum_partials_to_affine_host(ctx, group, xyzz_dev, 1, out, out_inf, st);
}

int sum_partials_to_affine_host(g16_ctx* ctx, int group, const void* xyzz_dev, int count, uint64_t* out, int* out_inf, cudaStream_t st) {
    if (group == 1) {
        G16_LAUNCH(ctx, k_xyzz_to_affine<Fq>, 1, 32, 0, st, (const G1XYZZ*)xyzz_dev, (G1Affine*)affine_ptr(ctx), count);
        G1Affine h;
        G16_CUDA(ctx, cudaMemcpyAsync(&h, affine_ptr(ctx), sizeof(h), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(out, &h, sizeof(h));
        if (out_inf) *out_inf = h.is_inf();
    } else {
        G16_LAUNCH(ctx, k_xyzz_to_affine<Fq2>, 1, 32, 0, st, (const G2XYZZ*)xyzz_dev, (G2Affine*)affine_ptr(ctx), count);
        G2Affine h;
        G16_CUDA(ctx, cudaMemcpyAsync(&h, affine_ptr(ctx), sizeof(h), cudaMemcpyDeviceToHost, st));
        G16_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(out, &h, sizeof(h));
        if (out_inf) *out_inf = h.is_inf();
    }
    return G16_OK;
}

}  // namespace g16
