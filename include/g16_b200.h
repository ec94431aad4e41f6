/*
 * g16_b200.h -- C ABI of libg16b200.so: a B200 (sm_100a) Groth16 prover for BN254 that drops in for the hot
 * path of microsoft/crescent-credentials' `prove` step.
 *
 * Reference interface replaced (paths relative to the reference tree):
 *   forks/groth16/src/prover.rs:26-34     Groth16::create_proof_with_reduction_and_matrices  -> g16_prove
 *   forks/groth16/src/prover.rs:54-136    create_proof_with_assignment (5 MSMs + assembly)   -> g16_prove (second half),
 *                                                                                               g16_prove_shard / g16_prove_combine
 *   forks/groth16/src/r1cs_to_qap.rs:150-213  LibsnarkReduction::witness_map_from_matrices   -> g16_witness_map(reduction=0)
 *   forks/circom-compat/src/circom/qap.rs:25-90  CircomReduction::witness_map_from_matrices  -> g16_witness_map(reduction=1)
 *   ark-ec VariableBaseMSM::msm_bigint call sites prover.rs:66,74,266                         -> g16_msm_g1 / g16_msm_g2
 *   ark-poly EvaluationDomain::{fft,ifft}_in_place (+ coset) call sites r1cs_to_qap.rs:179-210 -> g16_ntt
 *   ark-ff Fp mul/add/sub/inverse (K1 parity hook)                                            -> g16_field_op
 *   forks/groth16/src/generator.rs:133-194 FixedBase::msm (key minting, "next" row f-2)       -> g16_fixed_base_g1 / _g2
 *   forks/groth16/src/verifier.rs:13-20   prepare_verifying_key ("next" row f-4)              -> g16_ctx_load_vk, g16_vk_alpha_beta
 *   forks/groth16/src/verifier.rs:25-39   Groth16::prepare_inputs                             -> g16_prepare_inputs
 *   forks/groth16/src/verifier.rs:44-76   verify_proof[_with_prepared_inputs], one call per proof -> g16_verify_batch (n proofs per call)
 *   ark-ec Pairing::pairing / multi_miller_loop + final_exponentiation call sites verifier.rs:17,48-62 -> g16_pairing
 *
 * Data layout at the boundary (all host pointers unless a function says "dev"):
 *   Fr / Fq element : 4 x uint64 little-endian limbs in Montgomery form (R = 2^256) -- byte-identical to arkworks'
 *                     in-memory `Fp<MontBackend<_,4>,4>` (BigInt<4>) and to 8 x uint32 limbs.
 *   G1 affine point : x || y, 8 x uint64.  Infinity is (0, 0) (arkworks keeps a separate bool: the shim repacks).
 *   G2 affine point : x.c0 || x.c1 || y.c0 || y.c1, 16 x uint64.  Infinity is all-zero.
 *   CSR matrix      : row_ptr uint64[nc+1], col uint32[nnz], val 4 x uint64[nnz] (Montgomery) -- the flattening of
 *                     ark-relations ConstraintMatrices {a,b,c}: Vec<Vec<(Fr, usize)>>.
 * All functions return 0 on success or a G16_ERR_* code; no proof bytes are written on error.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with G16_ERR_NO_DEVICE.
 * Thread safety: any thread may call; calls on one context are serialised by an internal mutex.
 */
#ifndef G16_B200_H
#define G16_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G16_OK 0
#define G16_ERR_DEGREE_TOO_LARGE 1 /* SynthesisError::PolynomialDegreeTooLarge (r1cs_to_qap.rs:156-157) */
#define G16_ERR_BAD_ARG 2          /* size mismatch / null pointer / not loaded */
#define G16_ERR_CUDA 3
#define G16_ERR_OOM 4
#define G16_ERR_NO_DEVICE 5
#define G16_ERR_VANISHING_ZERO 6 /* the reference's `.inverse().unwrap()` panic at r1cs_to_qap.rs:201-204 */

#define G16_REDUCTION_LIBSNARK 0
#define G16_REDUCTION_CIRCOM 1

#define G16_ENC_MONTGOMERY 0 /* arkworks in-memory limbs */
#define G16_ENC_CANONICAL 1  /* arkworks serialised (little-endian canonical integers, flag bits already cleared) */

/* field ids / ops for g16_field_op */
#define G16_FIELD_FR 0
#define G16_FIELD_FQ 1
#define G16_FIELD_FQ2 2
#define G16_OP_MUL 0
#define G16_OP_ADD 1
#define G16_OP_SUB 2
#define G16_OP_NEG 3
#define G16_OP_INV 4
#define G16_OP_TO_MONT 5
#define G16_OP_FROM_MONT 6
#define G16_OP_SQR 7
#define G16_OP_MUL_BCAST 8 /* out[i] = a[i] * b[0] */
#define G16_OP_ADD_BCAST 9 /* out[i] = a[i] + b[0] */

typedef struct g16_ctx g16_ctx;

/* Proving-key view: forks/groth16/src/data_structures.rs:101-118 (ProvingKey) + :31-44 (the vk fields the prover reads). */
typedef struct g16_pk_view {
    const uint64_t* a_query;    size_t a_len;    /* G1, m points (query[0] is the constant-1 wire) */
    const uint64_t* b_g1_query; size_t b_g1_len; /* G1, m points */
    const uint64_t* b_g2_query; size_t b_g2_len; /* G2, m points */
    const uint64_t* h_query;    size_t h_len;    /* G1, n-1 points */
    const uint64_t* l_query;    size_t l_len;    /* G1, m - num_instance points */
    const uint64_t* alpha_g1;                    /* vk.alpha_g1 */
    const uint64_t* beta_g1;
    const uint64_t* delta_g1;
    const uint64_t* beta_g2;                     /* vk.beta_g2 */
    const uint64_t* delta_g2;                    /* vk.delta_g2 */
    int encoding;                                /* G16_ENC_* of every coordinate above */
} g16_pk_view;

/* R1CS view: the three ConstraintMatrices flattened to CSR. */
typedef struct g16_r1cs_view {
    uint64_t num_constraints;
    uint64_t num_instance; /* l, including the constant 1 */
    uint64_t num_wires;    /* m = instance + witness */
    const uint64_t* row_ptr[3];
    const uint32_t* col[3];
    const uint64_t* val[3];
    int encoding; /* G16_ENC_* of val */
} g16_r1cs_view;

/* Proof{a, b, c} (data_structures.rs:7-14), affine, Montgomery limbs; *_inf != 0 marks the point at infinity. */
typedef struct g16_proof {
    uint64_t a[8];
    uint64_t b[16];
    uint64_t c[8];
    int32_t a_inf, b_inf, c_inf;
    int32_t _pad;
} g16_proof;

/* Per-rank partial results of the five sharded MSMs (XYZZ, Montgomery): h, l, a, s*a, r*b_g1 (16 x u64 each), b_g2
 * (32 x u64).  Every rank scales its own a / b_g1 partial by s / r while its other MSMs are still running, so that the
 * combine step only adds (s*g_a and r*g1_b of prover.rs:98,118 by linearity). */
#define G16_PARTIAL_U64 (5 * 16 + 32)
typedef struct g16_partial {
    uint64_t w[G16_PARTIAL_U64];
} g16_partial;

/* Device-side stage timings of the last g16_prove / g16_prove_shard on this context, milliseconds (CUDA events on the
 * context's streams).  Stage names follow the reference's start_timer! labels (prover.rs:35-36,62,93,103,115,123). */
typedef struct g16_timings {
    float h2d_ms;         /* witness upload */
    float witness_map_ms; /* "R1CS to QAP witness map" */
    float msm_h_ms, msm_l_ms, msm_a_ms, msm_b_g1_ms, msm_b_g2_ms;
    float assemble_ms;    /* scalar muls, "Finish C", normalisation */
    float total_ms;       /* "Groth16::Prover" */
    /* with option "kernel_events": duration of the bucket accumulation of each MSM (h, l, a, b_g1, b_g2) */
    float acc_ms[5];
    float assemble_kernel_ms; /* k_assemble_post alone (assemble_ms also covers the proof read-back) */
    float h_wait_ms;  /* end of the witness-map slot -> start of the h MSM on the main stream (staggered plan: the h scatter) */
    float h_start_ms; /* start of the call -> start of the h MSM */
} g16_timings;

/* ---- lifecycle ---------------------------------------------------------------------------------------------------- */
int g16_device_count(void);
/* Creates a context bound to CUDA device `device`.  main_stream may be NULL (library-owned stream) or a cudaStream_t the
 * caller owns (e.g. torch's current stream): all work of later calls is ordered on it. */
int g16_ctx_create(g16_ctx** out, int device, void* main_stream);
void g16_ctx_destroy(g16_ctx* ctx);
const char* g16_last_error(const g16_ctx* ctx); /* ctx may be NULL: last error of g16_ctx_create on this thread */
const char* g16_version(void);

/* Uploads the proving key.  With shard_count > 1 this context keeps only the contiguous point range
 * [rank*N/G, (rank+1)*N/G) of each query; the single points are kept by every rank.
 * precompute != 0 additionally stores 2^(c*j) multiples of every base so that all Pippenger windows share one bucket
 * set (trades HBM for the per-window reduction and the Horner tail). */
int g16_ctx_load_pk(g16_ctx* ctx, const g16_pk_view* pk, int shard_rank, int shard_count, int precompute);
/* Same with explicit ranges instead of the uniform split: h_range = [lo, hi) of h_query, z_range = [lo, hi) of the wire
 * index space of a_query[1..] / b_g1_query[1..] / b_g2_query[1..] (l_query follows a's range).  The ranks' ranges must
 * partition each query; an empty range is allowed.  Used by the staggered plan (g16_prove_shard_begin_dev below): the
 * rank that runs the witness map takes a smaller share of the wire MSMs. */
int g16_ctx_load_pk_ranges(g16_ctx* ctx, const g16_pk_view* pk, int shard_rank, int shard_count, const uint64_t h_range[2],
                           const uint64_t z_range[2], int precompute);
/* Uploads the R1CS matrices (once; proof-independent -- SURVEY 8f-1). */
int g16_ctx_load_r1cs(g16_ctx* ctx, const g16_r1cs_view* r1cs);

/* ---- the hot path ------------------------------------------------------------------------------------------------- */
/* create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment):
 * z = full_assignment (m Montgomery elements, z[0] = 1).  reduction selects the R1CSToQAP implementation. */
int g16_prove(g16_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], int reduction, g16_proof* out);
/* Same, witness already resident on the device (g16_upload_witness): no H2D inside. */
int g16_upload_witness(g16_ctx* ctx, const uint64_t* z);
int g16_prove_resident(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction, g16_proof* out);
/* Stream-ordered witness upload for the sharded path: queues the copy on the context's main stream and returns without
 * synchronising (z must stay valid, ideally page-locked, until the stream has passed it; the prove calls that follow on the same
 * context are ordered behind it).  shard_only != 0 copies only the slice of z this rank's wire MSMs read (z[1 + lo, 1 + hi) of its
 * a / b_g1 / b_g2 / l ranges): enough for every rank that does not run the witness map itself (g16_prove_shard_begin_dev with
 * run_witness_map = 0), which must not pay for the other 32 * m * (1 - 1/G) bytes. */
int g16_upload_witness_async(g16_ctx* ctx, const uint64_t* z, int shard_only);
/* The whole witness from DEVICE memory (m Montgomery elements; stream-ordered copy on the context's main stream): lets the
 * host glue of a sharded run bring z to the witness-map rank over several PCIe links + NVLink (every rank uploads 1/G of it,
 * one gather) instead of over that rank's own link alone. */
int g16_upload_witness_dev(g16_ctx* ctx, const void* z_dev);
/* cudaMemcpyAsync(host -> device) on a raw cudaStream_t (the current device of the calling thread): for host glue that
 * addresses page-locked memory by pointer. */
int g16_memcpy_h2d_async(void* dst_dev, const void* src_host, size_t bytes, void* stream);

/* MSM-sharded proving: every rank calls g16_prove_shard on its context (loaded with its shard) with the same (r, s), the
 * G partials are gathered (one small NCCL gather by the host glue) and rank 0 calls g16_prove_combine. */
int g16_prove_shard(g16_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], int reduction, g16_partial* out);
int g16_prove_combine(g16_ctx* ctx, const g16_partial* partials, int count, const uint64_t r[4], const uint64_t s[4],
                      g16_proof* out);
/* Device pointer + byte size of this context's partial buffer, for a device-side NCCL gather. */
int g16_partial_dev(g16_ctx* ctx, void** dev_ptr, size_t* bytes);
int g16_prove_shard_dev(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction); /* witness resident; result left in the device partial */
int g16_copy_partial_dev(g16_ctx* ctx, void* dst_dev);         /* stream-ordered D2D copy of the partial (e.g. into an NCCL buffer) */
/* g16_prove_shard_dev in two halves, for the staggered multi-GPU plan: only ONE rank runs the witness map
 * (run_witness_map != 0) while the others spend that time on their larger share of the wire MSMs; the host glue then
 * scatters h (g16_copy_h_dev into the collective's buffer, one NCCL scatter on the context's main stream) and every rank
 * finishes with the h MSM over its h range, reading the coefficients h[h_first, h_first + h_count) from h_dev (Montgomery
 * Fr; must cover the rank's h range; NULL = this context's own witness-map output).  begin: forks the wire MSMs (+ witness map); finish: h MSM + join; the partial is then in the device buffer. */
int g16_prove_shard_begin_dev(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4], int reduction, int run_witness_map);
int g16_prove_shard_finish_dev(g16_ctx* ctx, const void* h_dev, size_t h_first, size_t h_count); /* h_dev[i] = h[h_first + i], i < h_count */
/* The LibsnarkReduction witness map in parts, for host glue that spreads it over several GPUs (the a, b and c pipelines are
 * independent until the last transform): `parts` is a mask of G16_WM_PART_*; each part reads / leaves its vector in the context
 * (a: A, FINAL; b: B; c: C), FINAL needs a, b and c and leaves h in a (g16_copy_h_dev).  The whole witness must be on the device.
 * A | B | C | FINAL in one call equals g16_witness_map's device work bit for bit.  g16_wm_vector_copy_dev moves one of the three
 * n-element vectors between the context and caller-owned device memory (to_ctx != 0: into the context), stream-ordered. */
#define G16_WM_PART_A 1
#define G16_WM_PART_B 2
#define G16_WM_PART_C 4
#define G16_WM_PART_FINAL 8
int g16_witness_map_part_dev(g16_ctx* ctx, int parts);
int g16_wm_vector_copy_dev(g16_ctx* ctx, int which /* 0 = a, 1 = b, 2 = c */, void* ext_dev, size_t capacity_elems, int to_ctx);
int g16_copy_h_dev(g16_ctx* ctx, void* dst_dev, size_t capacity_elems); /* stream-ordered D2D copy of h (n elements) */
/* Optional, rank 0: starts the (r, s, pk)-only scalar multiplications on a side stream so that they overlap the shard work;
 * a later g16_prove_combine[_dev] with the same (r, s) joins them instead of running them serially. */
int g16_prove_prepare(g16_ctx* ctx, const uint64_t r[4], const uint64_t s[4]);
int g16_prove_combine_dev(g16_ctx* ctx, const void* dev_partials, int count, const uint64_t r[4], const uint64_t s[4],
                          g16_proof* out);

/* witness_map_from_matrices: h_out receives n Montgomery elements (n = domain size, returned in *n_out). */
int g16_witness_map(g16_ctx* ctx, const uint64_t* z, int reduction, uint64_t* h_out, size_t h_capacity, size_t* n_out);
int g16_domain_size(g16_ctx* ctx, size_t* n_out);
int g16_get_timings(g16_ctx* ctx, g16_timings* out);
/* Options: "serialize" = 1 runs every stage on the main stream (no overlap; for per-kernel timing),
 * "kernel_events" = 1 brackets the bucket accumulation of every MSM (batched-affine levels + XYZZ tail) with CUDA events
 *   (fills g16_timings.acc_ms); = 2 brackets only the first level's k_ba_add launch (the dominant kernel),
 * "window_bits" = c forces the Pippenger window of bases loaded afterwards (0 = automatic: round(log2 n) - 3 with window tables
 *   from 2^17 points on, log2 n - 1 below, log2 n - 5 capped at 16 without tables),
 * "ba_levels" = L runs L pairwise batched-affine levels before the XYZZ tail for bases loaded afterwards (-1 = default 5,
 *   0 = XYZZ only), "share_digits" = 0 disables the reuse of one digit stage by a/l and b_g1/b_g2,
 * "split_chains" = 0 queues the MSM that reuses a digit stage behind the one that built it (default 1: beside it),
 * "wm_priority" = 1 runs the witness map and the h MSM on a high-priority stream, = 2 only the witness map (default 0),
 * "wm_first" = 1 / 0 starts the z-only MSM chains only after the witness map, which then runs alone / beside it (default -1:
 *   after it for domains of 2^20 and more),
 * "ntt_radix4" = 0 / 1 forces the radix-2 / radix-4 transform passes (default -1: radix-4 only when no MSM runs beside),
 * "spmv_sell" = 0 selects the row-per-thread CSR kernel instead of the sliced-ELL one (default 1),
 * "ntt_batch" = 0 / 1 forces one launch per transform and pass / batched launches for the witness map (default -1: batched
 *   only when no MSM runs beside the transforms),
 * "graph" = 0 queues every launch of a proof eagerly instead of replaying the captured launch sequences as CUDA graphs
 *   (default 1; the first run of a sequence is always eager, the second is captured),
 * "asm_tables" = 0 computes the (r, s)-only points of the assembly with one lane per scalar multiplication instead of the
 *   per-key fixed-base tables (default 1; matters for small circuits only: 4.4 ms -> 0.3 ms of latency),
 * "verify_occupancy" = 8 / 12 / 16 selects the k_verify build for that many resident warps per SM (255 / 168 / 128 registers).
 * No option changes a result bit.  Unknown keys return G16_ERR_BAD_ARG. */
int g16_set_option(g16_ctx* ctx, const char* key, int value);
/* CUDA-graph replay counters: out[0] = graph launches so far, out[1] = launch sequences captured, out[2] = captures that failed
 * (after one the context queues its launches eagerly).  g16_launch_count counts the kernels a replay stands for. */
int g16_graph_stats(const g16_ctx* ctx, uint64_t out[3]);

/* ---- building blocks (parity hooks and the synthetic sweep) ------------------------------------------------------- */
/* Sigma scalars[i] * points[i]; scalars Montgomery Fr; result affine Montgomery (+ infinity flag). */
int g16_msm_g1(g16_ctx* ctx, const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t out[8], int* out_inf);
int g16_msm_g2(g16_ctx* ctx, const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t out[16], int* out_inf);
/* Resident variant: bases stay on the device in `slot` (0..7); scalars_dev is a device pointer.  window_bits = 0 picks
 * the default; precompute as in g16_ctx_load_pk. */
int g16_msm_set_bases(g16_ctx* ctx, int slot, int group /*1|2*/, const uint64_t* points, size_t n, int window_bits,
                      int precompute);
int g16_msm_set_bases_dev(g16_ctx* ctx, int slot, int group, const void* points_dev, size_t n, int window_bits,
                          int precompute);
int g16_msm_run_dev(g16_ctx* ctx, int slot, const void* scalars_dev, size_t n, uint64_t* out, int* out_inf);
/* The window (bits per signed digit) the library picks for n bases when window_bits = 0: with window tables (precompute)
 * round(log2 n) - 3 from 2^17 points on and log2 n - 1 below (one bucket set serves all windows, so a smaller window buys
 * fuller buckets); without tables log2 n - 5, at most 16.  The tables hold ceil(254 / c) copies of the bases.  Pure function:
 * needs no context and no device. */
int g16_msm_window_bits(size_t n, int precompute);
/* Point-range sharding of a stand-alone MSM (the pattern of g16_prove_shard / g16_prove_combine for one sum): every
 * rank runs g16_msm_run_dev over its contiguous share of the pairs with out = NULL, copies its partial sum (XYZZ: 16 u64 words
 * for G1, 32 for G2) to dst_dev with g16_msm_copy_result_dev -- stream-ordered, so that one all_gather on the same stream can
 * follow -- and one rank adds the `count` gathered partials and normalises with g16_msm_combine_dev. */
int g16_msm_copy_result_dev(g16_ctx* ctx, int slot, void* dst_dev);
int g16_msm_combine_dev(g16_ctx* ctx, int group, const void* partials_dev, int count, uint64_t* out, int* out_inf);

/* In-place NTT over Fr of size 2^log_n, natural order in and out (arkworks semantics): inverse includes 1/n; coset
 * applies the shift g = Fr::GENERATOR = 5 (fft: scale then transform; ifft: transform then unscale). */
int g16_ntt(g16_ctx* ctx, uint64_t* data, unsigned log_n, int inverse, int coset);
int g16_ntt_dev(g16_ctx* ctx, void* data_dev, unsigned log_n, int inverse, int coset);

/* Element-wise field arithmetic on n elements (Fq2: 8 x u64 per element). */
int g16_field_op(g16_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);

/* k_i * G for the canonical generators (generator.rs:34-35); scalars Montgomery Fr; outputs affine Montgomery. */
int g16_fixed_base_g1(g16_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out_points);
int g16_fixed_base_g2(g16_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out_points);
int g16_fixed_base_g1_dev(g16_ctx* ctx, const void* scalars_dev, size_t n, void* out_points_dev);
int g16_fixed_base_g2_dev(g16_ctx* ctx, const void* scalars_dev, size_t n, void* out_points_dev);

/* out[i] = scale * base^i, i < n (Montgomery Fr) -- powers of tau for the key generator (r1cs_to_qap.rs:215-225). */
int g16_pow_table(g16_ctx* ctx, const uint64_t base[4], const uint64_t scale[4], size_t n, uint64_t* out);

/* Sparse R1CS evaluation a = A z, b = B z, c = C z over the loaded matrices (evaluate_constraint,
 * r1cs_to_qap.rs:16-45); outputs nc Montgomery elements each (any may be NULL). */
int g16_r1cs_eval(g16_ctx* ctx, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz);

/* ---- verification ("next" row f-4; forks/groth16/src/verifier.rs) --------------------------------------------------- */
/* VerifyingKey view (data_structures.rs:31-44; delta_g1 is not read by the verifier).  gamma_abc_g1: gamma_abc_len G1 points,
 * entry 0 pairs with the constant-1 wire. */
typedef struct g16_vk_view {
    const uint64_t* alpha_g1;     /* G1 */
    const uint64_t* beta_g2;      /* G2 */
    const uint64_t* gamma_g2;     /* G2 */
    const uint64_t* delta_g2;     /* G2 */
    const uint64_t* gamma_abc_g1; size_t gamma_abc_len;
    int encoding;                 /* G16_ENC_* of every coordinate above */
} g16_vk_view;

#define G16_VERDICT_REJECT 0
#define G16_VERDICT_ACCEPT 1
#define G16_VERDICT_UNEXPECTED_IDENTITY 2 /* final_exponentiation returned None (SynthesisError::UnexpectedIdentity, verifier.rs:62) */

/* prepare_verifying_key (verifier.rs:13-20): uploads the key and computes on the device e(alpha_g1, beta_g2), the line
 * coefficients of -gamma_g2 and -delta_g2 (G2Prepared) and fixed-base window tables of gamma_abc_g1[1..]. */
int g16_ctx_load_vk(g16_ctx* ctx, const g16_vk_view* vk);
/* PreparedVerifyingKey.alpha_g1_beta_g2: an Fq12 as 12 Montgomery Fq in ark-serialize order (c0.c0.c0 ... c1.c2.c1). */
int g16_vk_alpha_beta(g16_ctx* ctx, uint64_t out[48]);
/* prepare_inputs (verifier.rs:25-39) for n instances: public_inputs holds n * (gamma_abc_len - 1) Montgomery Fr, instance-major
 * (the constant-1 wire is NOT passed, as in the reference); out receives n affine G1 points (Montgomery, infinity = (0, 0)). */
int g16_prepare_inputs(g16_ctx* ctx, const uint64_t* public_inputs, size_t n, uint64_t* out_points);
/* verify_proof (verifier.rs:69-76) for n (proof, instance) pairs under the loaded key, one device thread per proof:
 * verdict[i] = G16_VERDICT_* of pair i -- exactly the reference's verdict for that pair (no random linear combination across
 * proofs, so one bad proof cannot hide and none can be blamed wrongly).  As in the reference, points are not checked for
 * curve or subgroup membership. */
int g16_verify_batch(g16_ctx* ctx, const g16_proof* proofs, const uint64_t* public_inputs, size_t n, uint8_t* verdict);
/* verify_proof_with_prepared_inputs (verifier.rs:44-65): as g16_verify_batch, with the n prepared-input points (affine G1,
 * Montgomery, 8 x u64 each, (0, 0) = infinity) supplied by the caller instead of the public inputs. */
int g16_verify_batch_prepared(g16_ctx* ctx, const g16_proof* proofs, const uint64_t* prepared_inputs, size_t n, uint8_t* verdict);
/* Same with everything resident on the device (g16_proof[n], Fr[n * (gamma_abc_len - 1)], uint8_t[n]); stream-ordered. */
int g16_verify_batch_dev(g16_ctx* ctx, const void* proofs_dev, const void* public_inputs_dev, size_t n, void* verdict_dev);
/* E::pairing(p_i, q_i).0 for n pairs (parity hook): gt_out receives n Fq12 (48 x u64 each, layout as g16_vk_alpha_beta).
 * A pair holding a point at infinity yields one. */
int g16_pairing(g16_ctx* ctx, const uint64_t* g1_points, const uint64_t* g2_points, size_t n, uint64_t* gt_out);

/* ---- device memory + micro-benchmarks ------------------------------------------------------------------------------ */
/* Page-locked host memory for the per-proof witness: a witness built in such a buffer is uploaded by DMA at PCIe / C2C speed
 * (46 MB in 0.7 ms); from pageable memory the driver stages it through its own bounce buffers (~6 ms).  Free with g16_host_free. */
int g16_host_alloc(g16_ctx* ctx, size_t bytes, void** host_ptr);
int g16_host_free(g16_ctx* ctx, void* host_ptr);
int g16_dev_alloc(g16_ctx* ctx, size_t bytes, void** dev_ptr);
int g16_dev_free(g16_ctx* ctx, void* dev_ptr);
int g16_dev_upload(g16_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int g16_dev_download(g16_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
int g16_sync(g16_ctx* ctx);
/* Integer-pipe peak probes: returns achieved giga-operations per second of (0) IMAD 32-bit, (1) IMAD.WIDE.U32 chains,
 * (2) Fr Montgomery multiplications, (3) Fq Montgomery multiplications -- the measured denominators of the integer
 * roofline (MEASURED_PEAKS.json carries none). */
int g16_bench_int_pipe(g16_ctx* ctx, int which, double* gops_out);
/* Geometry of the last run of one of the proof's MSMs (which: 0 = h, 1 = l, 2 = a, 3 = b_g1, 4 = b_g2), for the bench's
 * per-launch work figures: out[0] points, [1] window bits c, [2] windows, [3] batched-affine levels, [4] buckets,
 * [5] 1 if the digit stage is shared with another MSM, [6] XYZZ tail tasks, [8..8+levels) points left after each level
 * (= additions-or-copies the level's kernels process).  capacity >= 16.  Synchronises. */
int g16_get_msm_stats(g16_ctx* ctx, int which, uint64_t* out, int capacity);
/* Number of kernels this library has launched on this context since creation (for the bench's gpu_launches claim). */
uint64_t g16_launch_count(const g16_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* G16_B200_H */
