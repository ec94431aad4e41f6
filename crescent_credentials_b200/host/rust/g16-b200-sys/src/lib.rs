//! FFI binding + safe wrapper for libg16b200.so (include/g16_b200.h).
//!
//! forks/groth16 is `#![forbid(unsafe_code)]` (forks/groth16/src/lib.rs:13), so the `extern "C"` block and every `unsafe`
//! call live here; forks/groth16/src/prover.rs calls only the safe `B200Prover` below, behind the cargo feature `b200`.
//! arkworks' `Fp<MontBackend<_,4>,4>` is not `repr(C)` and `Affine<P>` carries a separate `infinity: bool`, so points
//! and scalars are *repacked* into dense `u64` arrays (once per proving key, once per proof for the witness) rather than
//! transmuted.  NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the image).
use ark_bn254::{Bn254, Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ec::AffineRepr;
use ark_ff::{BigInt, PrimeField, Zero};
use ark_relations::r1cs::{ConstraintMatrices, SynthesisError};
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct g16_pk_view {
    a_query: *const u64, a_len: usize,
    b_g1_query: *const u64, b_g1_len: usize,
    b_g2_query: *const u64, b_g2_len: usize,
    h_query: *const u64, h_len: usize,
    l_query: *const u64, l_len: usize,
    alpha_g1: *const u64, beta_g1: *const u64, delta_g1: *const u64,
    beta_g2: *const u64, delta_g2: *const u64,
    encoding: c_int,
}
#[repr(C)]
pub struct g16_r1cs_view {
    num_constraints: u64, num_instance: u64, num_wires: u64,
    row_ptr: [*const u64; 3], col: [*const u32; 3], val: [*const u64; 3],
    encoding: c_int,
}
#[repr(C)]
#[derive(Default)]
pub struct g16_proof { a: [u64; 8], b: [u64; 16], c: [u64; 8], a_inf: i32, b_inf: i32, c_inf: i32, _pad: i32 }

extern "C" {
    fn g16_ctx_create(out: *mut *mut c_void, device: c_int, main_stream: *mut c_void) -> c_int;
    fn g16_ctx_destroy(ctx: *mut c_void);
    fn g16_last_error(ctx: *const c_void) -> *const c_char;
    fn g16_ctx_load_pk(ctx: *mut c_void, pk: *const g16_pk_view, rank: c_int, count: c_int, precompute: c_int) -> c_int;
    fn g16_ctx_load_r1cs(ctx: *mut c_void, r1cs: *const g16_r1cs_view) -> c_int;
    fn g16_prove(ctx: *mut c_void, z: *const u64, r: *const u64, s: *const u64, reduction: c_int, out: *mut g16_proof) -> c_int;
    fn g16_ctx_load_vk(ctx: *mut c_void, vk: *const g16_vk_view) -> c_int;
    fn g16_vk_alpha_beta(ctx: *mut c_void, out: *mut u64) -> c_int;
    fn g16_verify_batch(ctx: *mut c_void, proofs: *const g16_proof, public_inputs: *const u64, n: usize, verdict: *mut u8) -> c_int;
}

#[repr(C)]
pub struct g16_vk_view {
    alpha_g1: *const u64, beta_g2: *const u64, gamma_g2: *const u64, delta_g2: *const u64,
    gamma_abc_g1: *const u64, gamma_abc_len: usize, encoding: c_int,
}

const G16_ERR_DEGREE_TOO_LARGE: c_int = 1;

// Montgomery limbs exactly as arkworks holds them: Fp(BigInt([u64; 4]), _)
fn fq_limbs(x: &Fq) -> [u64; 4] { (x.0).0 }
fn fr_limbs(x: &Fr) -> [u64; 4] { (x.0).0 }
fn pack_g1(p: &G1Affine, out: &mut Vec<u64>) {
    if p.is_zero() { out.extend_from_slice(&[0u64; 8]); return; }   // infinity -> (0, 0)
    out.extend_from_slice(&fq_limbs(&p.x)); out.extend_from_slice(&fq_limbs(&p.y));
}
fn pack_g2(p: &G2Affine, out: &mut Vec<u64>) {
    if p.is_zero() { out.extend_from_slice(&[0u64; 16]); return; }
    for c in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1] { out.extend_from_slice(&fq_limbs(c)); }
}
fn g1_from(l: &[u64; 8], inf: bool) -> G1Affine {
    if inf { return G1Affine::identity(); }
    let f = |w: &[u64]| Fq::new_unchecked(BigInt([w[0], w[1], w[2], w[3]]));   // already Montgomery
    G1Affine::new_unchecked(f(&l[0..4]), f(&l[4..8]))
}
fn g2_from(l: &[u64; 16], inf: bool) -> G2Affine {
    if inf { return G2Affine::identity(); }
    let f = |w: &[u64]| Fq::new_unchecked(BigInt([w[0], w[1], w[2], w[3]]));
    G2Affine::new_unchecked(Fq2::new(f(&l[0..4]), f(&l[4..8])), Fq2::new(f(&l[8..12]), f(&l[12..16])))
}

/// Persistent GPU prover: owns device copies of one ProvingKey and one set of constraint matrices.
pub struct B200Prover { ctx: *mut c_void, wires: usize }
unsafe impl Send for B200Prover {}   // the C context serialises calls with its own mutex; no thread affinity

impl B200Prover {
    pub fn new(device: i32) -> Self {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { g16_ctx_create(&mut ctx, device, std::ptr::null_mut()) };
        if rc != 0 { panic!("g16_ctx_create: {}", Self::err(std::ptr::null())); }
        Self { ctx, wires: 0 }
    }
    fn err(ctx: *const c_void) -> String { unsafe { CStr::from_ptr(g16_last_error(ctx)).to_string_lossy().into_owned() } }

    /// `pk` as in forks/groth16/src/data_structures.rs:101-118 (generic over E = Bn254 here).
    pub fn load_pk(&mut self, alpha_g1: &G1Affine, beta_g1: &G1Affine, delta_g1: &G1Affine, beta_g2: &G2Affine,
                   delta_g2: &G2Affine, a: &[G1Affine], b1: &[G1Affine], b2: &[G2Affine], h: &[G1Affine], l: &[G1Affine]) {
        let pack1 = |v: &[G1Affine]| { let mut o = Vec::with_capacity(v.len() * 8); v.iter().for_each(|p| pack_g1(p, &mut o)); o };
        let (qa, qb1, qh, ql) = (pack1(a), pack1(b1), pack1(h), pack1(l));
        let mut qb2 = Vec::with_capacity(b2.len() * 16); b2.iter().for_each(|p| pack_g2(p, &mut qb2));
        let (sa, sb, sd) = (pack1(&[*alpha_g1]), pack1(&[*beta_g1]), pack1(&[*delta_g1]));
        let (mut tb, mut td) = (vec![], vec![]); pack_g2(beta_g2, &mut tb); pack_g2(delta_g2, &mut td);
        let view = g16_pk_view { a_query: qa.as_ptr(), a_len: a.len(), b_g1_query: qb1.as_ptr(), b_g1_len: b1.len(),
            b_g2_query: qb2.as_ptr(), b_g2_len: b2.len(), h_query: qh.as_ptr(), h_len: h.len(), l_query: ql.as_ptr(), l_len: l.len(),
            alpha_g1: sa.as_ptr(), beta_g1: sb.as_ptr(), delta_g1: sd.as_ptr(), beta_g2: tb.as_ptr(), delta_g2: td.as_ptr(), encoding: 0 };
        let rc = unsafe { g16_ctx_load_pk(self.ctx, &view, 0, 1, 1) };
        if rc != 0 { panic!("g16_ctx_load_pk: {}", Self::err(self.ctx)); }   // device failures panic like today's unwraps
    }

    /// Flattens ConstraintMatrices {a, b, c}: Vec<Vec<(Fr, usize)>> (forks/circom-compat/src/zkey.rs:181-193 shape) to CSR.
    pub fn load_matrices(&mut self, m: &ConstraintMatrices<Fr>) -> Result<(), SynthesisError> {
        let flat = |rows: &Vec<Vec<(Fr, usize)>>| {
            let (mut ptr, mut col, mut val) = (vec![0u64], Vec::new(), Vec::new());
            for row in rows { for (c, j) in row { col.push(*j as u32); val.extend_from_slice(&fr_limbs(c)); } ptr.push(col.len() as u64); }
            (ptr, col, val)
        };
        let (a, b, c) = (flat(&m.a), flat(&m.b), flat(&m.c));
        self.wires = m.num_instance_variables + m.num_witness_variables;
        let view = g16_r1cs_view { num_constraints: m.num_constraints as u64, num_instance: m.num_instance_variables as u64,
            num_wires: self.wires as u64, row_ptr: [a.0.as_ptr(), b.0.as_ptr(), c.0.as_ptr()],
            col: [a.1.as_ptr(), b.1.as_ptr(), c.1.as_ptr()], val: [a.2.as_ptr(), b.2.as_ptr(), c.2.as_ptr()], encoding: 0 };
        match unsafe { g16_ctx_load_r1cs(self.ctx, &view) } {
            0 => Ok(()),
            G16_ERR_DEGREE_TOO_LARGE => Err(SynthesisError::PolynomialDegreeTooLarge),   // r1cs_to_qap.rs:156-157
            _ => panic!("g16_ctx_load_r1cs: {}", Self::err(self.ctx)),
        }
    }

    /// create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment)
    pub fn prove(&self, r: Fr, s: Fr, full_assignment: &[Fr]) -> Result<(G1Affine, G2Affine, G1Affine), SynthesisError> {
        assert_eq!(full_assignment.len(), self.wires);
        let mut z = Vec::with_capacity(full_assignment.len() * 4);
        full_assignment.iter().for_each(|x| z.extend_from_slice(&fr_limbs(x)));
        let mut out = g16_proof::default();
        match unsafe { g16_prove(self.ctx, z.as_ptr(), fr_limbs(&r).as_ptr(), fr_limbs(&s).as_ptr(), 0, &mut out) } {
            0 => Ok((g1_from(&out.a, out.a_inf != 0), g2_from(&out.b, out.b_inf != 0), g1_from(&out.c, out.c_inf != 0))),
            G16_ERR_DEGREE_TOO_LARGE => Err(SynthesisError::PolynomialDegreeTooLarge),
            _ => panic!("g16_prove: {}", Self::err(self.ctx)),
        }
    }
}
impl Drop for B200Prover { fn drop(&mut self) { unsafe { g16_ctx_destroy(self.ctx) } } }

// ---- prover_for: the per-(pk, matrices) cache the fork's prover.rs calls ---------------------------
// `create_proof_with_reduction_and_matrices` receives `&ProvingKey` and `&ConstraintMatrices` on every call and keeps nothing
// between calls (SURVEY 8b "Ownership"); the device copies must outlive the call.  Identity cannot be the address (a key that
// was dropped and reloaded may land elsewhere, and another key may land on the old address), so the cache is keyed on a CONTENT
// fingerprint: lengths plus a 64-bit FNV-1a over the first / last / strided sample of every query and of the matrices'
// coefficient words -- cheap (a few thousand words per call) and collision-safe for the handful of keys a process ever holds.
// One entry per fingerprint; a Mutex per entry serialises proofs on that context (the C side locks as well).
use std::collections::HashMap;
use std::sync::{Arc, Mutex, OnceLock};

#[derive(Clone, Copy, PartialEq, Eq, Hash)]
pub struct Fingerprint(u64, usize, usize);

fn fnv(h: &mut u64, words: &[u64]) { for w in words { *h = (*h ^ *w).wrapping_mul(0x100000001b3); } }
fn sample_g1(h: &mut u64, v: &[G1Affine]) {
    let n = v.len();
    let step = (n / 64).max(1);
    for i in (0..n).step_by(step).chain(n.saturating_sub(1)..n) { let mut o = Vec::with_capacity(8); pack_g1(&v[i], &mut o); fnv(h, &o); }
}

/// Fingerprint of (pk, matrices) -- the cache key of `prover_for`.
pub fn fingerprint(a: &[G1Affine], b1: &[G1Affine], h_query: &[G1Affine], l: &[G1Affine], delta_g1: &G1Affine,
                   m: &ConstraintMatrices<Fr>) -> Fingerprint {
    let mut h = 0xcbf29ce484222325u64;
    sample_g1(&mut h, a); sample_g1(&mut h, b1); sample_g1(&mut h, h_query); sample_g1(&mut h, l); sample_g1(&mut h, &[*delta_g1]);
    fnv(&mut h, &[m.num_constraints as u64, m.num_instance_variables as u64, m.num_witness_variables as u64,
                  m.a_num_non_zero as u64, m.b_num_non_zero as u64, m.c_num_non_zero as u64]);
    for rows in [&m.a, &m.b, &m.c] {
        let step = (rows.len() / 256).max(1);
        for row in rows.iter().step_by(step) { for (c, j) in row { fnv(&mut h, &fr_limbs(c)); fnv(&mut h, &[*j as u64]); } }
    }
    Fingerprint(h, a.len(), h_query.len())
}

static PROVERS: OnceLock<Mutex<HashMap<Fingerprint, Arc<Mutex<B200Prover>>>>> = OnceLock::new();

/// The persistent prover for this (pk, matrices): created, loaded (key upload + window tables + CSR upload: seconds, once) and
/// cached on first use, reused by every later proof.  `device` < 0 picks device 0.  Called by forks/groth16/src/b200.rs.
#[allow(clippy::too_many_arguments)]
pub fn prover_for(device: i32, alpha_g1: &G1Affine, beta_g1: &G1Affine, delta_g1: &G1Affine, beta_g2: &G2Affine, delta_g2: &G2Affine,
                  a: &[G1Affine], b1: &[G1Affine], b2: &[G2Affine], h_query: &[G1Affine], l: &[G1Affine],
                  m: &ConstraintMatrices<Fr>) -> Result<Arc<Mutex<B200Prover>>, SynthesisError> {
    let key = fingerprint(a, b1, h_query, l, delta_g1, m);
    let map = PROVERS.get_or_init(|| Mutex::new(HashMap::new()));
    let mut guard = map.lock().unwrap();
    if let Some(p) = guard.get(&key) { return Ok(p.clone()); }
    let mut p = B200Prover::new(device.max(0));
    p.load_matrices(m)?;                                   // PolynomialDegreeTooLarge surfaces here (r1cs_to_qap.rs:156-157)
    p.load_pk(alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2, a, b1, b2, h_query, l);
    let p = Arc::new(Mutex::new(p));
    guard.insert(key, p.clone());
    Ok(p)
}

/// Drops every cached context (frees the device copies); e.g. between circuits in a long-lived service.
pub fn clear_provers() { if let Some(m) = PROVERS.get() { m.lock().unwrap().clear(); } }

/// Persistent GPU verifier for one VerifyingKey (forks/groth16/src/verifier.rs): `verify_proofs` checks n (proof, inputs)
/// pairs per call, one device thread per proof, and returns the reference's verdict for each.
pub struct B200Verifier { ctx: *mut c_void, inputs: usize }
unsafe impl Send for B200Verifier {}

impl B200Verifier {
    /// prepare_verifying_key (verifier.rs:13-20) on the device.  The key's fields are passed one by one (data_structures.rs:
    /// 31-44): this crate cannot name the fork's `VerifyingKey` without a dependency cycle (the fork depends on it).
    pub fn new(device: i32, alpha_g1: &G1Affine, beta_g2: &G2Affine, gamma_g2: &G2Affine, delta_g2: &G2Affine,
               gamma_abc_g1: &[G1Affine]) -> Self {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { g16_ctx_create(&mut ctx, device, std::ptr::null_mut()) };
        if rc != 0 { panic!("g16_ctx_create: {}", B200Prover::err(std::ptr::null())); }
        let (mut a, mut b, mut g, mut d, mut abc) = (vec![], vec![], vec![], vec![], vec![]);
        pack_g1(alpha_g1, &mut a); pack_g2(beta_g2, &mut b); pack_g2(gamma_g2, &mut g); pack_g2(delta_g2, &mut d);
        gamma_abc_g1.iter().for_each(|p| pack_g1(p, &mut abc));
        let view = g16_vk_view { alpha_g1: a.as_ptr(), beta_g2: b.as_ptr(), gamma_g2: g.as_ptr(), delta_g2: d.as_ptr(),
            gamma_abc_g1: abc.as_ptr(), gamma_abc_len: gamma_abc_g1.len(), encoding: 0 };
        if unsafe { g16_ctx_load_vk(ctx, &view) } != 0 { panic!("g16_ctx_load_vk: {}", B200Prover::err(ctx)); }
        Self { ctx, inputs: gamma_abc_g1.len() - 1 }
    }
    /// PreparedVerifyingKey.alpha_g1_beta_g2 as 12 Montgomery Fq (c0.c0.c0 ... c1.c2.c1).
    pub fn alpha_g1_beta_g2(&self) -> [u64; 48] {
        let mut out = [0u64; 48];
        if unsafe { g16_vk_alpha_beta(self.ctx, out.as_mut_ptr()) } != 0 { panic!("g16_vk_alpha_beta: {}", B200Prover::err(self.ctx)); }
        out
    }
    /// Groth16::verify_proof (verifier.rs:69-76) for every pair; a proof is its (a, b, c) (data_structures.rs:7-14).
    pub fn verify_proofs(&self, proofs: &[(G1Affine, G2Affine, G1Affine)], public_inputs: &[Vec<Fr>]) -> Result<Vec<bool>, SynthesisError> {
        assert_eq!(proofs.len(), public_inputs.len());
        let mut ps = Vec::with_capacity(proofs.len());
        let mut xs = Vec::with_capacity(proofs.len() * self.inputs * 4);
        for (p, x) in proofs.iter().zip(public_inputs) {
            if x.len() != self.inputs { return Err(SynthesisError::MalformedVerifyingKey); }          // verifier.rs:29-31
            let (mut a, mut b, mut c) = (vec![], vec![], vec![]);
            pack_g1(&p.0, &mut a); pack_g2(&p.1, &mut b); pack_g1(&p.2, &mut c);
            let mut q = g16_proof::default();
            q.a.copy_from_slice(&a); q.b.copy_from_slice(&b); q.c.copy_from_slice(&c);
            q.a_inf = p.0.is_zero() as i32; q.b_inf = p.1.is_zero() as i32; q.c_inf = p.2.is_zero() as i32;
            ps.push(q);
            x.iter().for_each(|v| xs.extend_from_slice(&fr_limbs(v)));
        }
        let mut verdict = vec![0u8; proofs.len()];
        if unsafe { g16_verify_batch(self.ctx, ps.as_ptr(), xs.as_ptr(), ps.len(), verdict.as_mut_ptr()) } != 0 {
            panic!("g16_verify_batch: {}", B200Prover::err(self.ctx));
        }
        if verdict.iter().any(|&v| v == 2) { return Err(SynthesisError::UnexpectedIdentity); }       // verifier.rs:62
        Ok(verdict.iter().map(|&v| v == 1).collect())
    }
}
impl Drop for B200Verifier { fn drop(&mut self) { unsafe { g16_ctx_destroy(self.ctx) } } }
