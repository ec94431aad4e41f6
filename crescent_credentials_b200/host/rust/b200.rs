//! forks/groth16/src/b200.rs -- the module `prover.rs` calls behind the cargo feature `b200`.
//! Safe code only (`forks/groth16` is `#![forbid(unsafe_code)]`, lib.rs:13): everything unsafe lives in the g16-b200-sys crate.
//! Add `#[cfg(feature = "b200")] pub mod b200;` to forks/groth16/src/lib.rs.  SOURCE ONLY: no Rust toolchain in the build image.
use crate::{Proof, ProvingKey};
use ark_bn254::{Bn254, Fr};
use ark_relations::r1cs::{ConstraintMatrices, Result as R1CSResult};
use g16_b200_sys::{prover_for as sys_prover_for, B200Prover};
use std::sync::{Arc, Mutex};

/// The persistent GPU prover for this (pk, matrices): device copies are created on first use and reused by every later proof
/// (`create_proof_with_reduction_and_matrices` itself keeps nothing between calls).  Device from `G16_DEVICE` (default 0).
pub fn prover_for(pk: &ProvingKey<Bn254>, matrices: &ConstraintMatrices<Fr>) -> R1CSResult<Arc<Mutex<B200Prover>>> {
    let device = std::env::var("G16_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
    sys_prover_for(device, &pk.vk.alpha_g1, &pk.beta_g1, &pk.delta_g1, &pk.vk.beta_g2, &pk.vk.delta_g2, &pk.a_query,
                   &pk.b_g1_query, &pk.b_g2_query, &pk.h_query, &pk.l_query, matrices)
}

/// Body of `create_proof_with_reduction_and_matrices` (prover.rs:26-51) for E = Bn254 under the feature:
/// ```ignore
/// #[cfg(feature = "b200")]
/// { return crate::b200::prove(pk, r, s, matrices, num_inputs, num_constraints, full_assignment); }
/// ```
pub fn prove(pk: &ProvingKey<Bn254>, r: Fr, s: Fr, matrices: &ConstraintMatrices<Fr>, num_inputs: usize, num_constraints: usize,
             full_assignment: &[Fr]) -> R1CSResult<Proof<Bn254>> {
    assert_eq!(num_inputs, matrices.num_instance_variables);
    assert_eq!(num_constraints, matrices.num_constraints);
    let prover = prover_for(pk, matrices)?;
    let (a, b, c) = prover.lock().unwrap().prove(r, s, full_assignment)?;   // -> g16_prove; device failures panic like .unwrap()
    Ok(Proof { a, b, c })
}
