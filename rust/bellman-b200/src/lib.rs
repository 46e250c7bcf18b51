//! bellman -> libbellman_b200.so.  Source only (no Rust toolchain in the build image).
//!
//! The seam is the pair of call sites `groth16::create_proof` makes into `bellman::multiexp` and
//! `bellman::domain` (groth16/src/prover.rs:221-318); everything upstream of it -- `Circuit`,
//! `ConstraintSystem`, `ProvingAssignment`, `create_random_proof`'s signature, `Proof::write` -- is unchanged.
//! A maintainer enables this crate behind a `b200` feature and routes the body of `create_proof` after
//! synthesis to [`gpu::create_proof_b200`].
pub mod gpu;
