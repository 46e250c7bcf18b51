//! Safe layer over `bellman_b200_sys`.  Every item cites the bellman item it stands for
//! (paths relative to the bellman source tree).
use std::any::Any;
use std::ffi::{CStr, CString};
use std::io;
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::ptr;
use std::sync::Arc;

use bellman::groth16::{Parameters, Proof};
use bellman::SynthesisError;
use bellman_b200_sys::*;
use bls12_381::{Bls12, G1Affine, G1Projective, G2Affine, G2Projective, Scalar};
use ff::PrimeField;
use group::{prime::PrimeCurveAffine, Curve, UncompressedEncoding};

fn last_error() -> String {
    unsafe { CStr::from_ptr(bb_last_error()).to_string_lossy().into_owned() }
}

/// bb_status -> SynthesisError (src/lib.rs:304-319)
fn status(rc: i32) -> Result<(), SynthesisError> {
    match rc {
        BB_OK => Ok(()),
        BB_ERR_POLYNOMIAL_DEGREE_TOO_LARGE => Err(SynthesisError::PolynomialDegreeTooLarge),
        BB_ERR_UNEXPECTED_IDENTITY => Err(SynthesisError::UnexpectedIdentity),
        BB_ERR_IO_UNEXPECTED_EOF => Err(io::Error::new(io::ErrorKind::UnexpectedEof, "expected more bases from source").into()),
        BB_ERR_DENSITY_MISMATCH => panic!("assertion failed: query_size == exponents.len()"), // the assert! of multiexp.rs:324-329
        _ => Err(io::Error::new(io::ErrorKind::Other, last_error()).into()),
    }
}

/// multicore::Worker (src/multicore.rs:21-92): one CUDA device.
pub struct GpuWorker { pub(crate) ctx: *mut bb_ctx }
unsafe impl Send for GpuWorker {}
unsafe impl Sync for GpuWorker {}

impl GpuWorker {
    /// Worker::new().  Fails when no CUDA device is usable: there is no CPU fallback.
    pub fn new(device: i32) -> Result<Self, SynthesisError> {
        let mut ctx = ptr::null_mut();
        status(unsafe { bb_ctx_create(device, &mut ctx) })?;
        Ok(GpuWorker { ctx })
    }
    pub fn set_option(&self, key: &str, value: i64) -> Result<(), SynthesisError> {
        let k = CString::new(key).unwrap();
        status(unsafe { bb_ctx_set_option(self.ctx, k.as_ptr(), value as _) })
    }
}
impl Drop for GpuWorker {
    fn drop(&mut self) { unsafe { bb_ctx_destroy(self.ctx) } }
}

/// Coordinates of an affine point in the ABI's form: little-endian Montgomery limbs, identity = zeros.
/// The uncompressed ZCash encoding is big-endian canonical; bb_fp_convert turns canonical limbs into
/// Montgomery ones, so no private layout of the bls12_381 crate is relied on.
pub trait B200Affine: PrimeCurveAffine + UncompressedEncoding {
    const GROUP: i32;
    const WORDS: usize;          // u64 limbs per point: 12 (G1) / 24 (G2)
    /// order of the 48-byte coordinates inside the uncompressed encoding -> ABI order
    const COORD_ORDER: &'static [usize];
    fn to_b200(&self, out: &mut [u64]) {
        if bool::from(self.is_identity()) { out.iter_mut().for_each(|w| *w = 0); return; }
        let enc = self.to_uncompressed();
        let bytes = enc.as_ref();
        for (k, &src) in Self::COORD_ORDER.iter().enumerate() {
            let be = &bytes[48 * src..48 * (src + 1)];
            for limb in 0..6 {
                let mut w = 0u64;
                for b in 0..8 { w |= (be[47 - (8 * limb + b)] as u64) << (8 * b); }
                out[6 * k + limb] = w;
            }
        }
        // (the three flag bits of byte 0 are clear in the uncompressed encoding of a non-identity point)
        unsafe { bb_fp_convert(out.as_mut_ptr() as *mut c_void, Self::WORDS / 6, 1) };
    }
    fn from_b200(words: &[u64]) -> Self {
        if words.iter().all(|w| *w == 0) { return Self::identity(); }
        let mut w = words.to_vec();
        unsafe { bb_fp_convert(w.as_mut_ptr() as *mut c_void, Self::WORDS / 6, 0) };
        let mut enc = Self::Uncompressed::default();
        {
            let bytes = enc.as_mut();
            for (k, &dst) in Self::COORD_ORDER.iter().enumerate() {
                for limb in 0..6 {
                    for b in 0..8 { bytes[48 * dst + 47 - (8 * limb + b)] = (w[6 * k + limb] >> (8 * b)) as u8; }
                }
            }
        }
        Option::from(Self::from_uncompressed_unchecked(&enc)).expect("device returned a valid encoding")
    }
}
impl B200Affine for G1Affine { const GROUP: i32 = BB_G1; const WORDS: usize = 12; const COORD_ORDER: &'static [usize] = &[0, 1]; }
// uncompressed G2: x.c1 | x.c0 | y.c1 | y.c0; ABI: x.c0 | x.c1 | y.c0 | y.c1
impl B200Affine for G2Affine { const GROUP: i32 = BB_G2; const WORDS: usize = 24; const COORD_ORDER: &'static [usize] = &[1, 0, 3, 2]; }

/// Device-resident `Arc<Vec<G::Affine>>` (groth16/src/lib.rs:227-243), uploaded once per key.
pub struct GpuBases<A: B200Affine> { pub(crate) handle: *mut bb_bases, _a: PhantomData<A> }
unsafe impl<A: B200Affine> Send for GpuBases<A> {}
unsafe impl<A: B200Affine> Sync for GpuBases<A> {}

impl<A: B200Affine> GpuBases<A> {
    pub fn upload(worker: &GpuWorker, points: &[A]) -> Result<Self, SynthesisError> {
        let mut flat = vec![0u64; points.len() * A::WORDS];
        for (p, out) in points.iter().zip(flat.chunks_mut(A::WORDS)) { p.to_b200(out); }
        let mut handle = ptr::null_mut();
        status(unsafe { bb_bases_upload(worker.ctx, A::GROUP, flat.as_ptr() as *const c_void, points.len(), 0, points.len(), &mut handle) })?;
        Ok(GpuBases { handle, _a: PhantomData })
    }
}
impl<A: B200Affine> Drop for GpuBases<A> {
    fn drop(&mut self) { unsafe { bb_bases_free(self.handle) } }
}

/// `DensityTracker { bv: BitVec }` / `FullDensity` (src/multiexp.rs:88-157) as the ABI wants it:
/// raw LSB-first words and the length in bits; `FullDensity` is (null, 0).
pub trait DensityWords { fn raw_words(&self) -> (*const u64, usize); }
pub struct FullDensity;
impl DensityWords for FullDensity { fn raw_words(&self) -> (*const u64, usize) { (ptr::null(), 0) } }
impl DensityWords for bitvec::vec::BitVec<usize, bitvec::order::Lsb0> {
    fn raw_words(&self) -> (*const u64, usize) { (self.as_raw_slice().as_ptr() as *const u64, self.len()) }   // 64-bit targets
}

/// `Waiter<Result<G, SynthesisError>>` (src/multicore.rs:94-118)
pub struct GpuWaiter<A: B200Affine> { job: *mut bb_msm_job, _keep: (Arc<dyn Any + Send + Sync>, Arc<dyn Any + Send + Sync>), _a: PhantomData<A> }

impl<A: B200Affine> GpuWaiter<A> {
    /// Waiter::wait (src/multicore.rs:98-108): blocks on the job's CUDA stream, folds, returns the point.
    pub fn wait(self) -> Result<A::Curve, SynthesisError> {
        let mut out = vec![0u64; A::WORDS];
        status(unsafe { bb_msm_wait(self.job, out.as_mut_ptr() as *mut c_void) })?;
        Ok(A::from_b200(&out).to_curve())
    }
}

/// Drop-in for `bellman::multiexp::multiexp` (src/multiexp.rs:305-332).  `exponents` are the canonical
/// little-endian integers `Exponent::Bits` wraps (`to_le_bits`, :179); zero and one need no special
/// casing, the kernels classify them (`Exponent::Zero` / `Exponent::One`, :172-182).
pub fn multiexp<A, D>(pool: &GpuWorker, bases: (Arc<GpuBases<A>>, usize), density_map: &D, exponents: Arc<Vec<[u64; 4]>>) -> GpuWaiter<A>
where A: B200Affine + 'static, D: DensityWords {
    let (words, len) = density_map.raw_words();
    let mut job = ptr::null_mut();
    let rc = unsafe {
        bb_msm_async(pool.ctx, bases.0.handle, bases.1, words, len, exponents.as_ptr() as *const c_void, exponents.len(),
                     BB_FORM_CANONICAL, &mut job)
    };
    assert_eq!(rc, BB_OK, "{}", last_error());                 // argument errors only; run-time errors surface in wait()
    GpuWaiter { job, _keep: (bases.0, exponents), _a: PhantomData }
}

/// EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} (src/domain.rs:81-125) in place on `coeffs`
/// (length 2^log_n, the in-memory `Scalar`s are passed through their canonical representation).
pub fn ntt(pool: &GpuWorker, coeffs: &mut [Scalar], log_n: u32, mode: i32) -> Result<(), SynthesisError> {
    assert_eq!(coeffs.len(), 1usize << log_n);
    let mut flat: Vec<[u8; 32]> = coeffs.iter().map(|s| s.to_repr()).collect();
    status(unsafe { bb_ntt(pool.ctx, flat.as_mut_ptr() as *mut c_void, log_n, mode, BB_FORM_CANONICAL) })?;
    for (c, r) in coeffs.iter_mut().zip(flat.iter()) { *c = Option::from(Scalar::from_repr(*r)).expect("canonical output"); }
    Ok(())
}

/// Device-resident `groth16::Parameters` (groth16/src/lib.rs:222-244).
pub struct GpuParameters { pub(crate) ctx: *mut bb_ctx, pub(crate) handle: *mut bb_crs }
unsafe impl Send for GpuParameters {}
unsafe impl Sync for GpuParameters {}

impl GpuParameters {
    pub fn upload(worker: &GpuWorker, p: &Parameters<Bls12>) -> Result<Self, SynthesisError> {
        fn flat<A: B200Affine>(v: &[A]) -> Vec<u64> {
            let mut f = vec![0u64; v.len() * A::WORDS];
            for (p, out) in v.iter().zip(f.chunks_mut(A::WORDS)) { p.to_b200(out); }
            f
        }
        let (alpha, beta1, delta1) = (flat(&[p.vk.alpha_g1]), flat(&[p.vk.beta_g1]), flat(&[p.vk.delta_g1]));
        let (beta2, delta2) = (flat(&[p.vk.beta_g2]), flat(&[p.vk.delta_g2]));
        let (h, l, a, b1, b2) = (flat(&p.h[..]), flat(&p.l[..]), flat(&p.a[..]), flat(&p.b_g1[..]), flat(&p.b_g2[..]));
        let desc = bb_crs_desc {
            alpha_g1: alpha.as_ptr() as _, beta_g1: beta1.as_ptr() as _, delta_g1: delta1.as_ptr() as _,
            beta_g2: beta2.as_ptr() as _, delta_g2: delta2.as_ptr() as _,
            h: h.as_ptr() as _, h_len: p.h.len(), l: l.as_ptr() as _, l_len: p.l.len(), a: a.as_ptr() as _, a_len: p.a.len(),
            b_g1: b1.as_ptr() as _, b_g1_len: p.b_g1.len(), b_g2: b2.as_ptr() as _, b_g2_len: p.b_g2.len(),
            shard_index: 0, shard_count: 1,
        };
        let mut handle = ptr::null_mut();
        status(unsafe { bb_crs_create(worker.ctx, &desc, &mut handle) })?;
        Ok(GpuParameters { ctx: worker.ctx, handle })
    }
}
impl Drop for GpuParameters {
    fn drop(&mut self) { unsafe { bb_crs_destroy(self.handle) } }
}

/// What `ProvingAssignment` holds when synthesis is done (groth16/src/prover.rs:57-71,193-215); the fields
/// are private upstream, so the `b200` feature exposes them to this module (`pub(crate)` in-tree).
pub struct WitnessView<'a> {
    pub a: &'a [Scalar], pub b: &'a [Scalar], pub c: &'a [Scalar],
    pub input_assignment: &'a [Scalar], pub aux_assignment: &'a [Scalar],
    pub a_aux_density: &'a bitvec::vec::BitVec<usize, bitvec::order::Lsb0>,
    pub b_input_density: &'a bitvec::vec::BitVec<usize, bitvec::order::Lsb0>,
    pub b_aux_density: &'a bitvec::vec::BitVec<usize, bitvec::order::Lsb0>,
}

impl<'a> WitnessView<'a> {
    fn as_ffi(&self) -> bb_witness {
        const _: () = assert!(std::mem::size_of::<Scalar>() == 32);
        bb_witness {
            a: self.a.as_ptr() as _, b: self.b.as_ptr() as _, c: self.c.as_ptr() as _, n_constraints: self.a.len(),
            input_assignment: self.input_assignment.as_ptr() as _, n_inputs: self.input_assignment.len(),
            aux_assignment: self.aux_assignment.as_ptr() as _, n_aux: self.aux_assignment.len(),
            a_aux_density: self.a_aux_density.as_raw_slice().as_ptr() as _,
            b_input_density: self.b_input_density.as_raw_slice().as_ptr() as _,
            b_aux_density: self.b_aux_density.as_raw_slice().as_ptr() as _,
            on_device: 0,
        }
    }
}

impl GpuParameters {
    /// Once per key (a key serves many proofs): proves `w` with every MSM form of the back-end on this device and
    /// leaves the key configured for the fastest form whose partial sums are byte-identical to the default form's
    /// (`bb_groth16_autotune`).  Returns the index chosen and the milliseconds measured per form.
    pub fn autotune(&mut self, w: &WitnessView, reps: i32) -> Result<(usize, Vec<f64>), SynthesisError> {
        let n = unsafe { bb_tuning_count() } as usize;
        let mut ms = vec![0f64; n];
        let mut chosen = 0i32;
        let wit = w.as_ffi();
        status(unsafe { bb_groth16_autotune(self.ctx, self.handle, &wit, reps, &mut chosen, ms.as_mut_ptr()) })?;
        Ok((chosen as usize, ms))
    }
}

/// Drop-in for the body of `groth16::create_proof` after synthesis (groth16/src/prover.rs:217-360).
/// `Scalar` is `[u64; 4]` in Montgomery form (R = 2^256), which is BB_FORM_MONTGOMERY byte for byte; the
/// cast below is version-pinned (`bls12_381 = "=0.8.0"`) because `#[repr(transparent)]` is not promised.
pub fn create_proof_b200(w: &WitnessView, crs: &GpuParameters, r: Scalar, s: Scalar) -> Result<Proof<Bls12>, SynthesisError> {
    let wit = w.as_ffi();
    let mut bytes = [0u8; 192];
    status(unsafe { bb_groth16_prove(crs.ctx, crs.handle, &wit, r.to_repr().as_ptr(), s.to_repr().as_ptr(), bytes.as_mut_ptr()) })?;
    Proof::read(&bytes[..]).map_err(Into::into)               // the same 192 bytes Proof::write emits (lib.rs:39-45)
}

#[allow(dead_code)]
fn _projective_types_are_what_wait_returns(_: G1Projective, _: G2Projective) {}
