/*
 * bellman_b200 -- C ABI of the B200-native Groth16 proving back-end for bellman.
 *
 * This is the drop-in boundary for the two hot paths behind groth16::create_proof
 * (SURVEY.md 8b).  bellman has no FFI seam of its own; each entry point below
 * cites the Rust item a thin shim would forward to it (paths relative to the
 * bellman source tree), and INTEGRATION.md shows that shim.
 *
 * Plain pointers and sizes only.  All functions return a bb_status; no exception
 * or abort crosses the boundary.  bb_last_error() gives a thread-local message.
 *
 * Data formats
 *   Fr element   32 bytes, 4 x u64 little-endian limbs.
 *                BB_FORM_MONTGOMERY: value * 2^256 mod r (bls12_381::Scalar's in-memory
 *                form, what EvaluationDomain's Vec<Scalar<Fr>> holds);
 *                BB_FORM_CANONICAL: the integer itself (PrimeField::to_repr /
 *                PrimeFieldBits::to_le_bits, what Exponent::Bits holds,
 *                src/multiexp.rs:166-182).
 *   G1 affine    96 bytes: x | y, each 6 x u64 LE limbs, Montgomery (value * 2^384 mod p).
 *   G2 affine    192 bytes: x.c0 | x.c1 | y.c0 | y.c1.
 *                The identity is encoded as all-zero bytes ((0,0) is on neither curve).
 *   density      bit i of word i/64 (LSB first) = "variable i appears in this query"
 *                (bitvec::BitVec<usize, Lsb0> raw storage of DensityTracker,
 *                src/multiexp.rs:118-157).  NULL = FullDensity (:97-116).
 */
#ifndef BELLMAN_B200_H
#define BELLMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bb_status {
    BB_OK = 0,
    /* SynthesisError values that can originate on this path (src/lib.rs:304-319) */
    BB_ERR_POLYNOMIAL_DEGREE_TOO_LARGE = 1, /* domain.rs:57-59 */
    BB_ERR_UNEXPECTED_IDENTITY = 2,         /* multiexp.rs:63-65, prover.rs:320-324 */
    BB_ERR_IO_UNEXPECTED_EOF = 3,           /* multiexp.rs:55-61,74-80 */
    BB_ERR_DENSITY_MISMATCH = 5,            /* the assert! at multiexp.rs:324-329 (a panic upstream) */
    /* back-end errors */
    BB_ERR_ARG = 16,
    BB_ERR_CUDA = 17,
    BB_ERR_NO_DEVICE = 18,
    BB_ERR_OOM = 19
} bb_status;

typedef enum bb_form { BB_FORM_CANONICAL = 0, BB_FORM_MONTGOMERY = 1 } bb_form;

/* EvaluationDomain methods, src/domain.rs:81-125 */
typedef enum bb_ntt_mode { BB_NTT_FFT = 0, BB_NTT_IFFT = 1, BB_NTT_COSET_FFT = 2, BB_NTT_ICOSET_FFT = 3 } bb_ntt_mode;

typedef enum bb_group { BB_G1 = 1, BB_G2 = 2 } bb_group;

typedef struct bb_ctx bb_ctx;         /* one CUDA device; stands where multicore::Worker stands (src/multicore.rs:21-92) */
typedef struct bb_bases bb_bases;     /* device-resident Arc<Vec<G::Affine>> (groth16/src/lib.rs:227-243) */
typedef struct bb_msm_job bb_msm_job; /* Waiter<Result<G, SynthesisError>> (src/multicore.rs:94-118) */
typedef struct bb_crs bb_crs;         /* device-resident groth16::Parameters (groth16/src/lib.rs:222-244) */
typedef struct bb_prove bb_prove;     /* one create_proof in flight between bb_groth16_prove_begin and _end */

const char* bb_last_error(void);
int bb_version(void);

/* ---- context ------------------------------------------------------------------------ */
/* Worker::new() (src/multicore.rs:24-27).  `device` is a CUDA ordinal.  Fails with
 * BB_ERR_NO_DEVICE when no CUDA device is usable: there is no CPU fallback. */
int bb_ctx_create(int device, bb_ctx** out);
void bb_ctx_destroy(bb_ctx* ctx);
/* tuning knobs; results never depend on them.  Keys:
 *   msm_window_bits    window size c of the signed-digit Pippenger (0 = by size)
 *   msm_affine_rounds  batched-affine halving rounds before the XYZZ stage (-1 = by mean bucket fill, 0 = none)
 *   msm_affine_batch   pairs per thread in those rounds (default 16)
 *   msm_affine_tma     1 = dense rounds of G1 jobs staged by cp.async.bulk + mbarrier (default 0: measured neutral)
 *   msm_reduce_2d      0 = bucket reduction by the serial recursion over whole windows (default 1: row/column sums first)
 *   msm_reduce_k, msm_reduce_k1   entries per thread of the serial recursion (powers of two, default 4)
 *   msm_big_cap        bucket size above which a bucket is cut into tasks (0 = by mean fill)
 *   msm_precompute     keep window multiples of every base vector resident (see bb_bases_precompute): 1 = one bucket
 *                      array per window, added slot-wise; 2 = ONE bucket array for all windows and halving rounds by its fill
 *   msm_unified_rows_log   msm_precompute = 2: the rounds stop at about 2^this rows per bucket (default 3)
 *   msm_precompute_groups  which vectors use the tables: mask of 1 (G1) and 2 (G2), default 3
 *   shard_windows      multi-GPU: window groups per base range (default 4; 1 = base ranges only)
 *   ntt_radix8         1 = register radix-8 NTT windows (default 0: measured slower), ntt_tile_log, ntt_col_bits
 *   profile            1 = CUDA-event timing of the MSM stages (bb_profile_read) */
int bb_ctx_set_option(bb_ctx* ctx, const char* key, long value);
int bb_ctx_synchronize(bb_ctx* ctx);
/* counters for the harness: number of kernels this context has launched */
uint64_t bb_ctx_kernel_launches(const bb_ctx* ctx);

/* device buffers, so a caller can keep inputs resident across calls (bench "value" leg) */
int bb_device_alloc(bb_ctx* ctx, size_t bytes, void** d_out);
int bb_device_free(bb_ctx* ctx, void* d_ptr);
int bb_device_upload(bb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int bb_device_download(bb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* ---- NTT: EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} (src/domain.rs:81-125),
 *      i.e. best_fft (:261-269) plus the fused scale / distribute_powers passes --------- */
/* In place on a host buffer of 2^log_n Fr elements.  log_n >= 32 (= Fr::S) returns
 * BB_ERR_POLYNOMIAL_DEGREE_TOO_LARGE like from_coeffs (:55-59). */
int bb_ntt(bb_ctx* ctx, void* fr_inout, uint32_t log_n, int mode, int form);
/* same, on a device buffer (Montgomery form) */
int bb_ntt_device(bb_ctx* ctx, void* d_fr_inout, uint32_t log_n, int mode);

/* EvaluationDomain's O(n) methods as stand-alone calls on host buffers (Montgomery Fr); inside
 * bb_h_poly / bb_groth16_prove they are fused into the transforms instead.
 *   op 0  mul_assign   a[i] *= b[i]          (src/domain.rs:154-170)
 *   op 1  sub_assign   a[i] -= b[i]          (:173-189)
 *   op 2  scale        a[i] *= k             (divide_by_z_on_coset with k = 1/z(g), :139-151)
 *   op 3  distribute_powers  a[i] *= k^i     (:101-113) */
int bb_domain_pointwise(bb_ctx* ctx, int op, void* fr_a_inout, const void* fr_b, size_t n, const void* fr_k);

/* The H-polynomial block of create_proof (groth16/src/prover.rs:221-240): from_coeffs on
 * a,b,c (pads to m = 2^k >= n), 3x ifft, 3x coset_fft, mul_assign, sub_assign,
 * divide_by_z_on_coset, icoset_fft, truncate to m-1 coefficients.
 * a,b,c: n_constraints Fr each (Montgomery, host).  h_out: room for m-1 Fr, canonical
 * integers (the Exponent::Bits form multiexp consumes, prover.rs:242).  *m_out = m. */
int bb_h_poly(bb_ctx* ctx, const void* a, const void* b, const void* c, size_t n_constraints,
              void* h_out, size_t* m_out);

/* ---- bases ---------------------------------------------------------------------------- */
/* Uploads n affine points.  `global_offset`/`global_len` describe a shard of a larger base
 * vector for multi-GPU runs (the shard holds indices [global_offset, global_offset+n) of a
 * vector of global_len points); single-GPU callers pass 0 and n. */
int bb_bases_upload(bb_ctx* ctx, int group, const void* affine, size_t n, size_t global_offset,
                    size_t global_len, bb_bases** out);
void bb_bases_free(bb_bases* b);
/* Builds the table of window multiples 2^(c w) P_i of a resident base vector (c chosen from its
 * length).  MSMs over these bases then accumulate every window into one bucket set: one summation
 * by parts and no Horner fold instead of one per window (multiexp.rs:271-300), at (windows x) the
 * base storage.  Results are unchanged.  With bb_ctx_set_option(ctx, "msm_precompute", 1 or 2) the table
 * is built on first use instead; 2 also sorts the digits of all windows into ONE bucket array (deeper halving
 * rounds, a sixteenth of the bucket reduction) -- the form bb_groth16_autotune measures against the default. */
int bb_bases_precompute(bb_ctx* ctx, bb_bases* bases);
/* Frees that table again.  No MSM over these bases may be in flight. */
int bb_bases_drop_table(bb_bases* bases);

/* ---- MSM: multiexp() (src/multiexp.rs:305-332) ---------------------------------------- */
/* Sum over i with density bit set of scalars[i] * bases[base_offset + rank(i)], where rank(i)
 * counts set density bits below i (bases are compacted to set bits, App. C.3).
 * Semantics preserved from multiexp_inner (:242-265): a zero scalar skips its base even if
 * that base is the identity; an identity base under a non-zero scalar is
 * BB_ERR_UNEXPECTED_IDENTITY; running out of bases is BB_ERR_IO_UNEXPECTED_EOF; a density of
 * the wrong length is BB_ERR_DENSITY_MISMATCH.  Returns immediately; the job runs on its own
 * CUDA stream (prover.rs:244-318 starts 8 of these before the first wait()).
 * Host buffers must stay valid until bb_msm_wait returns. */
int bb_msm_async(bb_ctx* ctx, const bb_bases* bases, size_t base_offset,
                 const uint64_t* density_bits, size_t density_len,
                 const void* scalars, size_t n_scalars, int form, bb_msm_job** out);
/* scalars already in HBM (canonical or Montgomery per `form`) */
int bb_msm_async_device(bb_ctx* ctx, const bb_bases* bases, size_t base_offset,
                        const uint64_t* density_bits, size_t density_len,
                        const void* d_scalars, size_t n_scalars, int form, bb_msm_job** out);
/* Waiter::wait() (src/multicore.rs:98-108).  Writes the result as an affine point in the
 * format above (G1 96 B / G2 192 B), frees the job.  With a sharded `bases` the result is
 * this shard's partial sum. */
int bb_msm_wait(bb_msm_job* job, void* out_affine);

/* small group helpers for the host side of a multi-GPU reduction (affine in/out) */
int bb_point_add(int group, const void* a_affine, const void* b_affine, void* out_affine);
int bb_point_mul(int group, const void* a_affine, const void* fr_scalar, int form, void* out_affine);
/* GroupEncoding::to_bytes: G1 48 B / G2 96 B compressed ZCash encoding */
int bb_point_compress(int group, const void* affine, uint8_t* out);
/* in-place canonical <-> Montgomery conversion of n Fp coordinates (host; parameter files) */
int bb_fp_convert(void* fp_inout, size_t n, int to_montgomery);
/* The checks G1Affine/G2Affine::from_uncompressed applies after decoding -- what Parameters::read does
 * with checked = true and VerifyingKey::read always (groth16/src/lib.rs:158-183,289-330): every point on
 * the curve and, with check_subgroup != 0, of order r.  n affine points in the ABI format, host memory;
 * computed on the device.  *first_bad = index of the first offending point (SIZE_MAX if none), *why = 1
 * off the curve, 2 outside the subgroup.  The identity passes (callers decide where it is allowed). */
int bb_points_validate(bb_ctx* ctx, int group, const void* affine, size_t n, int check_subgroup, size_t* first_bad, int* why);
/* out[i] = [k_i] * generator, computed on the device (fixed-base); used to manufacture
 * synthetic CRS material of benchmark size (generator.rs:271-296,398-415 equivalent). */
int bb_fixed_base_mul(bb_ctx* ctx, int group, const void* fr_scalars, size_t n, int form, void* out_affine);

/* ---- whole prover: create_proof after synthesis (groth16/src/prover.rs:217-360) -------- */
typedef struct bb_crs_desc {
    /* VerifyingKey parts the prover reads (prover.rs:219,320-337) */
    const void* alpha_g1; const void* beta_g1; const void* delta_g1;   /* 96 B each  */
    const void* beta_g2;  const void* delta_g2;                        /* 192 B each */
    /* Parameters vectors (groth16/src/lib.rs:227-243) */
    const void* h;    size_t h_len;
    const void* l;    size_t l_len;
    const void* a;    size_t a_len;
    const void* b_g1; size_t b_g1_len;
    const void* b_g2; size_t b_g2_len;
    /* multi-GPU: this process is shard `shard_index` of `shard_count`.  A shard is a contiguous
     * base range of every vector above, crossed with a subset of the MSM windows (any partition
     * of the (base, window) pairs folds to the same point): the library splits the windows into
     * up to 4 groups first and the base vectors second -- 8 shards = 2 base ranges x 4 window
     * groups; bb_ctx_set_option(ctx, "shard_windows", 1) gives base ranges alone.  The arrays
     * passed are the full vectors; only this shard's base range is uploaded.  Single GPU: 0 / 1. */
    uint32_t shard_index; uint32_t shard_count;
} bb_crs_desc;

int bb_crs_create(bb_ctx* ctx, const bb_crs_desc* desc, bb_crs** out);
void bb_crs_destroy(bb_crs* crs);
/* bb_bases_precompute / bb_bases_drop_table over the five vectors of a resident key */
int bb_crs_precompute(bb_ctx* ctx, bb_crs* crs);
int bb_crs_drop_tables(bb_crs* crs);

/* What ProvingAssignment holds when synthesis is done (prover.rs:57-71,193-215), i.e.
 * including the trailing "input * 0 = 0" constraints. */
typedef struct bb_witness {
    const void* a; const void* b; const void* c; size_t n_constraints;   /* Fr, Montgomery */
    const void* input_assignment; size_t n_inputs;                        /* Fr, Montgomery */
    const void* aux_assignment;   size_t n_aux;                           /* Fr, Montgomery */
    const uint64_t* a_aux_density;   /* n_aux bits   */
    const uint64_t* b_input_density; /* n_inputs bits */
    const uint64_t* b_aux_density;   /* n_aux bits   */
    /* non-zero: a, b, c, input_assignment and aux_assignment are DEVICE pointers (inputs already
     * resident in HBM; a, b, c are left untouched).  The density maps are always host memory. */
    int on_device;
} bb_witness;

/* Eight partial MSM results of one proof, in the order prover.rs starts them:
 * h, l, a_inputs, a_aux, b_g1_inputs, b_g1_aux (G1, 96 B each) then b_g2_inputs, b_g2_aux
 * (G2, 192 B each): 6*96 + 2*192 = 960 bytes. */
#define BB_PARTIALS_BYTES 960
/* NTT pipeline + the 8 MSMs over this context's CRS shard. */
int bb_groth16_prove_partials(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, uint8_t partials[BB_PARTIALS_BYTES]);
/* The same in two steps, for multi-GPU provers that split the H pipeline by polynomial.  _begin uploads the
 * assignments and queues the seven witness MSMs (prover.rs:263-318) -- they need nothing from the H pipeline
 * -- and returns at once.  Each of a, b, c is taken through from_coeffs / ifft / coset_fft
 * (prover.rs:225-230) by ONE rank with bb_h_coset_evals and the three result vectors (m Fr each, device
 * memory) are broadcast (NCCL).  _end(state, evals_a, evals_b, evals_c, partials) then runs mul_assign /
 * sub_assign / divide_by_z_on_coset / icoset_fft (:232-237) and the h MSM (:238-244), waits for all eight MSMs
 * and frees the state; with three NULL pointers it runs the whole H pipeline from w->a, w->b, w->c itself.
 * evals_a is clobbered.  The witness arrays must stay valid until _end returns. */
int bb_groth16_prove_begin(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, bb_prove** out);
int bb_groth16_prove_end(bb_prove* state, const void* d_evals_a, const void* d_evals_b, const void* d_evals_c,
                         uint8_t partials[BB_PARTIALS_BYTES]);
/* d_out[0..m) = coset_fft(ifft(from_coeffs(poly))) for one polynomial of n_constraints Fr (Montgomery; host
 * memory, or device memory with on_device != 0); m = next power of two >= n_constraints.  Blocking. */
int bb_h_coset_evals(bb_ctx* ctx, const void* poly, size_t n_constraints, int on_device, void* d_out);
/* The same without blocking: _async queues the copy and the two transforms on the context's high-priority stream
 * and returns at once (poly and d_out must stay valid); _wait blocks until every evaluation queued so far is
 * complete.  Lets a rank launch its transforms FIRST and queue the witness MSMs (bb_groth16_prove_begin, ~1 ms of
 * host time) while they run. */
int bb_h_coset_evals_async(bb_ctx* ctx, const void* poly, size_t n_constraints, int on_device, void* d_out);
int bb_h_coset_evals_wait(bb_ctx* ctx);
/* Sums `count` partial sets (one per shard) and applies prover.rs:320-360 + Proof::write
 * (groth16/src/lib.rs:39-45).  r, s: 32-byte canonical little-endian scalars. */
int bb_groth16_finalize(const bb_crs* crs, const uint8_t* partials, size_t count,
                        const uint8_t r[32], const uint8_t s[32], uint8_t proof[192]);
/* The same in two steps, so that the five scalar multiplications that need no MSM result
 * (delta*r, delta*s, delta*rs, alpha*s, beta*r -- prover.rs:326-337) can run on a host thread while
 * the devices compute the partial sums: finalize_static fills BB_PROOF_STATIC_BYTES opaque bytes,
 * finalize_with consumes them.  bb_groth16_finalize == finalize_static + finalize_with;
 * bb_groth16_prove overlaps them internally.
 * Side channels: the host-side scalar multiplications by r, s and r s run a REGULAR ladder -- signed odd digits
 * (none zero, none skipped), the table read in full under a mask, even scalars as (k + 1) P - P -- so neither the
 * sequence of operations nor any address depends on the scalar.  Not covered: the field products end with a
 * data-dependent final subtraction (the reference uses bls12_381's constant-time arithmetic throughout). */
#define BB_PROOF_STATIC_BYTES 768
int bb_groth16_finalize_static(const bb_crs* crs, const uint8_t r[32], const uint8_t s[32], uint8_t* static_out);
int bb_groth16_finalize_with(const bb_crs* crs, const uint8_t* partials, size_t count, const uint8_t r[32],
                             const uint8_t s[32], const uint8_t* static_in, uint8_t* proof);
/* single-GPU create_proof after synthesis (groth16/src/prover.rs:217-360): partials + finalize.
 * When several errors apply the one the reference would return is reported: the density assert
 * of multiexp() first (call order), then delta = identity (prover.rs:320-324, checked before any
 * wait), then the MSM errors in the order of the reference's waits: a_inputs, a_aux,
 * b_g1_inputs, b_g1_aux, b_g2_inputs, b_g2_aux, h, l (:339-354). */
int bb_groth16_prove(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w,
                     const uint8_t r[32], const uint8_t s[32], uint8_t proof[192]);

/* ---- tuning per key ------------------------------------------------------------------------
 * The MSM has several forms with identical results (per-window bucket sets; one bucket set over resident window
 * multiples with deeper halving rounds, which costs (windows x) the base storage).  Which is fastest depends on
 * the device, the key size and the shard count, and a key serves many proofs, so the choice is MEASURED once per
 * key: bb_groth16_autotune proves the given witness with every form (one checked proof, then `reps` timed ones,
 * fastest taken), admits a form only if its 960 bytes of partial sums equal the default form's, leaves the context
 * and the key configured for the fastest one and frees the tables if they lost.  ms_out (bb_tuning_count()
 * doubles, may be NULL): milliseconds per form; -1 = not available (no room for the tables), -2 = failed,
 * -3 = different partial sums.  *chosen = index of the fastest eligible form, which is now active.
 * Sharded keys: all shards MUST run the same form (the forms cut the scalars into windows of different sizes, so a
 * shard's (base, window) pairs -- and its partial sums -- depend on the form; only the sum over all shards does not).
 * For a key with shard_count > 1 this call therefore only measures (no comparison, *chosen = the fastest form here)
 * and leaves the DEFAULT form active; the shards agree on one index -- time the whole sharded proof, compare the
 * final proof bytes -- and each selects it with bb_crs_apply_tuning.  Form 0 is the default. */
int bb_tuning_count(void);
const char* bb_tuning_name(int index);
int bb_crs_apply_tuning(bb_ctx* ctx, bb_crs* crs, int index);
int bb_groth16_autotune(bb_ctx* ctx, bb_crs* crs, const bb_witness* w, int reps, int* chosen, double* ms_out);

/* ---- profiling: CUDA-event timing of the dominant kernels, on the stream they run on ------ */
/* bb_ctx_set_option(ctx, "profile", 1) makes every MSM job bracket its bucket-accumulation
 * kernel with cudaEvents.  bb_profile_read returns, for `what` = "msm_accumulate_g1" |
 * "msm_accumulate_g2" | "msm_total_g1" | "msm_total_g2" | "ntt_pass", the summed device
 * milliseconds, the number of launches (or jobs) and the number of units they processed
 * ((base, scalar) pairs; NTT points) since the last bb_profile_reset. */
int bb_profile_read(bb_ctx* ctx, const char* what, double* ms, uint64_t* launches, uint64_t* units);
int bb_profile_reset(bb_ctx* ctx);
/* bytes the MSM / prover entry points have copied host->device and device->host so far */
int bb_ctx_bytes_copied(const bb_ctx* ctx, uint64_t* h2d, uint64_t* d2h);

/* Bench generators (bb_synth_*) and parity-test diagnostics (bb_selftest_*, bb_diag_*) are not part
 * of the bellman-facing surface: they are declared in bellman_b200_diag.h. */

#ifdef __cplusplus
}
#endif
#endif /* BELLMAN_B200_H */
