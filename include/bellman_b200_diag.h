/*
 * bellman_b200 -- bench generators and parity-test diagnostics.
 *
 * NOT part of the drop-in boundary (that is bellman_b200.h): nothing here stands for a bellman
 * item.  bb_synth_* manufacture benchmark-size inputs in HBM, bb_selftest_* / bb_diag_* expose
 * single device routines to tests/ so that each can be compared with the CPU oracle.
 */
#ifndef BELLMAN_B200_DIAG_H
#define BELLMAN_B200_DIAG_H

#include "bellman_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic workload (bench only): the MiMC chain of groth16/tests/common/mod.rs:48-129
 *      through ProvingAssignment's bookkeeping (groth16/src/prover.rs:73-145,193-215) -------- */
/* pseudorandom canonical scalars (< 2^254) generated in HBM */
int bb_synth_scalars_device(bb_ctx* ctx, uint64_t seed, size_t n, void* d_out);
/* base vector [k_i]G with counter-based pseudorandom k_i, made on the device straight into a
 * bb_bases / bb_crs (shard-consistent across ranks) */
int bb_synth_bases(bb_ctx* ctx, int group, uint64_t seed, size_t n, size_t global_offset, size_t global_len, bb_bases** out);
int bb_synth_crs(bb_ctx* ctx, uint64_t seed, size_t h_len, size_t l_len, size_t a_len, size_t b_len,
                 uint32_t shard_index, uint32_t shard_count, bb_crs** out);
int bb_synth_mimc_shape(size_t rounds, uint64_t shape[7]);
int bb_synth_mimc_witness(size_t rounds, uint64_t seed, uint64_t* a, uint64_t* b, uint64_t* c,
                          uint64_t* inputs, uint64_t* aux, uint64_t* a_aux_density,
                          uint64_t* b_input_density, uint64_t* b_aux_density);

/* ---- diagnostics (used by the parity tests; not part of the bellman-facing surface) ------ */
/* element-wise on the device: field 0 = Fr, 1 = Fp (Montgomery); op 0 mul, 1 add, 2 sub, 3 sqr;
 * Fp only: 4 = inverse by the binary Euclidean algorithm (fp_inv_gcd), 5 = inverse as a^(p-2); 0 -> 0 */
int bb_selftest_field(bb_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n);
/* element-wise on affine points: op 0: a + b (mixed add); 1: 2a + b; 2: a + (2b - b) */
int bb_selftest_point(bb_ctx* ctx, int group, int op, const void* a, const void* b, void* out, size_t n);
/* out = sum_{i<D} (i+1) * P_i through the bucket-reduction kernels, K buckets per thread (G1) */
int bb_selftest_bucket_reduce(bb_ctx* ctx, const void* affine_pts, uint32_t D, uint32_t K, void* out_affine);

/* out = sum_i a[i] * b[i] mod r over two device arrays of canonical Fr (result canonical, host).
 * With bases [k_i]G from bb_synth_bases(seed) -- k = bb_synth_scalars_device(seed) -- and scalars e_i, [sum k_i e_i]G is what an MSM must return
 * (the naive == fast property of src/multiexp.rs:334-378 at any size). */
int bb_diag_fr_dot(bb_ctx* ctx, const void* d_a, const void* d_b, size_t n, void* out_fr);

#ifdef __cplusplus
}
#endif
#endif /* BELLMAN_B200_DIAG_H */
