// bellman_b200: the device side of groth16::create_proof after synthesis
// (/root/reference/groth16/src/prover.rs:217-360): CRS residency, the H pipeline, the eight
// multiexps in flight, and the host finalisation.
#include <chrono>
#include <functional>
#include <memory>
#include <string>
#include <thread>

#include "bb_internal.cuh"

using namespace bb;

struct bb_crs {
    bb_ctx* ctx = nullptr;
    bb_bases *h = nullptr, *l = nullptr, *a = nullptr, *b_g1 = nullptr, *b_g2 = nullptr;
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, delta_g2;
    uint32_t shard_index = 0, shard_count = 1;
};

namespace {

// Shard `idx` of `cnt` -> (base range b of nb) x (window shard v of nv).  Any partition of the
// (base, window) pairs of an MSM folds to the same point, so the shards split the windows first
// (that also divides the per-window bucket reduction, which base ranges alone do not) and the
// base vector second: nv = the largest divisor of cnt not above ctx->opt_shard_windows (4),
// nb = cnt / nv.  8 GPUs: 2 base ranges x 4 window shards.
void shard_policy(const bb_ctx* ctx, uint32_t idx, uint32_t cnt, uint32_t* b, uint32_t* nb, uint32_t* v, uint32_t* nv) {
    uint32_t lim = (uint32_t)(ctx->opt_shard_windows < 1 ? 1 : ctx->opt_shard_windows), w = 1;
    for (uint32_t d = 1; d <= lim && d <= cnt; d++) if (cnt % d == 0) w = d;
    *nv = w; *nb = cnt / w; *v = idx % w; *b = idx / w;
}

// scalar multiplication on the host (the five Affine * Fr of prover.rs:326-337 and the two
// MulAssign<Fr> of :342,351); k is a canonical little-endian 256-bit integer.  The scalars here are the
// proof's blinding factors r, s (and r s), so the sequence of operations does not depend on them:
//   * regular signed-digit recoding (Joye-Tunstall): with k made odd, k = 16^64 + sum_{i<64} d_i 16^i where
//     d_i = (((k >> 4i) | 1) & 31) - 16 is odd and non-zero -- every digit costs four doublings and ONE
//     addition, none is skipped;
//   * the table of odd multiples P, 3P, ..., 15P is read in full for every digit and the entry (and its sign) is
//     selected under a mask -- no secret-dependent address or branch;
//   * an even k is handled as (k + 1) P - P, both results computed, one selected under a mask.
// What is left: the XYZZ formulas branch on exceptional operands (identity, P = +-Q), which no digit of a
// random scalar produces for a point of prime order, and the host field arithmetic ends its products with a
// data-dependent final subtraction.  (The reference multiplies with the bls12_381 crate's constant-time routine.)
template <class T>
void masked_select(T* dst, const T& a, const T& b, uint32_t take_b) {     // take_b = 0 or 0xffffffff
    static_assert(sizeof(T) % 4 == 0, "whole words");
    uint32_t wa[sizeof(T) / 4], wb[sizeof(T) / 4];
    std::memcpy(wa, &a, sizeof(T));
    std::memcpy(wb, &b, sizeof(T));
    for (size_t i = 0; i < sizeof(T) / 4; i++) wa[i] = (wa[i] & ~take_b) | (wb[i] & take_b);
    std::memcpy(dst, wa, sizeof(T));
}

template <class F>
XYZZ<F> host_mul(const XYZZ<F>& p, const uint32_t* k_in) {
    uint32_t k[9];
    const uint32_t even = (~k_in[0]) & 1u;                  // k -> k + 1 when even (no carry out: the low bit is clear)
    for (int i = 0; i < 8; i++) k[i] = k_in[i];
    k[0] |= 1u;
    k[8] = 0;
    XYZZ<F> tab[8];                                         // (2j + 1) P
    const XYZZ<F> p2 = p.dbl();
    tab[0] = p;
    for (int j = 1; j < 8; j++) { tab[j] = tab[j - 1]; tab[j].add(p2); }
    XYZZ<F> acc = p;                                        // the leading digit is always 1 (bit 256)
    for (int i = 63; i >= 0; i--) {
        for (int d = 0; d < 4; d++) acc = acc.dbl();
        const uint32_t pos = 4u * (uint32_t)i, word = pos >> 5, off = pos & 31u;
        const uint64_t two = (uint64_t)k[word] | ((uint64_t)k[word + 1] << 32);
        const uint32_t v = ((uint32_t)(two >> off) | 1u) & 31u;          // odd, 1..31; digit = v - 16
        const uint32_t negative = 0u - (uint32_t)(v < 16u);               // mask
        const uint32_t mag = ((16u - v) & negative) | ((v - 16u) & ~negative);   // |digit|: odd, 1..15
        const uint32_t idx = mag >> 1;
        XYZZ<F> t = tab[0];
        for (uint32_t j = 1; j < 8; j++) masked_select(&t, t, tab[j], 0u - (uint32_t)(j == idx));
        const XYZZ<F> tn = t.neg();
        masked_select(&t, t, tn, negative);
        acc.add(t);
    }
    XYZZ<F> minus = acc;
    minus.add(p.neg());
    masked_select(&acc, acc, minus, 0u - even);
    return acc;
}

void fp_to_be(const Fp& v, uint8_t* out) {
    Fp c = fp_to_canonical(v);
    for (int i = 0; i < 12; i++)
        for (int b = 0; b < 4; b++) out[47 - (4 * i + b)] = (uint8_t)(c.l[i] >> (8 * b));
}
bool fp_lexi_larger(const Fp& v) {               // canonical(v) > (p-1)/2
    static const uint32_t half[12] = {BBC_FP_HALF_LIST};
    Fp c = fp_to_canonical(v);
    for (int i = 11; i >= 0; i--) {
        if (c.l[i] > half[i]) return true;
        if (c.l[i] < half[i]) return false;
    }
    return false;
}

}  // namespace

namespace bb {
void g1_compress(const G1Affine& p, uint8_t* out) {
    if (p.is_identity()) { std::memset(out, 0, 48); out[0] = 0xC0; return; }
    fp_to_be(p.x, out);
    out[0] |= 0x80;
    if (fp_lexi_larger(p.y)) out[0] |= 0x20;
}
void g2_compress(const G2Affine& p, uint8_t* out) {
    if (p.is_identity()) { std::memset(out, 0, 96); out[0] = 0xC0; return; }
    fp_to_be(p.x.c1, out);
    fp_to_be(p.x.c0, out + 48);
    out[0] |= 0x80;
    bool larger = p.y.c1.is_zero() ? fp_lexi_larger(p.y.c0) : fp_lexi_larger(p.y.c1);
    if (larger) out[0] |= 0x20;
}
G1X g1_host_mul(const G1X& p, const uint32_t* k) { return host_mul<Fp>(p, k); }
G2X g2_host_mul(const G2X& p, const uint32_t* k) { return host_mul<Fp2>(p, k); }
}  // namespace bb

extern "C" {

int bb_synth_bases(bb_ctx* ctx, int group, uint64_t seed, size_t n, size_t global_offset, size_t global_len, bb_bases** out);
void bb_crs_destroy(bb_crs* crs);

int bb_crs_create(bb_ctx* ctx, const bb_crs_desc* d, bb_crs** out) {
    if (!ctx || !d || !out) { set_error("bb_crs_create: null argument"); return BB_ERR_ARG; }
    if (!d->alpha_g1 || !d->beta_g1 || !d->delta_g1 || !d->beta_g2 || !d->delta_g2) { set_error("bb_crs_create: missing vk element"); return BB_ERR_ARG; }
    uint32_t cnt = d->shard_count ? d->shard_count : 1, idx = d->shard_index;
    if (idx >= cnt) { set_error("bb_crs_create: shard %u of %u", idx, cnt); return BB_ERR_ARG; }
    bb_crs* crs = new bb_crs();
    crs->ctx = ctx; crs->shard_index = idx; crs->shard_count = cnt;
    std::memcpy(&crs->alpha_g1, d->alpha_g1, 96); std::memcpy(&crs->beta_g1, d->beta_g1, 96);
    std::memcpy(&crs->delta_g1, d->delta_g1, 96);
    std::memcpy(&crs->beta_g2, d->beta_g2, 192); std::memcpy(&crs->delta_g2, d->delta_g2, 192);
    uint32_t sb, snb, sv, snv;
    shard_policy(ctx, idx, cnt, &sb, &snb, &sv, &snv);
    auto up = [&](int group, const void* pts, size_t len, bb_bases** dst) -> int {
        size_t lo = len * sb / snb, hi = len * (sb + 1) / snb;
        size_t stride = group == BB_G1 ? 96 : 192;
        int rc = bb_bases_upload(ctx, group, (const char*)pts + lo * stride, hi - lo, lo, len, dst);
        if (rc == BB_OK) { (*dst)->win_index = sv; (*dst)->win_count = snv; }
        return rc;
    };
    int s;
    if ((s = up(BB_G1, d->h, d->h_len, &crs->h)) || (s = up(BB_G1, d->l, d->l_len, &crs->l)) ||
        (s = up(BB_G1, d->a, d->a_len, &crs->a)) || (s = up(BB_G1, d->b_g1, d->b_g1_len, &crs->b_g1)) ||
        (s = up(BB_G2, d->b_g2, d->b_g2_len, &crs->b_g2))) {
        bb_crs_destroy(crs);
        return s;
    }
    *out = crs;
    return BB_OK;
}

// Synthetic Parameters for the benchmark: every vector is [k_i]G with counter-based pseudorandom
// k_i, generated on the device directly into this rank's base-range shard.  Same vector lengths a
// real key has; not a trapdoor CRS (parity with a valid CRS at this size lives in tests/).
int bb_synth_crs(bb_ctx* ctx, uint64_t seed, size_t h_len, size_t l_len, size_t a_len, size_t b_len,
                 uint32_t shard_index, uint32_t shard_count, bb_crs** out) {
    if (!ctx || !out || !shard_count || shard_index >= shard_count) return BB_ERR_ARG;
    bb_crs* crs = new bb_crs();
    crs->ctx = ctx; crs->shard_index = shard_index; crs->shard_count = shard_count;
    uint64_t k1[3][4], k2[2][4];
    for (int i = 0; i < 3; i++) { k1[i][0] = seed * 7 + i + 11; k1[i][1] = seed + 3 * i; k1[i][2] = 5 + i; k1[i][3] = 1 + i; }
    for (int i = 0; i < 2; i++) { k2[i][0] = seed * 13 + i + 17; k2[i][1] = seed + 5 * i; k2[i][2] = 9 + i; k2[i][3] = 2 + i; }
    G1Affine v1[3]; G2Affine v2[2];
    int s = bb_fixed_base_mul(ctx, BB_G1, k1, 3, BB_FORM_CANONICAL, v1);
    if (s == BB_OK) s = bb_fixed_base_mul(ctx, BB_G2, k2, 2, BB_FORM_CANONICAL, v2);
    crs->alpha_g1 = v1[0]; crs->beta_g1 = v1[1]; crs->delta_g1 = v1[2];
    crs->beta_g2 = v2[0]; crs->delta_g2 = v2[1];
    uint32_t sb, snb, sv, snv;
    shard_policy(ctx, shard_index, shard_count, &sb, &snb, &sv, &snv);
    auto gen = [&](int group, uint64_t sd, size_t len, bb_bases** dst) -> int {
        size_t lo = len * sb / snb, hi = len * (sb + 1) / snb;
        int rc = bb_synth_bases(ctx, group, sd, hi - lo, lo, len, dst);
        if (rc == BB_OK) { (*dst)->win_index = sv; (*dst)->win_count = snv; }
        return rc;
    };
    if (s == BB_OK) s = gen(BB_G1, seed + 1, h_len, &crs->h);
    if (s == BB_OK) s = gen(BB_G1, seed + 2, l_len, &crs->l);
    if (s == BB_OK) s = gen(BB_G1, seed + 3, a_len, &crs->a);
    if (s == BB_OK) s = gen(BB_G1, seed + 4, b_len, &crs->b_g1);
    if (s == BB_OK) s = gen(BB_G2, seed + 4, b_len, &crs->b_g2);       // same dlogs in G1 and G2, like a real B query
    if (s != BB_OK) { bb_crs_destroy(crs); return s; }
    *out = crs;
    return BB_OK;
}

void bb_crs_destroy(bb_crs* crs) {
    if (!crs) return;
    bb_bases_free(crs->h); bb_bases_free(crs->l); bb_bases_free(crs->a); bb_bases_free(crs->b_g1); bb_bases_free(crs->b_g2);
    delete crs;
}

}  // extern "C"

namespace {

// x_i / zz_i, y_i / zzz_i for a handful of points with ONE field inversion (Montgomery's trick)
template <class F>
void batch_to_affine(const XYZZ<F>* in, int n, Affine<F>* out) {
    F pre[8];
    F acc = FieldOps<F>::one();
    for (int i = 0; i < n; i++) { pre[i] = acc; if (!in[i].is_identity()) acc = acc * in[i].ZZZ; }
    F inv = FieldOps<F>::inv(acc);
    for (int i = n - 1; i >= 0; i--) {
        if (in[i].is_identity()) { out[i] = Affine<F>::identity(); continue; }
        F zi3 = inv * pre[i];
        inv = inv * in[i].ZZZ;
        F zi2 = (zi3 * in[i].ZZ).sqr();
        out[i] = {in[i].X * zi2, in[i].Y * zi3};
    }
}

bool delta_is_identity(const bb_crs* crs) {                                       // prover.rs:320-324
    if (!crs->delta_g1.is_identity() && !crs->delta_g2.is_identity()) return false;
    set_error("delta is the identity: subversion-CRS attack");
    return true;
}

}  // namespace

// One proof in flight: what prove_begin queued and prove_end collects.
struct bb_prove {
    bb_ctx* ctx = nullptr;
    const bb_crs* crs = nullptr;
    bb_witness w{};
    size_t m = 1, b_in_total = 0;
    uint32_t log_m = 0;
    DevBuf d_a, d_b, d_c, d_tmp, d_in, d_aux;
    cudaStream_t up = nullptr;
    cudaEvent_t ev_up = nullptr, ev_h = nullptr;
    bb_msm_job* jobs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int s = BB_OK;
    bool h_queued = false;
    ~bb_prove() {
        // every exit -- error paths included -- first drains every job and the two streams that touch the buffers
        // above (the DevBufs go back to the cache when this object dies), then destroys the events
        for (auto& j : jobs) if (j) { MsmResult r; msm_wait_result(j, &r); j = nullptr; }
        if (ctx) cudaStreamSynchronize(ctx->main_stream);
        if (up) cudaStreamSynchronize(up);
        if (ev_up) cudaEventDestroy(ev_up);
        if (ev_h) cudaEventDestroy(ev_h);
    }
};

namespace {

static const char* const job_tag[8] = {"h", "l", "a_inputs", "a_aux", "b_g1_inputs", "b_g1_aux", "b_g2_inputs", "b_g2_aux"};

void prove_start_job(bb_prove* P, int slot, const bb_bases* bases, size_t off, const uint64_t* dens, size_t dens_len, const void* d_sc, size_t cnt, cudaEvent_t ev) {
    if (P->s == BB_OK)
        P->s = msm_start(P->ctx, bases, off, dens, dens_len, d_sc, true, cnt, BB_FORM_MONTGOMERY, ev, &P->jobs[slot], job_tag[slot], slot == 0 ? 1 : slot == 7 ? 2 : 0);
}

// The H pipeline (prover.rs:221-240) on the high-priority main stream and the h MSM behind it.  With `evals`
// (device buffers of m coset evaluations of a, b, c -- computed elsewhere, bb_h_coset_evals) only the last
// transform remains: mul_assign, sub_assign, divide_by_z_on_coset, icoset_fft (:232-237).
int prove_queue_h(bb_prove* P, const void* const* evals) {
    NvtxRange range(evals ? "bb: H pipeline (last transform) + h MSM" : "bb: H pipeline + h MSM");
    bb_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->main_stream;
    const bb_witness* w = &P->w;
    const size_t n = w->n_constraints, m = P->m;
    const cudaMemcpyKind kind = w->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int s = BB_OK;
    if (evals) {
        s = h_poly_final_device(ctx, st, (Fr*)evals[0], (const Fr*)evals[1], (const Fr*)evals[2], P->d_tmp.as<Fr>(), P->log_m);
    } else {
        auto stage = [&](DevBuf& d, const void* src) -> int {
            if (m > n) BB_CUDA(cudaMemsetAsync((char*)d.p + n * 32, 0, (m - n) * 32, st));   // coeffs.resize(m, zero), domain.rs:69
            if (n) BB_CUDA(cudaMemcpyAsync(d.p, src, n * 32, kind, st));
            return BB_OK;
        };
        if (!w->on_device) ctx->h2d_bytes += 3 * n * 32;
        if ((s = stage(P->d_a, w->a)) == BB_OK && (s = stage(P->d_b, w->b)) == BB_OK && (s = stage(P->d_c, w->c)) == BB_OK)
            s = h_poly_device(ctx, st, P->d_a.as<Fr>(), P->d_b.as<Fr>(), P->d_c.as<Fr>(), P->d_tmp.as<Fr>(), P->log_m);
    }
    if (s == BB_OK && cudaEventRecord(P->ev_h, st) != cudaSuccess) { set_error("cudaEventRecord(H pipeline done) failed"); s = BB_ERR_CUDA; }
    if (s != BB_OK) { P->s = s; return s; }
    prove_start_job(P, 0, P->crs->h, 0, nullptr, 0, evals ? evals[0] : P->d_a.p, m - 1, P->ev_h);             // :238-244
    P->h_queued = true;
    return P->s;
}

// Queues the uploads and the MSMs of one proof (prover.rs:221-318).  h_inline: the H pipeline and the h MSM are
// queued here, FIRST -- they are the longest dependency chain of the proof and run on high-priority streams, the
// seven witness MSMs fill the machine around them.  Otherwise prove_end queues them (multi-GPU: the coset
// evaluations arrive from other ranks while the witness MSMs already run).
int prove_begin_impl(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, bool h_inline, bb_prove** out) {
    NvtxRange range("bb: prove_begin (uploads, queue MSMs)");
    if (!ctx || !crs || !w || !out) { set_error("bb_groth16_prove: null argument"); return BB_ERR_ARG; }
    BB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = w->n_constraints;
    size_t m = 1;
    uint32_t log_m = 0;
    while (m < n) {                                   // from_coeffs, domain.rs:49-60
        m *= 2;
        log_m += 1;
        if (log_m >= (uint32_t)bbc::FR_S) { set_error("PolynomialDegreeTooLarge"); return BB_ERR_POLYNOMIAL_DEGREE_TOO_LARGE; }
    }
    cudaStream_t st = ctx->main_stream;
    if (ctx->opt_profile) {                           // origin of the per-job device timeline (bb_profile_read "tl.<job>.<mark>")
        if (!ctx->epoch_ev) BB_CUDA(cudaEventCreate(&ctx->epoch_ev));
        BB_CUDA(cudaEventRecord(ctx->epoch_ev, st));
    }
    std::unique_ptr<bb_prove> P(new bb_prove());
    P->ctx = ctx; P->crs = crs; P->w = *w; P->m = m; P->log_m = log_m;
    if (h_inline) { BB_TRY(P->d_a.alloc(ctx, m * 32)); BB_TRY(P->d_b.alloc(ctx, m * 32)); BB_TRY(P->d_c.alloc(ctx, m * 32)); }
    BB_TRY(P->d_tmp.alloc(ctx, m * 32));
    BB_TRY(P->d_in.alloc(ctx, w->n_inputs * 32)); BB_TRY(P->d_aux.alloc(ctx, w->n_aux * 32));
    P->up = ctx->pick_stream();
    const cudaMemcpyKind kind = w->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (w->n_inputs) BB_CUDA(cudaMemcpyAsync(P->d_in.p, w->input_assignment, w->n_inputs * 32, kind, P->up));
    if (w->n_aux) BB_CUDA(cudaMemcpyAsync(P->d_aux.p, w->aux_assignment, w->n_aux * 32, kind, P->up));
    if (!w->on_device) ctx->h2d_bytes += (w->n_inputs + w->n_aux) * 32;
    BB_CUDA(cudaEventCreateWithFlags(&P->ev_up, cudaEventDisableTiming));
    BB_CUDA(cudaEventCreateWithFlags(&P->ev_h, cudaEventDisableTiming));
    BB_CUDA(cudaEventRecord(P->ev_up, P->up));
    for (size_t j = 0; j < (w->n_inputs + 63) / 64; j++) {       // get_total_density, prover.rs:288-291
        uint64_t wv = w->b_input_density[j];
        if (j == (w->n_inputs + 63) / 64 - 1 && (w->n_inputs & 63)) wv &= (1ull << (w->n_inputs & 63)) - 1ull;
        P->b_in_total += (size_t)__builtin_popcountll(wv);
    }
    bb_prove* p = P.get();
    if (h_inline) prove_queue_h(p, nullptr);
    const cudaEvent_t ev_up = P->ev_up;
    const size_t b_in_total = P->b_in_total;
    // all eight MSMs are in flight before the first wait (prover.rs:244-318); the G2 MSM has the longest chain
    // of the seven (Fp2 arithmetic): it goes first, on the other high-priority stream
    prove_start_job(p, 7, crs->b_g2, b_in_total, w->b_aux_density, w->n_aux, P->d_aux.p, w->n_aux, ev_up);      // :318
    prove_start_job(p, 1, crs->l, 0, nullptr, 0, P->d_aux.p, w->n_aux, ev_up);                                  // :263-268
    prove_start_job(p, 2, crs->a, 0, nullptr, 0, P->d_in.p, w->n_inputs, ev_up);                                // :275-280
    prove_start_job(p, 3, crs->a, w->n_inputs, w->a_aux_density, w->n_aux, P->d_aux.p, w->n_aux, ev_up);        // :281-286
    prove_start_job(p, 4, crs->b_g1, 0, w->b_input_density, w->n_inputs, P->d_in.p, w->n_inputs, ev_up);        // :296-301
    prove_start_job(p, 5, crs->b_g1, b_in_total, w->b_aux_density, w->n_aux, P->d_aux.p, w->n_aux, ev_up);      // :302-307
    prove_start_job(p, 6, crs->b_g2, 0, w->b_input_density, w->n_inputs, P->d_in.p, w->n_inputs, ev_up);        // :312-317
    *out = P.release();
    return BB_OK;                                      // a failure to queue is kept in the state and reported by prove_end
}

// Queues the H pipeline if prove_begin did not, waits for the eight MSMs and writes the 960-byte partial sums.
// Frees the state.  `while_device_runs`, if given, is called once everything is queued and before the first
// wait: host work that needs no MSM result goes there and overlaps the device.
int prove_end_impl(bb_prove* Praw, const void* const* evals, uint8_t* partials, const std::function<void()>* while_device_runs) {
    NvtxRange range("bb: prove_end (waits, folds)");
    std::unique_ptr<bb_prove> P(Praw);
    if (!Praw || !partials) { set_error("bb_groth16_prove_end: null argument"); return BB_ERR_ARG; }
    bb_ctx* ctx = P->ctx;
    const bb_crs* crs = P->crs;
    BB_CUDA(cudaSetDevice(ctx->device));
    if (!P->h_queued && P->s == BB_OK) {
        if (!evals) { BB_TRY(P->d_a.alloc(ctx, P->m * 32)); BB_TRY(P->d_b.alloc(ctx, P->m * 32)); BB_TRY(P->d_c.alloc(ctx, P->m * 32)); }
        prove_queue_h(P.get(), evals);
    }
    int s = P->s;
    if (while_device_runs && *while_device_runs) (*while_device_runs)();
    // wait() x8 (prover.rs:339-354); always drain every started job.  Each wait ends with a host-side Horner
    // fold of the job's window sums (255 doublings: 0.2 ms for G1, 0.5 ms for G2 on one core); jobs tend to
    // finish together, so the eight waits run on eight short-lived host threads and the folds overlap.
    G1X sums1[6];
    G2X sums2[2];
    for (auto& p1 : sums1) p1 = G1X::identity();
    for (auto& p2 : sums2) p2 = G2X::identity();
    int slot_status[8] = {BB_OK, BB_OK, BB_OK, BB_OK, BB_OK, BB_OK, BB_OK, BB_OK};
    {
        std::thread waiters[8];
        for (int k = 0; k < 8; k++) {
            if (!P->jobs[k]) continue;
            bb_msm_job* job = P->jobs[k];
            P->jobs[k] = nullptr;
            waiters[k] = std::thread([&, k, job] {
                cudaSetDevice(ctx->device);
                MsmResult r;
                slot_status[k] = msm_wait_result(job, &r);
                if (slot_status[k] == BB_OK) {
                    if (k >= 6) sums2[k - 6] = r.x2;
                    else sums1[k] = r.g1;
                }
            });
        }
        for (auto& t : waiters) if (t.joinable()) t.join();
    }
    if (s == BB_OK) {
        // Which error create_proof reports when several apply: the density assert fires inside the
        // multiexp() call itself (multiexp.rs:324-329; calls are made in slot order), the delta
        // check comes before the first wait (prover.rs:320-324), and the waits run a_inputs, a_aux,
        // b_g1_inputs, b_g1_aux, b_g2_inputs, b_g2_aux, h, l (:339-354).
        static const int ref_wait_order[8] = {2, 3, 4, 5, 6, 7, 0, 1};
        for (int k = 0; k < 8 && s == BB_OK; k++) if (slot_status[k] == BB_ERR_DENSITY_MISMATCH) s = slot_status[k];
        if (s == BB_OK && delta_is_identity(crs)) s = BB_ERR_UNEXPECTED_IDENTITY;
        for (int idx = 0; idx < 8 && s == BB_OK; idx++) s = slot_status[ref_wait_order[idx]];
        // the message of the error that is reported (another job may have failed later)
        if (s == BB_ERR_IO_UNEXPECTED_EOF) set_error("expected more bases from source");
        else if (s == BB_ERR_UNEXPECTED_IDENTITY && !crs->delta_g1.is_identity() && !crs->delta_g2.is_identity())
            set_error("encountered an identity element in the CRS");
        else if (s == BB_ERR_DENSITY_MISMATCH) set_error("density map length differs from the number of exponents");
    }
    if (s == BB_OK) {
        G1Affine a1[6];
        G2Affine a2[2];
        batch_to_affine<Fp>(sums1, 6, a1);
        batch_to_affine<Fp2>(sums2, 2, a2);
        std::memcpy(partials, a1, 6 * 96);
        std::memcpy(partials + 576, a2, 2 * 192);
    }
    return s;
}

int prove_partials_impl(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, uint8_t* partials, const std::function<void()>* while_device_runs) {
    if (!partials) { set_error("bb_groth16_prove_partials: null argument"); return BB_ERR_ARG; }
    bb_prove* P = nullptr;
    BB_TRY(prove_begin_impl(ctx, crs, w, true, &P));
    return prove_end_impl(P, nullptr, partials, while_device_runs);
}

// The terms of A, B, C that depend only on the key and on (r, s) -- prover.rs:326-337:
//   g_a = delta_g1 r + alpha_g1,  g_b = delta_g2 s + beta_g2,  g_c = delta_g1 rs + alpha_g1 s + beta_g1 r
struct ProofStatic { G1X g_a, g_c; G2X g_b; };

void finalize_static(const bb_crs* crs, const Fr& r, const Fr& s, ProofStatic* out) {
    Fr rs = fr_to_canonical(fr_from_canonical(r) * fr_from_canonical(s));        // rs.mul_assign(&s), :332-333
    out->g_a = g1_host_mul(G1X::from_affine(crs->delta_g1), r.l);                 // :326-327
    out->g_a.add_mixed(crs->alpha_g1);
    out->g_b = g2_host_mul(G2X::from_affine(crs->delta_g2), s.l);                 // :328-329
    out->g_b.add_mixed(crs->beta_g2);
    out->g_c = g1_host_mul(G1X::from_affine(crs->delta_g1), rs.l);                // :335-337
    out->g_c.add(g1_host_mul(G1X::from_affine(crs->alpha_g1), s.l));
    out->g_c.add(g1_host_mul(G1X::from_affine(crs->beta_g1), r.l));
}

int finalize_impl(const bb_crs* crs, const uint8_t* partials, size_t count, const Fr& r, const Fr& s, const ProofStatic& stat, uint8_t* proof) {
    G1X sum1[6];
    G2X sum2[2];
    for (auto& p : sum1) p = G1X::identity();
    for (auto& p : sum2) p = G2X::identity();
    for (size_t c = 0; c < count; c++) {
        const uint8_t* base = partials + c * BB_PARTIALS_BYTES;
        for (int k = 0; k < 6; k++) { G1Affine a; std::memcpy(&a, base + 96 * k, 96); sum1[k].add_mixed(a); }
        for (int k = 0; k < 2; k++) { G2Affine a; std::memcpy(&a, base + 576 + 192 * k, 192); sum2[k].add_mixed(a); }
    }
    G1X g_a = stat.g_a, g_c = stat.g_c;
    G2X g_b = stat.g_b;
    G1X a_answer = sum1[2];                                                       // :339-343
    a_answer.add(sum1[3]);
    g_a.add(a_answer);
    G1X b1_answer = sum1[4];                                                      // :345-354
    b1_answer.add(sum1[5]);
    // the two scalar multiplications that need MSM results sit between the last device result and the proof
    // bytes: they run side by side (0.2 ms each on one core)
    G1X a_s;
    std::thread side([&] { a_s = g1_host_mul(a_answer, s.l); });
    const G1X b_r = g1_host_mul(b1_answer, r.l);
    G2X b2_answer = sum2[0];
    b2_answer.add(sum2[1]);
    g_b.add(b2_answer);
    side.join();
    g_c.add(a_s);
    g_c.add(b_r);
    g_c.add(sum1[0]);
    g_c.add(sum1[1]);
    G1X ac[2] = {g_a, g_c};                                                       // :356-360: to_affine, one inversion for A and C
    G1Affine ac_aff[2];
    batch_to_affine<Fp>(ac, 2, ac_aff);
    g1_compress(ac_aff[0], proof);                                                // lib.rs:39-45
    g2_compress(g_b.to_affine(), proof + 48);
    g1_compress(ac_aff[1], proof + 144);
    return BB_OK;
}

}  // namespace

extern "C" {

int bb_groth16_prove_partials(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, uint8_t* partials) {
    return prove_partials_impl(ctx, crs, w, partials, nullptr);
}

int bb_groth16_prove_begin(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, bb_prove** out) {
    return prove_begin_impl(ctx, crs, w, false, out);
}

int bb_groth16_prove_end(bb_prove* state, const void* d_evals_a, const void* d_evals_b, const void* d_evals_c, uint8_t* partials) {
    const bool any = d_evals_a || d_evals_b || d_evals_c;
    if (any && !(d_evals_a && d_evals_b && d_evals_c)) { delete state; set_error("bb_groth16_prove_end: all three evaluation vectors or none"); return BB_ERR_ARG; }
    const void* ev[3] = {d_evals_a, d_evals_b, d_evals_c};
    return prove_end_impl(state, any ? ev : nullptr, partials, nullptr);
}

// from_coeffs (zero padding to m), ifft and coset_fft of ONE of the polynomials a, b, c (prover.rs:225-230) into a
// device buffer of m Fr.  _async queues the copy and the two transforms on the high-priority stream and returns;
// _wait blocks until every queued evaluation is there (and hands the scratch buffers back).
int bb_h_coset_evals_async(bb_ctx* ctx, const void* poly, size_t n_constraints, int on_device, void* d_out) {
    if (!ctx || !d_out || (n_constraints && !poly)) { set_error("bb_h_coset_evals: null argument"); return BB_ERR_ARG; }
    BB_CUDA(cudaSetDevice(ctx->device));
    size_t m = 1;
    uint32_t log_m = 0;
    while (m < n_constraints) {
        m *= 2;
        log_m += 1;
        if (log_m >= (uint32_t)bbc::FR_S) { set_error("PolynomialDegreeTooLarge"); return BB_ERR_POLYNOMIAL_DEGREE_TOO_LARGE; }
    }
    cudaStream_t st = ctx->main_stream;
    void* d_tmp = nullptr;
    BB_TRY(ctx->alloc(m * 32, &d_tmp));
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        ctx->h_evals_scratch.push_back(d_tmp);            // released by bb_h_coset_evals_wait, after the stream has drained
    }
    if (m > n_constraints) BB_CUDA(cudaMemsetAsync((char*)d_out + n_constraints * 32, 0, (m - n_constraints) * 32, st));
    if (n_constraints) BB_CUDA(cudaMemcpyAsync(d_out, poly, n_constraints * 32, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!on_device) ctx->h2d_bytes += n_constraints * 32;
    return h_poly_evals_device(ctx, st, (Fr*)d_out, (Fr*)d_tmp, log_m);
}

int bb_h_coset_evals_wait(bb_ctx* ctx) {
    if (!ctx) { set_error("bb_h_coset_evals_wait: null argument"); return BB_ERR_ARG; }
    BB_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaStreamSynchronize(ctx->main_stream);
    std::vector<void*> scratch;
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        scratch.swap(ctx->h_evals_scratch);
    }
    for (void* p : scratch) ctx->release(p);
    if (e != cudaSuccess) { set_error("bb_h_coset_evals: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
    return BB_OK;
}

int bb_h_coset_evals(bb_ctx* ctx, const void* poly, size_t n_constraints, int on_device, void* d_out) {
    int s = bb_h_coset_evals_async(ctx, poly, n_constraints, on_device, d_out);
    int w = ctx ? bb_h_coset_evals_wait(ctx) : BB_OK;
    return s != BB_OK ? s : w;
}

int bb_groth16_finalize(const bb_crs* crs, const uint8_t* partials, size_t count, const uint8_t* r_bytes, const uint8_t* s_bytes,
                        uint8_t* proof) {
    if (!crs || !partials || !count || !r_bytes || !s_bytes || !proof) { set_error("bb_groth16_finalize: null argument"); return BB_ERR_ARG; }
    if (delta_is_identity(crs)) return BB_ERR_UNEXPECTED_IDENTITY;
    Fr r, s;
    std::memcpy(r.l, r_bytes, 32);
    std::memcpy(s.l, s_bytes, 32);
    ProofStatic stat;
    finalize_static(crs, r, s, &stat);
    return finalize_impl(crs, partials, count, r, s, stat, proof);
}

static_assert(sizeof(ProofStatic) == BB_PROOF_STATIC_BYTES, "ProofStatic layout");

int bb_groth16_finalize_static(const bb_crs* crs, const uint8_t* r_bytes, const uint8_t* s_bytes, uint8_t* static_out) {
    if (!crs || !r_bytes || !s_bytes || !static_out) { set_error("bb_groth16_finalize_static: null argument"); return BB_ERR_ARG; }
    Fr r, s;
    std::memcpy(r.l, r_bytes, 32);
    std::memcpy(s.l, s_bytes, 32);
    ProofStatic stat;
    finalize_static(crs, r, s, &stat);
    std::memcpy(static_out, &stat, sizeof stat);
    return BB_OK;
}

int bb_groth16_finalize_with(const bb_crs* crs, const uint8_t* partials, size_t count, const uint8_t* r_bytes, const uint8_t* s_bytes,
                             const uint8_t* static_in, uint8_t* proof) {
    if (!crs || !partials || !count || !r_bytes || !s_bytes || !static_in || !proof) { set_error("bb_groth16_finalize_with: null argument"); return BB_ERR_ARG; }
    if (delta_is_identity(crs)) return BB_ERR_UNEXPECTED_IDENTITY;
    Fr r, s;
    std::memcpy(r.l, r_bytes, 32);
    std::memcpy(s.l, s_bytes, 32);
    ProofStatic stat;
    std::memcpy(&stat, static_in, sizeof stat);
    return finalize_impl(crs, partials, count, r, s, stat, proof);
}

// ---- tuning: measure, don't guess ---------------------------------------------------------------------
// The MSM has forms whose ranking depends on the machine, the key size and the number of shards (resident
// window multiples trade 16x the base storage and colder gathers for fewer buckets and deeper halving rounds).
// A key lives for many proofs, so the choice is made once per key by timing real proofs of each form on this
// device (the way FFT and convolution libraries plan), and a form is only eligible if its partial sums are
// byte-identical to the default form's.
namespace {
struct Tuning { const char* name; long precompute; long rows_log; long groups; };   // groups: 1 = G1 vectors, 2 = G2 vectors
const Tuning kTunings[] = {
    {"per-window bucket sets, no tables", 0, 3, 3},
    {"one bucket set over resident window multiples, halving rounds down to ~8 rows per bucket", 2, 3, 3},
    {"one bucket set over resident window multiples, halving rounds down to ~16 rows per bucket", 2, 4, 3},
    {"one bucket set over resident window multiples, halving rounds down to ~4 rows per bucket", 2, 2, 3},
    {"one bucket set over resident window multiples, halving rounds down to ~32 rows per bucket", 2, 5, 3},
    {"one bucket set over resident window multiples for the G2 vector only (~8 rows), per-window bucket sets for G1", 2, 3, 2},
    {"one bucket set over resident window multiples for the G1 vectors only (~8 rows), per-window bucket sets for G2", 2, 3, 1},
};
constexpr int kNumTunings = (int)(sizeof kTunings / sizeof kTunings[0]);
}  // namespace

int bb_tuning_count(void) { return kNumTunings; }
const char* bb_tuning_name(int index) { return index >= 0 && index < kNumTunings ? kTunings[index].name : nullptr; }

namespace {
// tables for the vectors of the groups in `groups` (bit 0 = G1, bit 1 = G2); the other vectors lose theirs
int crs_tables(bb_ctx* ctx, bb_crs* crs, long groups) {
    for (bb_bases* b : {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2}) {
        if (!b) continue;
        if (groups & (b->group == BB_G1 ? 1 : 2)) BB_TRY(bases_build_table(ctx, b));
        else bb_bases_drop_table(b);
    }
    return BB_OK;
}
}  // namespace

int bb_crs_precompute(bb_ctx* ctx, bb_crs* crs) {
    if (!ctx || !crs || crs->ctx != ctx) { set_error("bb_crs_precompute: bad argument"); return BB_ERR_ARG; }
    return crs_tables(ctx, crs, 3);
}
int bb_crs_drop_tables(bb_crs* crs) {
    if (!crs) { set_error("bb_crs_drop_tables: null argument"); return BB_ERR_ARG; }
    for (bb_bases* b : {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2})
        if (b) bb_bases_drop_table(b);
    return BB_OK;
}

int bb_crs_apply_tuning(bb_ctx* ctx, bb_crs* crs, int index) {
    if (!ctx || !crs || crs->ctx != ctx || index < 0 || index >= kNumTunings) { set_error("bb_crs_apply_tuning: bad argument"); return BB_ERR_ARG; }
    BB_CUDA(cudaSetDevice(ctx->device));
    BB_CUDA(cudaDeviceSynchronize());                  // nothing may still read a table that is about to go
    const Tuning& t = kTunings[index];
    int s = BB_OK;
    if (t.precompute) s = crs_tables(ctx, crs, t.groups);
    if (s != BB_OK || !t.precompute) bb_crs_drop_tables(crs);
    ctx->opt_msm_precompute = s == BB_OK ? t.precompute : 0;
    ctx->opt_msm_precompute_groups = s == BB_OK ? t.groups : 3;
    ctx->opt_msm_unified_rows_log = s == BB_OK ? t.rows_log : 3;
    return s;
}

int bb_groth16_autotune(bb_ctx* ctx, bb_crs* crs, const bb_witness* w, int reps, int* chosen, double* ms_out) {
    if (!ctx || !crs || crs->ctx != ctx || !w) { set_error("bb_groth16_autotune: bad argument"); return BB_ERR_ARG; }
    if (reps < 1) reps = 3;
    uint8_t ref[BB_PARTIALS_BYTES], got[BB_PARTIALS_BYTES];
    double ms[kNumTunings];
    int best = 0;
    // A shard of a window-sharded key owns the windows w = index mod count of its base range, and the window size differs
    // between the forms (the table forms take it from the vector, the default from the job): the (base, window) pairs a
    // shard sums -- and with them its partial sums -- are comparable only after all shards are added up.  For a sharded
    // key this call therefore only MEASURES (every form must still prove without an error) and leaves the default form
    // configured: all shards must run the same form, which is a collective decision (distributed.autotune_sharded).
    const bool whole_key = crs->shard_count == 1;
    for (int i = 0; i < kNumTunings; i++) {
        ms[i] = -1.0;                                   // -1: not available here (the tables do not fit)
        int s = bb_crs_apply_tuning(ctx, crs, i);
        if (s != BB_OK) { if (i == 0) return s; continue; }
        std::memset(got, 0, sizeof got);
        s = prove_partials_impl(ctx, crs, w, got, nullptr);             // first proof of this form: builds caches, checked
        if (s != BB_OK) {
            if (i == 0) return s;                       // the witness / key itself fails: nothing to tune
            ms[i] = -2.0;                               // -2: this form failed where the default did not
            continue;
        }
        if (i == 0) std::memcpy(ref, got, sizeof ref);
        else if (whole_key && std::memcmp(ref, got, sizeof ref) != 0) { ms[i] = -3.0; continue; }     // -3: different partial sums -- never eligible
        double fastest = 1e300;
        for (int k = 0; k < reps && s == BB_OK; k++) {
            const auto t0 = std::chrono::steady_clock::now();
            s = prove_partials_impl(ctx, crs, w, got, nullptr);
            const double dt = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (dt < fastest) fastest = dt;
        }
        if (s != BB_OK) { if (i == 0) return s; ms[i] = -2.0; continue; }
        ms[i] = fastest;
        if (ms[i] < ms[best]) best = i;
    }
    if (ms_out) for (int i = 0; i < kNumTunings; i++) ms_out[i] = ms[i];
    if (chosen) *chosen = best;
    return bb_crs_apply_tuning(ctx, crs, whole_key ? best : 0);
}

int bb_groth16_prove(bb_ctx* ctx, const bb_crs* crs, const bb_witness* w, const uint8_t* r_bytes, const uint8_t* s_bytes, uint8_t* proof) {
    if (!crs || !r_bytes || !s_bytes || !proof) { set_error("bb_groth16_prove: null argument"); return BB_ERR_ARG; }
    if (crs->shard_count != 1) { set_error("bb_groth16_prove needs an unsharded CRS; use prove_partials + finalize"); return BB_ERR_ARG; }
    uint8_t partials[BB_PARTIALS_BYTES];
    std::memset(partials, 0, sizeof partials);
    Fr r, s;
    std::memcpy(r.l, r_bytes, 32);
    std::memcpy(s.l, s_bytes, 32);
    // the five scalar multiplications that need no MSM result run on the host while the device works
    ProofStatic stat;
    const bool bad_delta = crs->delta_g1.is_identity() || crs->delta_g2.is_identity();
    // profile mode: host-side milestones of one prove, ms since entry (bb_profile_read "host.<mark>")
    const auto t_entry = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    double t_queued = 0, t_static = 0;
    std::function<void()> overlap = [&] {
        t_queued = since();
        if (!bad_delta) finalize_static(crs, r, s, &stat);
        t_static = since();
    };
    BB_TRY(prove_partials_impl(ctx, crs, w, partials, &overlap));     // reports delta = identity as well (prover.rs:320-324)
    const double t_waited = since();
    int rc = finalize_impl(crs, partials, 1, r, s, stat, proof);
    if (ctx && ctx->opt_profile) {
        ctx->prof_add("host.queued", t_queued, 1, 0);          // everything launched
        ctx->prof_add("host.static_done", t_static, 1, 0);     // MSM-independent scalar multiplications done
        ctx->prof_add("host.msms_done", t_waited, 1, 0);       // all eight results folded on the host
        ctx->prof_add("host.proof_done", since(), 1, 0);
    }
    return rc;
}

}  // extern "C"
