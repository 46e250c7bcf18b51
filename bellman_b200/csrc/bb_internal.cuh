// bellman_b200 internal: context, device memory cache, error plumbing.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bellman_b200_diag.h"
#include "curve.cuh"

// NVTX ranges (header-only nvtx3, no link dependency; a no-op unless a tool such as nsys is attached):
// one range per prove, per MSM job launch sequence and per H pipeline, so that a timeline shows the job graph.
#if defined(__CUDACC__)
#include <nvtx3/nvToolsExt.h>
namespace bb {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace bb
#else
namespace bb {
struct NvtxRange { explicit NvtxRange(const char*) {} };
}  // namespace bb
#endif

namespace bb {

void set_error(const char* fmt, ...);

#define BB_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            bb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return BB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define BB_TRY(call)                       \
    do {                                   \
        int s__ = (call);                  \
        if (s__ != BB_OK) return s__;      \
    } while (0)

struct NttTables;

}  // namespace bb

// Context = one device.  Device allocations are cached by size so that the steady state of a
// prover issues no cudaMalloc/cudaFree (both synchronise the device).
struct bb_ctx {
    int device = 0;
    int num_sms = 0;
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::map<void*, size_t> live_blocks;
    std::vector<cudaStream_t> streams;       // round-robin pool for async jobs
    size_t next_stream = 0;
    cudaStream_t main_stream = nullptr;      // high priority: NTTs / the H pipeline
    cudaStream_t crit_stream[2] = {nullptr, nullptr};   // high priority: the MSMs on a proof's critical path (h, and the G2 job)
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0};   // host<->device traffic of the hot-path calls
    long opt_msm_window_bits = 0;
    long opt_ntt_tile_log = 11;
    long opt_ntt_col_bits = 3;
    long opt_ntt_radix8 = 0;          // 1 = register radix-8 windows (k_ntt_pass8); measured 4-5 % slower than the radix-2 sweeps on B200
                                      // (2^24 fwd+inv 9.74 vs 9.34 ms, profiles/ab_r02_call4.txt), so it stays an option
    long opt_profile = 0;
    long opt_msm_acc_variant = 0;     // accepted and ignored: the launch-bound / prefetch variants of round 1's accumulate kernel
                                      // measured no gain (profiles/ab_r02_call1...) and were removed
    long opt_msm_reduce_2d = 1;      // bucket reduction through row / column sums (k_bucket_fold) for windows of >= 1024 buckets
    long opt_msm_reduce_k = 4;       // entries per thread and level of the bucket reduction: 4 halves the length of the
    long opt_msm_reduce_k1 = 4;      // dependent-addition chain of 16 (2K per level, log_K D levels) for 1.25x its additions
    long opt_msm_big_cap = 0;
    long opt_shard_windows = 4;      // multi-GPU: up to this many window shards per base range (1 = base ranges only)
    long opt_msm_affine_rounds = -1;  // batched-affine halving rounds per MSM: -1 = by size, 0 = none (XYZZ accumulation only)
    long opt_msm_affine_tma = 0;      // dense halving rounds of G1 jobs: operands staged by cp.async.bulk + mbarrier (k_aff_phase3_tma)
    long opt_msm_affine_batch = 16;   // pairs per thread sharing one link of the inversion chain
    long opt_msm_precompute = 0;     // resident window multiples 2^(c w) P of every base vector (msm.cu: bases_build_table):
                                     // 1 = one bucket array per window slot, 2 = ONE bucket array for all windows
    long opt_msm_precompute_groups = 3;  // which groups use the tables: bit 0 = G1 vectors, bit 1 = G2 vectors
    long opt_msm_unified_rows_log = 3;   // msm_precompute = 2: halving rounds until about 2^this rows per bucket are left
    struct ProfEntry { double ms = 0; uint64_t launches = 0, units = 0; };
    std::map<std::string, ProfEntry> prof;
    void prof_add(const char* what, double ms, uint64_t launches, uint64_t units) {
        std::lock_guard<std::mutex> g(mu);
        ProfEntry& e = prof[what];
        e.ms += ms; e.launches += launches; e.units += units;
    }
    std::map<uint32_t, bb::NttTables*> ntt_tables;   // by log_n
    std::vector<void*> h_evals_scratch;              // transform scratch of bb_h_coset_evals_async calls not yet waited for
    cudaEvent_t epoch_ev = nullptr;                  // profile mode: start of the current prove (device timeline origin)

    // small page-locked staging blocks for results (cudaMallocHost/cudaFreeHost synchronise the
    // device and are slow; jobs borrow fixed-size blocks instead)
    static constexpr size_t PINNED_BLOCK = 64 * 1024;
    std::vector<void*> pinned_free;
    int pinned_acquire(size_t bytes, void** out);
    void pinned_release(void* p);
    int alloc(size_t bytes, void** out);
    void release(void* p);
    cudaStream_t pick_stream();
    void count_launch(uint64_t k = 1) { launches.fetch_add(k, std::memory_order_relaxed); }
};

struct bb_bases {
    bb_ctx* ctx;
    int group;
    void* d_points;          // Affine<Fp> or Affine<Fp2>
    size_t n;                // points held here
    size_t global_offset;    // first global index held
    size_t global_len;       // length of the whole (unsharded) vector
    uint32_t win_index = 0;  // window shard: this device accumulates windows w % win_count == win_index
    uint32_t win_count = 1;
    // Optional table of window multiples: slot s = w / win_count of every owned window w holds
    // 2^(tab_c * w) * P_i at d_table[s * n + i] (affine).  With it all windows share one bucket set.
    void* d_table = nullptr;
    uint32_t tab_c = 0, tab_W = 0, tab_slots = 0;
    std::mutex tab_mu;
};

namespace bb {

// RAII device buffer from the context cache
struct DevBuf {
    bb_ctx* ctx = nullptr;
    void* p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { reset(); }
    int alloc(bb_ctx* c, size_t bytes) { reset(); ctx = c; return c->alloc(bytes ? bytes : 16, &p); }
    void reset() { if (p && ctx) ctx->release(p); p = nullptr; }
    template <class T> T* as() const { return (T*)p; }
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- ntt.cu ----
int ntt_run_device(bb_ctx* ctx, cudaStream_t st, Fr* d_data, Fr* d_tmp, uint32_t log_n, int mode);
int h_poly_device(bb_ctx* ctx, cudaStream_t st, Fr* d_a, Fr* d_b, Fr* d_c, Fr* d_tmp, uint32_t log_m);
int h_poly_evals_device(bb_ctx* ctx, cudaStream_t st, Fr* d_p, Fr* d_tmp, uint32_t log_m);
int h_poly_final_device(bb_ctx* ctx, cudaStream_t st, Fr* d_a, const Fr* d_b, const Fr* d_c, Fr* d_tmp, uint32_t log_m);
int fr_convert_device(bb_ctx* ctx, cudaStream_t st, Fr* d_data, size_t n, bool to_montgomery);
void ntt_free_tables(bb_ctx* ctx);
int domain_pointwise_device(bb_ctx* ctx, cudaStream_t st, Fr* d_a, const Fr* d_b, size_t n, int op, const Fr& k);

// ---- capi.cu ----
int fixed_base_mul_device(bb_ctx* ctx, int group, const Fr* d_scalars, size_t n, bool montgomery, void* d_out, cudaStream_t st);

// ---- msm.cu ----
struct MsmResult {
    int status = BB_OK;
    bool g2 = false;
    G1X g1;
    G2X x2;
};
int msm_start(bb_ctx* ctx, const bb_bases* bases, size_t base_offset, const uint64_t* density_bits, size_t density_len,
              const void* scalars, bool scalars_on_device, size_t n, int form, cudaEvent_t wait_for, bb_msm_job** out,
              const char* tag = nullptr, int critical = 0);
int msm_wait_result(bb_msm_job* job, MsmResult* res);
int bases_build_table(bb_ctx* ctx, bb_bases* bases);
// vals[i] <- 1 / vals[i] for n NON-ZERO field elements in HBM (Montgomery's trick, fan-in 32, one Fermat inversion);
// scratch: batch_invert_scratch(n) elements
int batch_invert_fp(bb_ctx* ctx, cudaStream_t st, Fp* vals, size_t n, Fp* scratch);
int batch_invert_fp2(bb_ctx* ctx, cudaStream_t st, Fp2* vals, size_t n, Fp2* scratch);
size_t batch_invert_scratch(size_t n);

}  // namespace bb
