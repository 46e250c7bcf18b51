// bellman_b200: bulk asynchronous copies (TMA engine, cp.async.bulk -> SASS UBLKCP) completing on shared-memory
// mbarriers -- the Hopper/Blackwell way to stage a contiguous tile without spending registers or issue slots of
// the compute warps on it.  Used by the dense rounds of the batched-affine accumulation (msm.cu:
// k_aff_phase3_tma), where every CTA walks contiguous 24 KB tiles of the previous round's output.
//
// Host compilers (tests/native: CUDA emulation) get synchronous stand-ins: a bulk copy is a memcpy that has
// completed when it returns, so waiting is a no-op.
#pragma once
#include <cstdint>
#include <cstring>

namespace bb {
namespace tma {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival + the number of bytes the copies issued next will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
#else
// Host stand-in with the mbarrier's phase semantics, so that a waiter that runs before the producer really
// waits (the CUDA emulation of tests/native runs the threads of a block as fibers and defines BB_EMU_YIELD).
// Low word of the barrier: completed phases; high word: bytes still expected in the current phase.
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { *bar += (uint64_t)bytes << 32; }
inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(smem_dst, gmem_src, bytes);
    *bar -= (uint64_t)bytes << 32;
    if ((*bar >> 32) == 0) *bar = (uint32_t)*bar + 1u;             // last byte of the phase delivered
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (((uint32_t)*bar & 1u) == parity) {
#ifdef BB_EMU_YIELD
        BB_EMU_YIELD();
#endif
    }
}
#endif

}  // namespace tma
}  // namespace bb
