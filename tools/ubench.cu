// Micro-benchmarks that decide the arithmetic design: integer multiply-add issue rates and
// field-multiplication throughput on the device.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -std=c++17 -I bellman_b200/csrc tools/ubench.cu -o gpurun_out/ubench
#include <cstdio>
#include <vector>
#include "field.cuh"
using namespace bb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long w0 = x0, w1 = x1, w2 = x2, w3 = x3, w4 = x4, w5 = x5, w6 = x6, w7 = x7;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {   // IMAD lo
#define S(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
            S(x0) S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        } else if (MODE == 1) {  // IMAD.HI
#define S(x) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
            S(x0) S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        } else if (MODE == 2) {  // IMAD.WIDE
// multiplicand = low word of the accumulator: a true dependency, nothing can be hoisted
#define S(x, y) asm volatile("{ .reg .u32 t; cvt.u32.u64 t, %0; mad.wide.u32 %0, t, %1, %0; }" : "+l"(x) : "r"(a));
            S(w0, x0) S(w1, x1) S(w2, x2) S(w3, x3) S(w4, x4) S(w5, x5) S(w6, x6) S(w7, x7)
#undef S
        } else if (MODE == 3) {  // IADD3 chain
#define S(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
            S(x0) S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        } else if (MODE == 6) {  // mul.wide only (IMAD.WIDE.U32 with RZ addend), dependent through the low word
#define S(x) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %1; }" : "+l"(x) : "r"(a));
            S(w0) S(w1) S(w2) S(w3) S(w4) S(w5) S(w6) S(w7)
#undef S
        } else if (MODE == 5) {  // mad.lo.cc/madc.hi.cc pairs in one carry chain (IMAD.WIDE.U32.X)
            asm volatile("mad.lo.cc.u32 %0, %0, %8, %0; madc.hi.cc.u32 %1, %0, %8, %1; madc.lo.cc.u32 %2, %2, %8, %2; madc.hi.cc.u32 %3, %2, %8, %3;"
                         "madc.lo.cc.u32 %4, %4, %8, %4; madc.hi.cc.u32 %5, %4, %8, %5; madc.lo.cc.u32 %6, %6, %8, %6; madc.hi.u32 %7, %6, %8, %7;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a));
        } else if (MODE == 4) {  // IMAD + IADD3 interleaved (dual pipe)
#define S(x, y) asm volatile("mad.lo.u32 %0, %0, %2, %3; add.u32 %1, %1, %2;" : "+r"(x), "+r"(y) : "r"(a), "r"(b));
            S(x0, x4) S(x1, x5) S(x2, x6) S(x3, x7)
#undef S
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7);
}

template <class FE, int CHAINS>
__global__ void __launch_bounds__(256) k_fmul(FE* out, const FE* in, int iters) {
    FE x[CHAINS];
    FE y = in[threadIdx.x & 31];
    for (int c = 0; c < CHAINS; c++) x[c] = in[(threadIdx.x + c) & 31];
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = x[c] * y;
    FE acc = x[0];
    for (int c = 1; c < CHAINS; c++) acc = acc + x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class FE>
__global__ void __launch_bounds__(256) k_fadd(FE* out, const FE* in, int iters) {
    FE x = in[threadIdx.x & 31], y = in[(threadIdx.x + 1) & 31];
    for (int i = 0; i < iters; i++) { x = x + y; y = y - x; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

template <class K>
float time_kernel(K launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    uint32_t* d_out; CK(cudaMalloc(&d_out, 64 << 20));
    const int iters = 4096;
    const char* names[] = {"IMAD.lo", "IMAD.hi", "IMAD.WIDE", "IADD", "IMAD+IADD", "lo/hi.cc x4", "MUL.WIDE"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        int blocks = sms * 2, threads = warps * 32 / 2;
        float ms[7];
        ms[0] = time_kernel([&] { k_imad<0><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[1] = time_kernel([&] { k_imad<1><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[2] = time_kernel([&] { k_imad<2><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[3] = time_kernel([&] { k_imad<3><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[4] = time_kernel([&] { k_imad<4><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[5] = time_kernel([&] { k_imad<5><<<blocks, threads>>>(d_out, 3, 5, iters); });
        ms[6] = time_kernel([&] { k_imad<6><<<blocks, threads>>>(d_out, 3, 5, iters); });
        for (int m = 0; m < 7; m++) {
            double ops = (double)blocks * threads * iters * 8;   // per-thread instructions of the named kind (mode 4: 4 IMAD + 4 IADD)
            printf("warps/SM %2d  %-10s %8.3f ms  %7.1f Gop/s  %6.1f op/clk/SM @1.9GHz\n", warps, names[m], ms[m], ops / ms[m] / 1e6,
                   ops / (ms[m] * 1e-3) / sms / 1.9e9);
        }
    }
    // field multiplication throughput
    std::vector<uint32_t> h(32 * 12);
    for (size_t i = 0; i < h.size(); i++) h[i] = 0x01234567u * (uint32_t)(i + 1) >> 2;
    void* d_in; CK(cudaMalloc(&d_in, h.size() * 4)); CK(cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const int fit = 512;
    for (int tpb : {64, 128, 256}) {
        for (int bps : {1, 2, 4, 8}) {
            int blocks = sms * bps;
            float m1 = time_kernel([&] { k_fmul<Fp, 1><<<blocks, tpb>>>((Fp*)d_out, (const Fp*)d_in, fit); });
            float m2 = time_kernel([&] { k_fmul<Fp, 2><<<blocks, tpb>>>((Fp*)d_out, (const Fp*)d_in, fit); });
            float m3 = time_kernel([&] { k_fmul<Fr, 1><<<blocks, tpb>>>((Fr*)d_out, (const Fr*)d_in, fit); });
            float m4 = time_kernel([&] { k_fmul<Fr, 2><<<blocks, tpb>>>((Fr*)d_out, (const Fr*)d_in, fit); });
            float m5 = time_kernel([&] { k_fadd<Fp><<<blocks, tpb>>>((Fp*)d_out, (const Fp*)d_in, fit); });
            double n1 = (double)blocks * tpb * fit;
            printf("tpb %3d blocks/SM %d (warps/SM %2d): Fp mul %6.2f G/s (2 chains %6.2f)  Fr mul %6.2f G/s (2 chains %6.2f)  Fp add+sub pairs %6.2f G/s\n",
                   tpb, bps, tpb * bps / 32, n1 / m1 / 1e6, 2 * n1 / m2 / 1e6, n1 / m3 / 1e6, 2 * n1 / m4 / 1e6, n1 / m5 / 1e6);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
