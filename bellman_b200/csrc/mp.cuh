// bellman_b200: multi-precision Montgomery arithmetic for sm_100a.
//
// Replaces the field arithmetic bellman's hot path pulls from the ff / bls12_381
// crates (call sites: /root/reference/src/domain.rs:250-258 mul/add/sub_assign,
// src/multiexp.rs:39 mixed add).  Elements are N little-endian 32-bit limbs in
// Montgomery form with R = 2^(32N) -- byte-identical to bls12_381's [u64; N/2]
// Montgomery limbs on a little-endian host.
//
// Device path: the product a*b*R^-1 is computed row by row with TWO accumulators,
// one collecting the limb products a[j]*b_i for even j and one for odd j.  Inside
// one accumulator the (lo,hi) halves of consecutive products land on consecutive
// limbs, so a whole row is a single mad.lo.cc / madc.hi.cc carry chain that ptxas
// can pair into IMAD.WIDE; the Montgomery reduction row m*p uses the same two
// chains, and the per-row division by 2^32 is a swap of accumulator roles plus a
// static two-limb register rename (no data movement).  The instruction-level model
// of this routine is checked in tools/emu_montmul.py.
//
// Host path (finalisation only: window fold, to_affine, encodings): portable
// 64-bit CIOS.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#define BB_D __device__ __forceinline__
// out-of-line on the device: keeps the G2 / reduction kernels (dozens of field products per
// point operation) from being inlined into multi-megabyte functions
#define BB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BB_HD inline
#define BB_D inline
#define BB_HD_NOINLINE inline
#endif

// Compile-time arithmetic variants.  Results are bit-identical in all of them; each is executed on
// the CPU by tests/test_emulated_device_field.py and was run through the GPU parity suite on B200.
//   BB_FP_WIDE_SQR  dedicated 12-limb squaring (wide_sqr + redc_wide) instead of a*a:  -29 % multiplier
//                   instructions per squaring.
//   BB_FP2_LAZY     lazily reduced Fp2 product (3 wide products, 2 reductions): -13 % multiplier
//                   instructions, +9 % instructions overall.
// Both were measured SLOWER on B200 in the 2^20 prove (50.7 / 50.5 ms against 48.9 ms, same box, round 1):
// the bucket kernels run 4-8 warps per SM and are bound by the latency of the serial carry chains,
// which both variants lengthen (24-limb add/shift/diagonal chains), not by multiplier issue slots.
// They stay off; with both at 0 the generated SASS is identical to the plain merged product.
#ifndef BB_FP_WIDE_SQR
#define BB_FP_WIDE_SQR 0
#endif
#ifndef BB_FP2_LAZY
#define BB_FP2_LAZY 0
#endif

namespace bb {

// BB_EMULATE_PTX (host compilers only, tests/native/): the carry-chain primitives below are
// modelled in plain C++ with an explicit carry flag, so that the exact limb/index logic of the
// device path can be executed and checked on a CPU.  Never defined in the product build.
#if defined(BB_EMULATE_PTX) && !defined(__CUDACC__)
#define BB_DEVPATH 1
namespace ptx {
static thread_local uint32_t cf = 0;
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin) { uint64_t s = (uint64_t)a + b + cin; cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin) { uint64_t d = (uint64_t)a - b - bin; cf = (uint32_t)(d >> 63); return (uint32_t)d; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, 0); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, cf); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cf); }
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, cf); }
inline uint32_t addc(uint32_t a, uint32_t b) { uint32_t c = cf; uint32_t r = a + b + c; return r; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, cf); }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf; }
}  // namespace ptx
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi) << 32 | lo) << (s & 31)) >> 32); }
#elif defined(__CUDA_ARCH__)
#define BB_DEVPATH 1
namespace ptx {
BB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
}  // namespace ptx
#endif

#if defined(BB_DEVPATH)
// E (aligned to limb 0) += even-index limb products of x*s ; fresh carry chain
template <int N, class XF>
BB_D void cmad_even(uint32_t* E, XF x, uint32_t s) {
    E[0] = ptx::mad_lo_cc(x(0), s, E[0]);
    E[1] = ptx::madc_hi_cc(x(0), s, E[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
        E[k] = ptx::madc_lo_cc(x(k), s, E[k]);
        E[k + 1] = ptx::madc_hi_cc(x(k), s, E[k + 1]);
    }
    E[N] = ptx::addc(E[N], 0);
}
// O (aligned to limb 1) += odd-index limb products of x*s ; fresh carry chain
template <int N, class XF>
BB_D void cmad_odd(uint32_t* O, XF x, uint32_t s) {
    O[0] = ptx::mad_lo_cc(x(1), s, O[0]);
    O[1] = ptx::madc_hi_cc(x(1), s, O[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
        O[k] = ptx::madc_lo_cc(x(k + 1), s, O[k]);
        O[k + 1] = ptx::madc_hi_cc(x(k + 1), s, O[k + 1]);
    }
    O[N] = ptx::addc(O[N], 0);
}

// One CIOS row.  On entry (E,O) hold the previous row's accumulators with E[0]==0;
// logically T/2^32 = O + (E >> 32).  On exit the roles are swapped: O is the new
// limb-0-aligned accumulator, E (shifted down two limbs in place) the odd one.
template <int N, class AF, class PF>
BB_D void mont_row(uint32_t* E, uint32_t* O, AF a, uint32_t bi, PF p, uint32_t inv) {
    O[0] = ptx::add_cc(O[0], E[1]);                       // fold E[1] into limb 0; carry -> limb 1
#pragma unroll
    for (int k = 0; k < N; k += 2) {                      // E := (E >> 64) + odd products + carry
        E[k] = ptx::madc_lo_cc(a(k + 1), bi, (k + 2 <= N) ? E[k + 2] : 0u);
        E[k + 1] = ptx::madc_hi_cc(a(k + 1), bi, (k + 3 <= N) ? E[k + 3] : 0u);
    }
    E[N] = ptx::addc(0, 0);
    cmad_even<N>(O, a, bi);
    uint32_t m = O[0] * inv;
    cmad_odd<N>(E, p, m);
    cmad_even<N>(O, p, m);
}
// ---- wide (2N-limb) squaring and stand-alone Montgomery reduction -------------------------
// a^2 needs only the N(N-1)/2 off-diagonal limb products (doubled) plus the N diagonal ones:
// 78 instead of 144 for Fp.  The products a[j]*a[i] are split by the parity of i+j into two
// accumulators (B is offset by one limb) so that, inside a row, the (lo,hi) halves of one class
// sit on consecutive limbs and form a single carry chain -- the same trick as in the merged
// multiplication above.  Instruction-level model: tools/emu_wide.py.
template <int N>
BB_D void wide_sqr(uint32_t* T, const uint32_t* a) {
    uint32_t A[2 * N + 1], B[2 * N + 1];
#pragma unroll
    for (int k = 0; k <= 2 * N; k++) { A[k] = 0; B[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
        // class 0: j = i+2, i+4, ... (i+j even) -> A at limbs i+j, i+j+1
        if (i + 2 < N) {
            A[2 * i + 2] = ptx::mad_lo_cc(a[i + 2], a[i], A[2 * i + 2]);
            A[2 * i + 3] = ptx::madc_hi_cc(a[i + 2], a[i], A[2 * i + 3]);
#pragma unroll
            for (int j = i + 4; j < N; j += 2) {
                A[i + j] = ptx::madc_lo_cc(a[j], a[i], A[i + j]);
                A[i + j + 1] = ptx::madc_hi_cc(a[j], a[i], A[i + j + 1]);
            }
            // last j of the class
            const int jl = i + 2 + 2 * ((N - 1 - (i + 2)) / 2);
            A[i + jl + 2] = ptx::addc(A[i + jl + 2], 0);
        }
        // class 1: j = i+1, i+3, ... (i+j odd) -> B at limbs i+j-1, i+j
        if (i + 1 < N) {
            B[2 * i] = ptx::mad_lo_cc(a[i + 1], a[i], B[2 * i]);
            B[2 * i + 1] = ptx::madc_hi_cc(a[i + 1], a[i], B[2 * i + 1]);
#pragma unroll
            for (int j = i + 3; j < N; j += 2) {
                B[i + j - 1] = ptx::madc_lo_cc(a[j], a[i], B[i + j - 1]);
                B[i + j] = ptx::madc_hi_cc(a[j], a[i], B[i + j]);
            }
            const int jl = i + 1 + 2 * ((N - 1 - (i + 1)) / 2);
            B[i + jl + 1] = ptx::addc(B[i + jl + 1], 0);
        }
    }
    // off = A + (B << 32)
    T[0] = A[0];
    T[1] = ptx::add_cc(A[1], B[0]);
#pragma unroll
    for (int k = 2; k < 2 * N; k++) T[k] = ptx::addc_cc(A[k], B[k - 1]);
    // 2 * off
#pragma unroll
    for (int k = 2 * N - 1; k >= 1; k--) T[k] = __funnelshift_l(T[k - 1], T[k], 1);
    T[0] <<= 1;
    // + diagonal
    T[0] = ptx::mad_lo_cc(a[0], a[0], T[0]);
    T[1] = ptx::madc_hi_cc(a[0], a[0], T[1]);
#pragma unroll
    for (int i = 1; i < N; i++) {
        T[2 * i] = ptx::madc_lo_cc(a[i], a[i], T[2 * i]);
        T[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], T[2 * i + 1]);
    }
}

// One row of the reduction-only CIOS (mont_row without the a*b products)
template <int N, class PF>
BB_D void redc_row(uint32_t* E, uint32_t* O, PF p, uint32_t inv) {
    O[0] = ptx::add_cc(O[0], E[1]);
#pragma unroll
    for (int k = 0; k < N; k++) E[k] = ptx::addc_cc((k + 2 <= N) ? E[k + 2] : 0u, 0u);
    E[N] = ptx::addc(0, 0);
    uint32_t m = O[0] * inv;
    cmad_odd<N>(E, p, m);
    cmad_even<N>(O, p, m);
}

// T[2N] = a * b (plain integer product of two N-limb values), same two-accumulator layout
template <int N>
BB_D void wide_mul(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t A[2 * N + 1], B[2 * N + 1];
#pragma unroll
    for (int k = 0; k <= 2 * N; k++) { A[k] = 0; B[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
        {   // i+j even -> A at limbs i+j, i+j+1
            const int j0 = i & 1;
            A[i + j0] = ptx::mad_lo_cc(a[j0], b[i], A[i + j0]);
            A[i + j0 + 1] = ptx::madc_hi_cc(a[j0], b[i], A[i + j0 + 1]);
#pragma unroll
            for (int j = j0 + 2; j < N; j += 2) {
                A[i + j] = ptx::madc_lo_cc(a[j], b[i], A[i + j]);
                A[i + j + 1] = ptx::madc_hi_cc(a[j], b[i], A[i + j + 1]);
            }
            const int jl = j0 + 2 * ((N - 1 - j0) / 2);
            A[i + jl + 2] = ptx::addc(A[i + jl + 2], 0);
        }
        {   // i+j odd -> B at limbs i+j-1, i+j
            const int j0 = 1 - (i & 1);
            B[i + j0 - 1] = ptx::mad_lo_cc(a[j0], b[i], B[i + j0 - 1]);
            B[i + j0] = ptx::madc_hi_cc(a[j0], b[i], B[i + j0]);
#pragma unroll
            for (int j = j0 + 2; j < N; j += 2) {
                B[i + j - 1] = ptx::madc_lo_cc(a[j], b[i], B[i + j - 1]);
                B[i + j] = ptx::madc_hi_cc(a[j], b[i], B[i + j]);
            }
            const int jl = j0 + 2 * ((N - 1 - j0) / 2);
            B[i + jl + 1] = ptx::addc(B[i + jl + 1], 0);
        }
    }
    T[0] = A[0];
    T[1] = ptx::add_cc(A[1], B[0]);
#pragma unroll
    for (int k = 2; k < 2 * N; k++) T[k] = ptx::addc_cc(A[k], B[k - 1]);
}

// r = T * R^-1 mod p for a 2N-limb T < p*R with T >> 32N < p (true for T < 2 p^2 with both BLS12-381
// moduli): REDC(T) = (T >> 32N) + redc(T mod R), one conditional subtraction.
template <class Cfg>
BB_D void redc_wide(uint32_t* r, const uint32_t* T) {
    constexpr int N = Cfg::N;
    uint32_t X[N + 1], Y[N + 1];
    auto p = [&](int k) { return Cfg::dmod(k); };
#pragma unroll
    for (int k = 0; k < N; k += 2) { X[k] = T[k]; X[k + 1] = 0; Y[k] = T[k + 1]; Y[k + 1] = 0; }
    X[N] = 0;
    Y[N] = 0;
    {
        uint32_t m = X[0] * Cfg::INV;
        cmad_odd<N>(Y, p, m);
        cmad_even<N>(X, p, m);
    }
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        redc_row<N>(X, Y, p, Cfg::INV);                   // now even = Y, odd = X
        if (i + 1 < N) redc_row<N>(Y, X, p, Cfg::INV);    // back to even = X
    }
    r[0] = ptx::add_cc(X[0], Y[1]);                       // redc(T mod R) <= p
#pragma unroll
    for (int k = 1; k < N; k++) r[k] = ptx::addc_cc(X[k], Y[k + 1]);
    r[0] = ptx::add_cc(r[0], T[N]);                       // + high half: sum < 2p, no carry out
#pragma unroll
    for (int k = 1; k < N; k++) r[k] = ptx::addc_cc(r[k], T[N + k]);
    uint32_t t[N];
    t[0] = ptx::sub_cc(r[0], Cfg::dmod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r[i], Cfg::dmod(i));
    uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}
#endif  // BB_DEVPATH

// Cfg supplies: N, INV (32-bit -p^-1), device accessor dmod(k) (constant memory),
// host pointer hmod() to the same limbs.
template <class Cfg>
struct alignas(16) Fe {
    static constexpr int N = Cfg::N;
    uint32_t l[N];

    BB_HD static Fe zero() { Fe r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    BB_HD bool is_zero() const { uint32_t a = 0; for (int i = 0; i < N; i++) a |= l[i]; return a == 0; }
    BB_HD bool operator==(const Fe& o) const { uint32_t a = 0; for (int i = 0; i < N; i++) a |= l[i] ^ o.l[i]; return a == 0; }
    BB_HD bool operator!=(const Fe& o) const { return !(*this == o); }

    BB_HD static uint32_t modl(int k) {
#if defined(BB_DEVPATH)
        return Cfg::dmod(k);
#else
        return Cfg::hmod()[k];
#endif
    }

    BB_HD Fe operator+(const Fe& o) const {
        Fe r;
#if defined(BB_DEVPATH)
        r.l[0] = ptx::add_cc(l[0], o.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = ptx::addc_cc(l[i], o.l[i]);
        // both moduli leave the top bit of limb N-1 clear: no carry out of a+b
        uint32_t t[N];
        t[0] = ptx::sub_cc(r.l[0], Cfg::dmod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r.l[i], Cfg::dmod(i));
        uint32_t borrow = ptx::subc(0, 0);                 // 0xffffffff if r < p
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : t[i];
#else
        uint64_t c = 0;
        for (int i = 0; i < N; i++) { c += (uint64_t)l[i] + o.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
        uint32_t t[N];
        int64_t b = 0;
        for (int i = 0; i < N; i++) { b += (int64_t)r.l[i] - Cfg::hmod()[i]; t[i] = (uint32_t)b; b >>= 32; }
        if (b == 0) for (int i = 0; i < N; i++) r.l[i] = t[i];
#endif
        return r;
    }
    BB_HD Fe operator-(const Fe& o) const {
        Fe r;
#if defined(BB_DEVPATH)
        r.l[0] = ptx::sub_cc(l[0], o.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(l[i], o.l[i]);
        uint32_t mask = ptx::subc(0, 0);                   // 0xffffffff if a < b
        r.l[0] = ptx::add_cc(r.l[0], Cfg::dmod(0) & mask);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], Cfg::dmod(i) & mask);
        r.l[N - 1] = ptx::addc(r.l[N - 1], Cfg::dmod(N - 1) & mask);
#else
        int64_t b = 0;
        for (int i = 0; i < N; i++) { b += (int64_t)l[i] - o.l[i]; r.l[i] = (uint32_t)b; b >>= 32; }
        if (b) {
            uint64_t c = 0;
            for (int i = 0; i < N; i++) { c += (uint64_t)r.l[i] + Cfg::hmod()[i]; r.l[i] = (uint32_t)c; c >>= 32; }
        }
#endif
        return r;
    }
    BB_HD Fe neg() const { return is_zero() ? *this : (zero() - *this); }
    BB_HD Fe dbl() const { return *this + *this; }

    BB_HD Fe operator*(const Fe& o) const {
        Fe r;
#if defined(BB_DEVPATH)
        uint32_t X[N + 1], Y[N + 1];
        auto a = [&](int k) { return l[k]; };
        auto p = [&](int k) { return Cfg::dmod(k); };
        const uint32_t b0 = o.l[0];
#pragma unroll
        for (int k = 0; k < N; k += 2) {
            X[k] = ptx::mul_lo(l[k], b0);
            X[k + 1] = ptx::mul_hi(l[k], b0);
            Y[k] = ptx::mul_lo(l[k + 1], b0);
            Y[k + 1] = ptx::mul_hi(l[k + 1], b0);
        }
        X[N] = 0;
        Y[N] = 0;
        {
            uint32_t m = X[0] * Cfg::INV;
            cmad_odd<N>(Y, p, m);
            cmad_even<N>(X, p, m);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            mont_row<N>(X, Y, a, o.l[i], p, Cfg::INV);               // now even = Y, odd = X
            if (i + 1 < N) mont_row<N>(Y, X, a, o.l[i + 1], p, Cfg::INV);   // back to even = X
        }
        // N even: after row N-1 the limb-0-aligned accumulator is Y (Y[0]==0), odd is X
        r.l[0] = ptx::add_cc(X[0], Y[1]);
#pragma unroll
        for (int k = 1; k < N; k++) r.l[k] = ptx::addc_cc(X[k], Y[k + 1]);
        uint32_t t[N];
        t[0] = ptx::sub_cc(r.l[0], Cfg::dmod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r.l[i], Cfg::dmod(i));
        uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : t[i];
#else
        constexpr int M = N / 2;
        typedef unsigned __int128 u128;
        uint64_t a[M], b[M], q[M], t[M + 2];
        std::memcpy(a, l, sizeof a);
        std::memcpy(b, o.l, sizeof b);
        std::memcpy(q, Cfg::hmod(), sizeof q);
        const uint64_t inv64 = Cfg::INV64;
        for (int i = 0; i < M + 2; i++) t[i] = 0;
        for (int i = 0; i < M; i++) {
            uint64_t c = 0;
            for (int j = 0; j < M; j++) { u128 s = (u128)a[j] * b[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
            u128 s = (u128)t[M] + c;
            t[M] = (uint64_t)s;
            t[M + 1] = (uint64_t)(s >> 64);
            uint64_t m = t[0] * inv64;
            u128 s2 = (u128)m * q[0] + t[0];
            c = (uint64_t)(s2 >> 64);
            for (int j = 1; j < M; j++) { s2 = (u128)m * q[j] + t[j] + c; t[j - 1] = (uint64_t)s2; c = (uint64_t)(s2 >> 64); }
            s2 = (u128)t[M] + c;
            t[M - 1] = (uint64_t)s2;
            t[M] = t[M + 1] + (uint64_t)(s2 >> 64);
        }
        bool ge = t[M] != 0;
        if (!ge) {
            ge = true;
            for (int i = M - 1; i >= 0; i--) { if (t[i] > q[i]) break; if (t[i] < q[i]) { ge = false; break; } }
        }
        if (ge) {
            uint64_t bo = 0;
            for (int i = 0; i < M; i++) { u128 d = (u128)t[i] - q[i] - bo; t[i] = (uint64_t)d; bo = (uint64_t)(d >> 64) & 1; }
        }
        std::memcpy(r.l, t, sizeof r.l);
#endif
        return r;
    }
    // device: dedicated squaring (wide_sqr + redc_wide) for the 12-limb field, where it saves a
    // quarter of the multiplier work; the 8-limb field keeps the merged product
    BB_HD Fe sqr() const {
#if defined(BB_DEVPATH) && BB_FP_WIDE_SQR
        if (N >= 12) {
            uint32_t T[2 * N];
            wide_sqr<N>(T, l);
            Fe r;
            redc_wide<Cfg>(r.l, T);
            return r;
        }
#endif
        return *this * *this;
    }
    BB_HD Fe& operator+=(const Fe& o) { *this = *this + o; return *this; }
    BB_HD Fe& operator-=(const Fe& o) { *this = *this - o; return *this; }
    BB_HD Fe& operator*=(const Fe& o) { *this = *this * o; return *this; }

    // variable-time pow by a little-endian u32-limb exponent
    BB_HD Fe pow(const uint32_t* e, int ne, const Fe& one) const {
        Fe res = one;
        for (int i = ne - 1; i >= 0; i--)
            for (int b = 31; b >= 0; b--) {
                res = res.sqr();
                if ((e[i] >> b) & 1) res = res * *this;
            }
        return res;
    }
    BB_HD Fe pow_u64(uint64_t e, const Fe& one) const {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(w, 2, one);
    }
};

}  // namespace bb
