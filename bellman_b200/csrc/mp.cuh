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

namespace bb {

#if defined(__CUDA_ARCH__)
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
#endif  // __CUDA_ARCH__

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
#if defined(__CUDA_ARCH__)
        return Cfg::dmod(k);
#else
        return Cfg::hmod()[k];
#endif
    }

    BB_HD Fe operator+(const Fe& o) const {
        Fe r;
#if defined(__CUDA_ARCH__)
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
#if defined(__CUDA_ARCH__)
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
#if defined(__CUDA_ARCH__)
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
    BB_HD Fe sqr() const { return *this * *this; }
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
