// Oracle-1 field arithmetic.  TEST INFRASTRUCTURE ONLY: nothing in the product
// (bellman_b200/) may include, link or call anything under oracle/.
//
// CPU restatement of the arithmetic the reference obtains from the external
// crates ff 0.13.0 / bls12_381 0.8.0 (Cargo.lock:105-108,310-313; sources are
// not in /root/reference).  Montgomery representation, 64-bit limbs, R = 2^(64N)
// -- the representation those crates use internally -- written from the
// published CIOS algorithm.  Checked limb-for-limb against oracle/oracle0
// (Python integers) in tests/test_oracle1.py.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>

#include "bls_constants.hpp"

namespace o1 {

typedef unsigned __int128 u128;

// P supplies: N, MOD[N], R[N], R2[N], INV
template <class P>
struct Mont {
    static constexpr int N = P::N;
    uint64_t l[N];

    static Mont zero() { Mont r; std::memset(r.l, 0, sizeof r.l); return r; }
    static Mont one() { Mont r; std::memcpy(r.l, P::R, sizeof r.l); return r; }
    static Mont from_raw(const uint64_t* p) { Mont r; std::memcpy(r.l, p, sizeof r.l); return r; }

    bool is_zero() const { uint64_t a = 0; for (int i = 0; i < N; i++) a |= l[i]; return a == 0; }
    bool operator==(const Mont& o) const { return std::memcmp(l, o.l, sizeof l) == 0; }
    bool operator!=(const Mont& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > P::MOD[i]) return true;
            if (a[i] < P::MOD[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a[i] - P::MOD[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }

    Mont operator+(const Mont& o) const {
        Mont r; uint64_t carry = 0;
        for (int i = 0; i < N; i++) {
            u128 s = (u128)l[i] + o.l[i] + carry;
            r.l[i] = (uint64_t)s; carry = (uint64_t)(s >> 64);
        }
        if (carry || geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    Mont operator-(const Mont& o) const {
        Mont r; uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)l[i] - o.l[i] - borrow;
            r.l[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            uint64_t carry = 0;
            for (int i = 0; i < N; i++) {
                u128 s = (u128)r.l[i] + P::MOD[i] + carry;
                r.l[i] = (uint64_t)s; carry = (uint64_t)(s >> 64);
            }
        }
        return r;
    }
    Mont neg() const { return is_zero() ? *this : (zero() - *this); }
    Mont dbl() const { return *this + *this; }

    // CIOS Montgomery product  a*b*R^-1 mod p.  Both moduli leave the top bit of the top limb clear,
    // so the two inner loops of the textbook CIOS merge and no (N+2)-limb accumulator is needed
    // (the multiplication and the reduction row share one pass; result < 2p before the final
    // conditional subtraction).
    Mont operator*(const Mont& o) const {
        static_assert((P::MOD[N - 1] >> 63) == 0, "modulus must leave the top bit clear");
        uint64_t t[N];
        std::memset(t, 0, sizeof t);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) {
            u128 p = (u128)l[0] * o.l[i] + t[0];
            uint64_t A = (uint64_t)(p >> 64);
            uint64_t m = (uint64_t)p * P::INV;
            u128 q = (u128)m * P::MOD[0] + (uint64_t)p;
            uint64_t C = (uint64_t)(q >> 64);
#pragma GCC unroll 8
            for (int j = 1; j < N; j++) {
                p = (u128)l[j] * o.l[i] + t[j] + A;
                A = (uint64_t)(p >> 64);
                q = (u128)m * P::MOD[j] + (uint64_t)p + C;
                C = (uint64_t)(q >> 64);
                t[j - 1] = (uint64_t)q;
            }
            t[N - 1] = C + A;
        }
        Mont r; std::memcpy(r.l, t, sizeof r.l);
        if (geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    Mont square() const { return *this * *this; }
    Mont& operator+=(const Mont& o) { *this = *this + o; return *this; }
    Mont& operator-=(const Mont& o) { *this = *this - o; return *this; }
    Mont& operator*=(const Mont& o) { *this = *this * o; return *this; }

    // pow by a little-endian multi-limb exponent (pow_vartime)
    Mont pow(const uint64_t* e, int ne) const {
        Mont res = one();
        for (int i = ne - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                res = res.square();
                if ((e[i] >> b) & 1) res = res * *this;
            }
        return res;
    }
    Mont pow(uint64_t e) const { return pow(&e, 1); }
    Mont inv() const {                       // Fermat: a^(p-2); caller checks non-zero
        uint64_t e[N];
        std::memcpy(e, P::MOD, sizeof e);
        e[0] -= 2;                           // both moduli end in ...01 / ...ab: no borrow
        return pow(e, N);
    }

    static Mont from_canonical(const uint64_t* c) {   // c < p, plain integer limbs
        Mont a = from_raw(c), r2 = from_raw(P::R2);
        return a * r2;
    }
    static Mont from_u64(uint64_t v) { uint64_t c[N] = {0}; c[0] = v; return from_canonical(c); }
    void to_canonical(uint64_t* out) const {
        Mont o; std::memset(o.l, 0, sizeof o.l); o.l[0] = 1;
        Mont r = *this * o;
        std::memcpy(out, r.l, sizeof r.l);
    }
};

struct FrParams {
    static constexpr int N = 4;
    static constexpr const uint64_t* MOD = o1c::FR_MOD;
    static constexpr const uint64_t* R = o1c::FR_R;
    static constexpr const uint64_t* R2 = o1c::FR_R2;
    static constexpr uint64_t INV = o1c::FR_INV;
};
struct FpParams {
    static constexpr int N = 6;
    static constexpr const uint64_t* MOD = o1c::FP_MOD;
    static constexpr const uint64_t* R = o1c::FP_R;
    static constexpr const uint64_t* R2 = o1c::FP_R2;
    static constexpr uint64_t INV = o1c::FP_INV;
};

typedef Mont<FpParams> Fp;

// bls12_381::Scalar stand-in with the ff::PrimeField surface the path uses
struct Fr : Mont<FrParams> {
    typedef Mont<FrParams> B;
    static constexpr uint32_t S = o1c::FR_S;
    static constexpr uint32_t NUM_BITS = o1c::FR_NUM_BITS;
    Fr() {}
    Fr(const B& b) : B(b) {}
    static Fr zero() { return B::zero(); }
    static Fr one() { return B::one(); }
    static Fr from_u64(uint64_t v) { return B::from_u64(v); }
    static Fr root_of_unity() { return B::from_raw(o1c::FR_ROOT_OF_UNITY_M); }
    static Fr generator() { return B::from_raw(o1c::FR_GENERATOR_M); }
    Fr operator+(const Fr& o) const { return B::operator+(o); }
    Fr operator-(const Fr& o) const { return B::operator-(o); }
    Fr operator*(const Fr& o) const { return B::operator*(o); }
    Fr square() const { return B::square(); }
    Fr inv() const { return B::inv(); }
    Fr pow(uint64_t e) const { return B::pow(e); }
    Fr neg() const { return B::neg(); }
    // PrimeFieldBits::to_le_bits: canonical integer, little-endian limbs
    std::array<uint64_t, 4> to_bits() const { std::array<uint64_t, 4> r; to_canonical(r.data()); return r; }
};

// Fp2 = Fp[u]/(u^2+1)
struct Fp2 {
    Fp c0, c1;
    static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fp2& o) const { return !(*this == o); }
    Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fp2 operator*(const Fp2& o) const {
        Fp aa = c0 * o.c0, bb = c1 * o.c1;
        Fp t = (c0 + c1) * (o.c0 + o.c1);
        return {aa - bb, t - aa - bb};
    }
    Fp2 square() const { return *this * *this; }
    Fp2 dbl() const { return *this + *this; }
    Fp2 neg() const { return {c0.neg(), c1.neg()}; }
    Fp2 inv() const {
        Fp n = (c0.square() + c1.square()).inv();
        return {c0 * n, (c1 * n).neg()};
    }
    Fp2& operator+=(const Fp2& o) { *this = *this + o; return *this; }
    Fp2& operator-=(const Fp2& o) { *this = *this - o; return *this; }
    Fp2& operator*=(const Fp2& o) { *this = *this * o; return *this; }
};

}  // namespace o1
