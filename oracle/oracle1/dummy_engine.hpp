// Oracle-1 toy engine (TEST INFRASTRUCTURE ONLY).  Restates the reference's
// test-only DummyEngine (groth16/src/tests/dummy_engine.rs): Fr = Z/64513
// (:15), S = 10, generator 5, root of unity 57751 (:297-320), G1 = G2 = Fr under
// addition with scalar-mul = field multiplication (:336-378).  It carries the
// only golden vectors the reference has for generator + domain + multiexp +
// prover (groth16/src/tests/mod.rs:91-373).
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace o1 {

struct DFr {
    static constexpr uint32_t Q = 64513;
    static constexpr uint32_t S = 10;
    static constexpr uint32_t NUM_BITS = 16;
    uint32_t v;
    static DFr zero() { return {0}; }
    static DFr one() { return {1}; }
    static DFr from_u64(uint64_t x) { return {(uint32_t)(x % Q)}; }
    static DFr root_of_unity() { return {57751}; }
    static DFr generator() { return {5}; }
    bool is_zero() const { return v == 0; }
    bool operator==(const DFr& o) const { return v == o.v; }
    DFr operator+(const DFr& o) const { return {(v + o.v) % Q}; }
    DFr operator-(const DFr& o) const { return {(v + Q - o.v) % Q}; }
    DFr operator*(const DFr& o) const { return {(uint32_t)((uint64_t)v * o.v % Q)}; }
    DFr square() const { return *this * *this; }
    DFr neg() const { return {(Q - v) % Q}; }
    DFr pow(uint64_t e) const {
        DFr r = one(), b = *this;
        while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
        return r;
    }
    DFr inv() const { return pow(Q - 2); }
    std::array<uint64_t, 4> to_bits() const { return {v, 0, 0, 0}; }
};

// additive group on DFr; affine and projective coincide
struct DG {
    DFr e;
    static DG identity() { return {DFr::zero()}; }
    static DG from_affine(const DG& a) { return a; }
    bool is_identity() const { return e.is_zero(); }
    DG add(const DG& o) const { return {e + o.e}; }
    DG add_mixed(const DG& o) const { return {e + o.e}; }
    DG dbl() const { return {e + e}; }
    DG mul(const DFr& k) const { return {e * k}; }
    DG to_affine() const { return *this; }
    bool operator==(const DG& o) const { return e == o.e; }
};

struct DummyEngine {
    typedef DFr Fr;
    typedef DG G1;
    typedef DG G1A;
    typedef DG G2;
    typedef DG G2A;
    static void batch_normalize1(const std::vector<DG>& in, DG* out) { for (size_t i = 0; i < in.size(); i++) out[i] = in[i]; }
    static void batch_normalize2(const std::vector<DG>& in, DG* out) { for (size_t i = 0; i < in.size(); i++) out[i] = in[i]; }
};

}  // namespace o1
