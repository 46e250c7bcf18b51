// Oracle-1 curve arithmetic (TEST INFRASTRUCTURE ONLY, see field.hpp).
//
// Short-Weierstrass y^2 = x^3 + b, a = 0, Jacobian coordinates (the reference's
// bls12_381 crate uses homogeneous projective; any complete coordinate system
// yields the same affine point and therefore the same bytes).  Call sites the
// reference makes into this arithmetic: mixed add multiexp.rs:39, projective add
// :273-274, double :299, is_identity :63, Affine*Fr prover.rs:326-337,
// to_affine prover.rs:357-359, to_bytes groth16/src/lib.rs:40-42.
#pragma once
#include <vector>

#include "field.hpp"

namespace o1 {

template <class Fq>
struct Affine {
    Fq x, y;
    bool inf;
    static Affine identity() { return {Fq::zero(), Fq::one(), true}; }
    bool is_identity() const { return inf; }
    bool operator==(const Affine& o) const {
        if (inf || o.inf) return inf == o.inf;
        return x == o.x && y == o.y;
    }
    Affine neg() const { return {x, y.neg(), inf}; }
};

template <class Fq>
struct Jac {
    typedef Affine<Fq> A;
    Fq X, Y, Z;
    static Jac identity() { return {Fq::zero(), Fq::one(), Fq::zero()}; }
    static Jac from_affine(const A& a) { return a.inf ? identity() : Jac{a.x, a.y, Fq::one()}; }
    bool is_identity() const { return Z.is_zero(); }

    Jac dbl() const {                       // dbl-2009-l
        if (is_identity()) return *this;
        Fq A_ = X.square(), B_ = Y.square(), C_ = B_.square();
        Fq D_ = ((X + B_).square() - A_ - C_).dbl();
        Fq E_ = A_.dbl() + A_;
        Fq F_ = E_.square();
        Jac r;
        r.X = F_ - D_.dbl();
        r.Y = E_ * (D_ - r.X) - C_.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    Jac add(const Jac& o) const {           // add-2007-bl with special cases
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq Z1Z1 = Z.square(), Z2Z2 = o.Z.square();
        Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return identity();
        }
        Fq H = U2 - U1, I = H.dbl().square(), J = H * I;
        Fq r_ = (S2 - S1).dbl(), V = U1 * I;
        Jac r;
        r.X = r_.square() - J - V.dbl();
        r.Y = r_ * (V - r.X) - (S1 * J).dbl();
        r.Z = ((Z + o.Z).square() - Z1Z1 - Z2Z2) * H;
        return r;
    }
    Jac add_mixed(const A& o) const {       // madd-2007-bl with special cases
        if (o.inf) return *this;
        if (is_identity()) return from_affine(o);
        Fq Z1Z1 = Z.square();
        Fq U2 = o.x * Z1Z1, S2 = o.y * Z * Z1Z1;
        if (X == U2) {
            if (Y == S2) return dbl();
            return identity();
        }
        Fq H = U2 - X, HH = H.square(), I = HH.dbl().dbl(), J = H * I;
        Fq r_ = (S2 - Y).dbl(), V = X * I;
        Jac r;
        r.X = r_.square() - J - V.dbl();
        r.Y = r_ * (V - r.X) - (Y * J).dbl();
        r.Z = (Z + H).square() - Z1Z1 - HH;
        return r;
    }
    Jac neg() const { return {X, Y.neg(), Z}; }
    A to_affine() const {
        if (is_identity()) return A::identity();
        Fq zi = Z.inv(), zi2 = zi.square();
        return {X * zi2, Y * zi2 * zi, false};
    }
    bool operator==(const Jac& o) const { return to_affine() == o.to_affine(); }

    // scalar given as canonical little-endian 256-bit integer
    Jac mul_bits(const uint64_t* k) const {
        Jac acc = identity();
        for (int i = 3; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                acc = acc.dbl();
                if ((k[i] >> b) & 1) acc = acc.add(*this);
            }
        return acc;
    }
    Jac mul(const Fr& k) const { auto b = k.to_bits(); return mul_bits(b.data()); }
};

// Curve::batch_normalize (generator.rs:293,419-422): Montgomery's trick
template <class Fq>
void batch_to_affine(const std::vector<Jac<Fq>>& in, Affine<Fq>* out) {
    size_t n = in.size();
    std::vector<Fq> pre(n);
    Fq acc = Fq::one();
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!in[i].is_identity()) acc = acc * in[i].Z;
    }
    Fq inv = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (in[i].is_identity()) { out[i] = Affine<Fq>::identity(); continue; }
        Fq zi = inv * pre[i];
        inv = inv * in[i].Z;
        Fq zi2 = zi.square();
        out[i] = {in[i].X * zi2, in[i].Y * zi2 * zi, false};
    }
}

// Fixed-base table for the generator's many [k]G (stands in for group::Wnaf,
// generator.rs:209-226; only the resulting group elements matter).
template <class Fq>
struct FixedBase {
    static constexpr int W = 8;
    std::vector<Affine<Fq>> table;   // table[w*255 + (d-1)] = d * 2^(8w) * G
    explicit FixedBase(const Jac<Fq>& g) {
        std::vector<Jac<Fq>> t(32 * 255);
        Jac<Fq> base = g;
        for (int w = 0; w < 32; w++) {
            Jac<Fq> cur = base;
            for (int d = 1; d <= 255; d++) {
                t[w * 255 + d - 1] = cur;
                cur = cur.add(base);
            }
            base = cur;               // 256 * base
        }
        table.resize(t.size());
        batch_to_affine(t, table.data());
    }
    Jac<Fq> mul(const Fr& k) const {
        auto bits = k.to_bits();
        Jac<Fq> acc = Jac<Fq>::identity();
        for (int w = 0; w < 32; w++) {
            unsigned d = (bits[w / 8] >> (8 * (w % 8))) & 0xff;
            if (d) acc = acc.add_mixed(table[w * 255 + d - 1]);
        }
        return acc;
    }
};

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef Jac<Fp> G1;
typedef Jac<Fp2> G2;

inline G1 g1_generator() {
    return G1::from_affine({Fp::from_raw(o1c::G1_GEN_X_M), Fp::from_raw(o1c::G1_GEN_Y_M), false});
}
inline G2 g2_generator() {
    return G2::from_affine({{Fp::from_raw(o1c::G2_GEN_X0_M), Fp::from_raw(o1c::G2_GEN_X1_M)},
                            {Fp::from_raw(o1c::G2_GEN_Y0_M), Fp::from_raw(o1c::G2_GEN_Y1_M)}, false});
}

inline bool g1_on_curve(const G1Affine& p) {
    return p.inf || p.y.square() == p.x.square() * p.x + Fp::from_raw(o1c::G1_B_M);
}
inline bool g2_on_curve(const G2Affine& p) {
    Fp b = Fp::from_raw(o1c::G1_B_M);
    return p.inf || p.y.square() == p.x.square() * p.x + Fp2{b, b};
}

// --- ZCash compressed encodings (GroupEncoding::to_bytes) --------------------
inline void fp_to_be(const Fp& v, uint8_t* out) {
    uint64_t c[6];
    v.to_canonical(c);
    for (int i = 0; i < 6; i++)
        for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(c[i] >> (8 * b));
}
inline bool fp_lexi_larger(const Fp& v) {       // v > (p-1)/2
    uint64_t c[6];
    v.to_canonical(c);
    for (int i = 5; i >= 0; i--) {
        if (c[i] > o1c::FP_HALF[i]) return true;
        if (c[i] < o1c::FP_HALF[i]) return false;
    }
    return false;
}
inline void g1_compress(const G1Affine& p, uint8_t* out) {
    if (p.inf) { std::memset(out, 0, 48); out[0] = 0xC0; return; }
    fp_to_be(p.x, out);
    out[0] |= 0x80;
    if (fp_lexi_larger(p.y)) out[0] |= 0x20;
}
inline void g2_compress(const G2Affine& p, uint8_t* out) {
    if (p.inf) { std::memset(out, 0, 96); out[0] = 0xC0; return; }
    fp_to_be(p.x.c1, out);
    fp_to_be(p.x.c0, out + 48);
    out[0] |= 0x80;
    bool larger = p.y.c1.is_zero() ? fp_lexi_larger(p.y.c0) : fp_lexi_larger(p.y.c1);
    if (larger) out[0] |= 0x20;
}

}  // namespace o1
