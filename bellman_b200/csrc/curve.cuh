// bellman_b200: BLS12-381 G1/G2 point arithmetic (a = 0 short Weierstrass).
//
// Stands in for the group operations multiexp() calls into the bls12_381 crate:
// bucket += affine base (/root/reference/src/multiexp.rs:39,253-262), projective
// adds of the summation by parts (:271-275), doublings of the window fold
// (:295-300), and the prover's final Affine*Fr / to_affine (groth16/src/prover.rs:
// 326-360).  Buckets use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ): a mixed addition
// costs 8M + 2S with no inversions and the identity is ZZ = 0.  Any coordinate
// system yields the same affine result, hence the same proof bytes.
#pragma once
#include "field.cuh"

namespace bb {

// Affine point as stored in HBM: x | y, Montgomery limbs; all-zero = identity
// (0,0) is not on either curve since b != 0.
template <class F>
struct Affine {
    F x, y;
    BB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
    BB_HD static Affine identity() { return {F::zero(), F::zero()}; }
    BB_HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;

    BB_HD static XYZZ identity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    BB_HD bool is_identity() const { return ZZ.is_zero(); }
    BB_HD static XYZZ from_affine(const Affine<F>& a) {
        if (a.is_identity()) return identity();
        return {a.x, a.y, FieldOps<F>::one(), FieldOps<F>::one()};
    }

    // 2*(x,y) for an affine input (dbl-2008-s-1 with ZZ = ZZZ = 1)
    BB_HD static XYZZ dbl_affine_impl(const Affine<F>& a) {
        F U = a.y.dbl();
        if (U.is_zero()) return identity();
        F V = U.sqr(), W = U * V, S = a.x * V;
        F xx = a.x.sqr();
        F M = xx.dbl() + xx;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * a.y;
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }
    BB_HD XYZZ dbl_impl() const {                        // dbl-2008-s-1
        if (is_identity()) return *this;
        F U = Y.dbl();
        F V = U.sqr(), W = U * V, S = X * V;
        F xx = X.sqr();
        F M = xx.dbl() + xx;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }
    // this += affine  (madd-2008-s), all special cases handled
    BB_HD void add_mixed(const Affine<F>& a) {
        if (a.is_identity()) return;
        if (is_identity()) { X = a.x; Y = a.y; ZZ = FieldOps<F>::one(); ZZZ = ZZ; return; }
        F U2 = a.x * ZZ, S2 = a.y * ZZZ;
        F P = U2 - X, R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(a);
            else *this = identity();
            return;
        }
        F PP = P.sqr(), PPP = P * PP, Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
    // this += o  (add-2008-s)
    BB_HD void add_impl(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F U1 = X * o.ZZ, U2 = o.X * ZZ;
        F S1 = Y * o.ZZZ, S2 = o.Y * ZZZ;
        F P = U2 - U1, R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F PP = P.sqr(), PPP = P * PP, Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }
    BB_HD XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

    // Out-of-line entry points.  On the device the heavy point operations are real function
    // calls (see BB_HD_NOINLINE in mp.cuh) taking and returning the points BY VALUE: handing
    // the address of a kernel-local point to a non-inlined member function made nvcc 12.9
    // drop a 16-byte store of that local afterwards (k_msm_reduce, seen in PTX and SASS).
    BB_HD XYZZ dbl() const;
    BB_HD void add(const XYZZ& o);
    BB_HD static XYZZ dbl_affine(const Affine<F>& a);

    // x = X/ZZ, y = Y/ZZZ
    BB_HD Affine<F> to_affine() const {
        if (is_identity()) return Affine<F>::identity();
        F zi3 = FieldOps<F>::inv(ZZZ);       // ZZ = z^2, ZZZ = z^3: 1/ZZZ = z^-3
        F zi2 = (zi3 * ZZ).sqr();            // (z^-3 z^2)^2 = z^-2
        return {X * zi2, Y * zi3};
    }

    // scalar given as canonical little-endian 8x32-bit integer
    BB_HD XYZZ mul_bits(const uint32_t* k) const {
        XYZZ acc = identity();
        for (int i = 7; i >= 0; i--)
            for (int b = 31; b >= 0; b--) {
                acc = acc.dbl();
                if ((k[i] >> b) & 1) acc.add(*this);
            }
        return acc;
    }
};

template <class F> BB_HD_NOINLINE XYZZ<F> xyzz_dbl_outlined(XYZZ<F> a) { return a.dbl_impl(); }
template <class F> BB_HD_NOINLINE XYZZ<F> xyzz_add_outlined(XYZZ<F> a, XYZZ<F> b) { a.add_impl(b); return a; }
template <class F> BB_HD_NOINLINE XYZZ<F> xyzz_dbl_affine_outlined(Affine<F> a) { return XYZZ<F>::dbl_affine_impl(a); }
template <class F> BB_HD XYZZ<F> XYZZ<F>::dbl() const { return xyzz_dbl_outlined<F>(*this); }
template <class F> BB_HD void XYZZ<F>::add(const XYZZ<F>& o) { *this = xyzz_add_outlined<F>(*this, o); }
template <class F> BB_HD XYZZ<F> XYZZ<F>::dbl_affine(const Affine<F>& a) { return xyzz_dbl_affine_outlined<F>(a); }

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1X;
typedef XYZZ<Fp2> G2X;

BB_HD G1Affine g1_generator() { return {Fp{{BBC_G1_GEN_X_M_LIST}}, Fp{{BBC_G1_GEN_Y_M_LIST}}}; }
BB_HD G2Affine g2_generator() {
    return {{Fp{{BBC_G2_GEN_X0_M_LIST}}, Fp{{BBC_G2_GEN_X1_M_LIST}}}, {Fp{{BBC_G2_GEN_Y0_M_LIST}}, Fp{{BBC_G2_GEN_Y1_M_LIST}}}};
}

}  // namespace bb
