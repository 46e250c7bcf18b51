// Host-side execution of the DEVICE arithmetic path of bellman_b200/csrc/{mp,field}.cuh.
// Built by tests/test_emulated_device_field.py with  g++ -DBB_EMULATE_PTX : the PTX carry-chain
// primitives are modelled in C++ (mp.cuh, namespace ptx), everything above them -- the merged
// Montgomery product, wide_mul / wide_sqr / redc_wide, the lazy Fp2 product -- is the very
// template code the kernels instantiate.  Test infrastructure only; not part of the product build.
#ifndef BB_EMULATE_PTX
#error "compile with -DBB_EMULATE_PTX"
#endif
#include "field.cuh"
#include "curve.cuh"

using namespace bb;

extern "C" {
// op: 0 add, 1 sub, 2 mul, 3 sqr, 4 neg, 5 dbl     (raw limbs in, raw limbs out)
void emu_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fr x = fr_from_limbs(a + 8 * i), y = fr_from_limbs(b + 8 * i), r;
        switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                      case 3: r = x.sqr(); break; case 4: r = x.neg(); break; default: r = x.dbl(); }
        for (int k = 0; k < 8; k++) out[8 * i + k] = r.l[k];
    }
}
void emu_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fp x = fp_from_limbs(a + 12 * i), y = fp_from_limbs(b + 12 * i), r;
        switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                      case 3: r = x.sqr(); break; case 4: r = x.neg(); break; default: r = x.dbl(); }
        for (int k = 0; k < 12; k++) out[12 * i + k] = r.l[k];
    }
}
void emu_fp2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fp2 x{fp_from_limbs(a + 24 * i), fp_from_limbs(a + 24 * i + 12)};
        Fp2 y{fp_from_limbs(b + 24 * i), fp_from_limbs(b + 24 * i + 12)}, r;
        switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                      case 3: r = x.sqr(); break; case 4: r = x.neg(); break; default: r = x.dbl(); }
        for (int k = 0; k < 12; k++) { out[24 * i + k] = r.c0.l[k]; out[24 * i + 12 + k] = r.c1.l[k]; }
    }
}
// wide primitives on raw integers
void emu_wide_mul12(const uint32_t* a, const uint32_t* b, uint32_t* out24) { wide_mul<12>(out24, a, b); }
void emu_wide_sqr12(const uint32_t* a, uint32_t* out24) { wide_sqr<12>(out24, a); }
void emu_wide_mul8(const uint32_t* a, const uint32_t* b, uint32_t* out16) { wide_mul<8>(out16, a, b); }
void emu_wide_sqr8(const uint32_t* a, uint32_t* out16) { wide_sqr<8>(out16, a); }
void emu_redc_wide12(const uint32_t* t24, uint32_t* out12) { redc_wide<FpCfg>(out12, t24); }
void emu_redc_wide8(const uint32_t* t16, uint32_t* out8) { redc_wide<FrCfg>(out8, t16); }

// mixed addition / doubling chain on G1 and G2 through the emulated field: acc = sum_i P_i
// (affine Montgomery inputs, XYZZ accumulator -> affine out); identity = all zero
void emu_g1_sum(const uint32_t* pts, size_t n, uint32_t* out24) {
    XYZZ<Fp> acc = XYZZ<Fp>::identity();
    for (size_t i = 0; i < n; i++) {
        Affine<Fp> p{fp_from_limbs(pts + 24 * i), fp_from_limbs(pts + 24 * i + 12)};
        acc.add_mixed(p);
    }
    Affine<Fp> r = acc.to_affine();
    for (int k = 0; k < 12; k++) { out24[k] = r.x.l[k]; out24[12 + k] = r.y.l[k]; }
}
void emu_g2_sum(const uint32_t* pts, size_t n, uint32_t* out48) {
    XYZZ<Fp2> acc = XYZZ<Fp2>::identity();
    for (size_t i = 0; i < n; i++) {
        Affine<Fp2> p{{fp_from_limbs(pts + 48 * i), fp_from_limbs(pts + 48 * i + 12)},
                      {fp_from_limbs(pts + 48 * i + 24), fp_from_limbs(pts + 48 * i + 36)}};
        acc.add_mixed(p);
    }
    Affine<Fp2> r = acc.to_affine();
    for (int k = 0; k < 12; k++) {
        out48[k] = r.x.c0.l[k]; out48[12 + k] = r.x.c1.l[k];
        out48[24 + k] = r.y.c0.l[k]; out48[36 + k] = r.y.c1.l[k];
    }
}
}
