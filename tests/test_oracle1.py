"""Pin oracle-1 (C++ limb-level restatement, the parity oracle for the CUDA path).

1. DummyEngine instantiation == the reference's golden vectors
   (groth16/src/tests/mod.rs:91-373 and :375-440).
2. BLS12-381 instantiation == oracle-0 (Python integers) on small inputs, plus
   the properties the reference itself tests (multiexp.rs:334-378 naive==fast,
   domain.rs:436-498 FFT round trips / parallel==serial, mimc.rs prove).
"""
import random

import numpy as np

from oracle import o1
from oracle.oracle0 import bellman as B
from oracle.oracle0 import fields as F

R = F.FR_MODULUS
P = F.FP_MODULUS


def test_dummy_xordemo_golden():
    rc, o = o1.dummy_xordemo()
    assert rc == 0
    fr = B.DummyEngine.fr
    alpha, beta, gamma, delta, tau = 48577, 22580, 53332, 5481, 3673
    r, s = 27134, 17146
    t_at_tau = (pow(tau, 8, fr.q) - 1) % fr.q
    coeff = fr.inv(delta) * t_at_tau % fr.q
    assert o[0:7] == [pow(tau, i, fr.q) * coeff % fr.q for i in range(7)]          # :160-173
    u_i = [59158, 48317, 21767, 10402]                                             # :216-227
    v_i = [0, 0, 60619, 30791]
    w_i = [0, 23320, 41193, 41193]
    assert o[9:13] == u_i                                                          # :229-231
    assert o[13:15] == [60619, 30791] and o[15:17] == [60619, 30791]               # :233-239
    ext = [(beta * u_i[i] + alpha * v_i[i] + w_i[i]) % fr.q for i in range(4)]
    assert o[17:19] == [ext[i] * fr.inv(gamma) % fr.q for i in range(2)]           # ic :251-255
    assert o[7:9] == [ext[i] * fr.inv(delta) % fr.q for i in range(2, 4)]          # l  :256-261
    assert o[19:25] == [alpha, beta, beta, gamma, delta, delta]                    # :265-270
    assert o[28:35] == [5040, 11763, 10755, 63633, 128, 9747, 8739]                # :358
    a, b, c = o[25:28]
    assert a == (delta * r + alpha + u_i[0] + u_i[1] + u_i[2]) % fr.q              # :293-301
    assert b == (delta * s + beta + v_i[0] + v_i[1] + v_i[2]) % fr.q               # :310-318
    exp_c = (a * s + b * r - delta * r * s + o[7]) % fr.q                          # :335-369
    for i, co in enumerate(o[28:35]):
        exp_c = (exp_c + o[i] * co) % fr.q
    assert c == exp_c
    assert (a, b, c) == (3269, 471, 8383)


def test_dummy_zero_coeff_golden():
    assert o1.dummy_zero_coeff(True) == 1                                          # :432-435
    assert o1.dummy_zero_coeff(False) == 1                                         # :437-440


def test_field_arithmetic_vs_python():
    rng = random.Random(11)
    edge_r = [0, 1, 2, R - 1, R - 2, (1 << 255) % R, (1 << 256) % R]
    xs = edge_r + [rng.randrange(R) for _ in range(200)]
    ys = [rng.choice(edge_r) for _ in edge_r] + [rng.randrange(R) for _ in range(200)]
    a, b = o1.fr_from_ints(xs), o1.fr_from_ints(ys)
    assert o1.fr_to_ints(a) == xs
    assert o1.fr_to_ints(o1.fr_mul(a, b)) == [x * y % R for x, y in zip(xs, ys)]
    assert o1.fr_to_ints(o1.fr_add(a, b)) == [(x + y) % R for x, y in zip(xs, ys)]
    assert o1.fr_to_ints(o1.fr_sub(a, b)) == [(x - y) % R for x, y in zip(xs, ys)]
    # Montgomery form is x * 2^256 mod r (bls12_381::Scalar's internal form)
    assert o1.limbs_to_ints(a) == [x * (1 << 256) % R for x in xs]
    edge_p = [0, 1, P - 1, P - 2, (1 << 381) % P, (1 << 384) % P]
    xs = edge_p + [rng.randrange(P) for _ in range(200)]
    ys = [rng.choice(edge_p) for _ in edge_p] + [rng.randrange(P) for _ in range(200)]
    a, b = o1.fp_from_ints(xs), o1.fp_from_ints(ys)
    assert o1.fp_to_ints(a) == xs
    assert o1.fp_to_ints(o1.fp_mul(a, b)) == [x * y % P for x, y in zip(xs, ys)]
    assert o1.fp_to_ints(o1.fp_add(a, b)) == [(x + y) % P for x, y in zip(xs, ys)]
    assert o1.fp_to_ints(o1.fp_sub(a, b)) == [(x - y) % P for x, y in zip(xs, ys)]


def test_curve_ops_vs_python():
    rng = random.Random(12)
    assert o1.g1_to_affine_ints(o1.g1_generator())[0] == F.G1_GEN
    assert o1.g2_to_affine_ints(o1.g2_generator())[0] == F.G2_GEN
    ks = [0, 1, 2, R - 1] + [rng.randrange(R) for _ in range(6)]
    km = o1.fr_from_ints(ks)
    g1s = np.repeat(o1.g1_generator(), len(ks), axis=0)
    pts = o1.g1_mul(g1s, km)
    assert o1.g1_to_affine_ints(pts) == [F.G1.mul(F.G1_GEN, k) for k in ks]
    assert o1.g1_to_affine_ints(o1.g1_fixed_mul(km)) == [F.G1.mul(F.G1_GEN, k) for k in ks]
    assert o1.g1_on_curve(pts)
    # add incl. doubling, inverse and identity cases
    a = pts
    b = np.roll(pts, 1, axis=0)
    b[4] = a[4]                                   # doubling
    neg = o1.g1_from_affine_ints([F.G1.neg(o1.g1_to_affine_ints(a[5:6])[0])])
    b[5] = neg[0]                                 # P + (-P)
    exp = [F.G1.add(x, y) for x, y in zip(o1.g1_to_affine_ints(a), o1.g1_to_affine_ints(b))]
    assert o1.g1_to_affine_ints(o1.g1_add(a, b)) == exp
    g2s = np.repeat(o1.g2_generator(), 5, axis=0)
    k2 = [0, 1, 5, R - 1, rng.randrange(R)]
    p2 = o1.g2_mul(g2s, o1.fr_from_ints(k2))
    assert o1.g2_to_affine_ints(p2) == [F.G2.mul(F.G2_GEN, k) for k in k2]
    assert o1.g2_to_affine_ints(o1.g2_fixed_mul(o1.fr_from_ints(k2))) == [F.G2.mul(F.G2_GEN, k) for k in k2]
    assert o1.g2_on_curve(p2)
    exp2 = [F.G2.add(x, y) for x, y in zip(o1.g2_to_affine_ints(p2), o1.g2_to_affine_ints(np.roll(p2, 1, axis=0)))]
    assert o1.g2_to_affine_ints(o1.g2_add(p2, np.roll(p2, 1, axis=0))) == exp2
    # ZCash compressed encodings
    assert [bytes(r) for r in o1.g1_compress(pts)] == [F.g1_compress(p) for p in o1.g1_to_affine_ints(pts)]
    assert [bytes(r) for r in o1.g2_compress(p2)] == [F.g2_compress(p) for p in o1.g2_to_affine_ints(p2)]


def test_fft_vs_python_and_properties():
    rng = random.Random(13)
    for log_n in (0, 1, 2, 3, 6):
        n = 1 << log_n
        v = [rng.randrange(R) for _ in range(n)]
        vm = o1.fr_from_ints(v)
        for mode, fn in ((o1.FFT, "fft"), (o1.IFFT, "ifft"), (o1.COSET_FFT, "coset_fft"), (o1.ICOSET_FFT, "icoset_fft")):
            d = B.EvaluationDomain(F.FR, v)
            getattr(d, fn)()
            for threads in (1, 4):
                o1.set_threads(threads)
                assert o1.fr_to_ints(o1.fft(vm, mode)) == d.coeffs, (log_n, fn, threads)
        s = list(v)
        B.serial_fft(F.FR, s, B.EvaluationDomain(F.FR, v).omega, log_n)
        assert o1.fr_to_ints(o1.serial_fft(vm)) == s
    o1.set_threads(8)
    # domain.rs:436-457 fft_composition at a larger size, split FFT active (8 threads -> log_cpus 3)
    v = o1.fr_random(5, 1 << 12)
    assert np.array_equal(o1.fft(o1.fft(v, o1.IFFT), o1.FFT), v)
    assert np.array_equal(o1.fft(o1.fft(v, o1.ICOSET_FFT), o1.COSET_FFT), v)
    assert np.array_equal(o1.fft(o1.fft(v, o1.COSET_FFT), o1.ICOSET_FFT), v)
    # domain.rs:460-498 parallel_fft_consistency
    assert np.array_equal(o1.fft(v, o1.FFT), o1.serial_fft(v))


def test_h_poly_vs_python():
    rng = random.Random(14)
    n = 13
    a = [rng.randrange(R) for _ in range(n)]
    b = [rng.randrange(R) for _ in range(n)]
    c = [x * y % R for x, y in zip(a, b)]
    h = o1.h_poly(o1.fr_from_ints(a), o1.fr_from_ints(b), o1.fr_from_ints(c))
    assert o1.fr_to_ints(h) == B.h_coefficients(F.FR, a, b, c)
    assert h.shape[0] == 15


def test_multiexp_vs_python_and_naive():
    rng = random.Random(15)
    n = 40
    ks = [rng.randrange(1, R) for _ in range(n)]
    bases = o1.g1_fixed_mul(o1.fr_from_ints(ks))
    exps = [rng.randrange(R) for _ in range(n)]
    exps[3], exps[7], exps[9] = 0, 1, R - 1
    em = o1.fr_from_ints(exps)
    rc, fast = o1.multiexp(1, bases, 0, None, em)
    assert rc == 0
    py = B.multiexp(F.FR, F.G1, o1.g1_to_affine_ints(bases), 0, None, exps)
    assert o1.g1_to_affine_ints(fast)[0] == py
    assert np.array_equal(fast, o1.naive_multiexp(1, bases, em))
    # the group element equals [sum k_i e_i] G
    tot = sum(k * e for k, e in zip(ks, exps)) % R
    assert np.array_equal(fast, o1.g1_fixed_mul(o1.fr_from_ints([tot])))
    # density + offset (prover.rs:281-286 style)
    dens = np.array([rng.random() < 0.5 for _ in range(n)], dtype=np.uint8)
    k = int(dens.sum())
    rc, r2 = o1.multiexp(1, bases, 3, dens, em)
    exp_pts = o1.g1_to_affine_ints(bases)
    if 3 + k <= n:
        assert rc == 0
        assert o1.g1_to_affine_ints(r2)[0] == B.multiexp(F.FR, F.G1, exp_pts, 3, [bool(x) for x in dens], exps)
    # G2
    b2 = o1.g2_fixed_mul(o1.fr_from_ints(ks[:9]))
    rc, f2 = o1.multiexp(2, b2, 0, None, em[:9])
    assert rc == 0 and np.array_equal(f2, o1.naive_multiexp(2, b2, em[:9]))
    # n >= 32 uses c = ceil(ln n); test_with_bls12's property at 2^10
    n = 1 << 10
    ks = o1.fr_random(77, n)
    bases = o1.g1_fixed_mul(ks)
    em = o1.fr_random(78, n)
    rc, fast = o1.multiexp(1, bases, 0, None, em)
    kk, ee = o1.fr_to_ints(ks), o1.fr_to_ints(em)
    tot = sum(x * y for x, y in zip(kk, ee)) % R
    assert rc == 0 and np.array_equal(fast, o1.g1_fixed_mul(o1.fr_from_ints([tot])))


def test_multiexp_error_semantics():
    # identity skipped iff scalar zero; EOF; density mismatch
    bases = o1.g1_fixed_mul(o1.fr_from_ints([5, 6, 7]))
    bases[0] = 0                                              # identity base
    rc, out = o1.multiexp(1, bases, 0, None, o1.fr_from_ints([0, 2, 3]))
    assert rc == 0
    assert np.array_equal(out, o1.g1_fixed_mul(o1.fr_from_ints([6 * 2 + 7 * 3])))
    rc, _ = o1.multiexp(1, bases, 0, None, o1.fr_from_ints([2, 2, 3]))
    assert o1.ERR_NAMES[rc] == "UnexpectedIdentity"
    rc, _ = o1.multiexp(1, bases, 0, None, o1.fr_from_ints([1, 2, 3]))
    assert o1.ERR_NAMES[rc] == "UnexpectedIdentity"
    rc, _ = o1.multiexp(1, bases[1:], 0, None, o1.fr_from_ints([4, 2, 3]))
    assert o1.ERR_NAMES[rc] == "IoError(UnexpectedEof)"
    rc, out = o1.multiexp(1, bases, 0, None, o1.fr_from_ints([]))
    assert rc == 0 and not out.any()                          # n = 0 -> identity


def test_mimc_reference_config_prove():
    """BASELINE.json configs[0]: MiMC-322 (groth16/tests/mimc.rs) on the CPU path."""
    mc = o1.Mimc(322, seed=1)
    assert (mc.num_constraints, mc.num_inputs, mc.num_aux, mc.m) == (646, 2, 645, 1024)
    assert (mc.a_aux_total, mc.b_in_total, mc.b_aux_total) == (644, 1, 322)
    rng = random.Random(16)
    toxic = [rng.randrange(1, R) for _ in range(5)]
    mc.set_toxic(toxic)
    mc.generate()
    p = mc.export_params()
    assert p["h"].shape[0] == 1023 and p["l"].shape[0] == 645 and p["ic"].shape[0] == 2
    assert p["a"].shape[0] == 646 and p["b_g1"].shape[0] == 323 and p["b_g2"].shape[0] == 323
    # CRS elements are [k]G for the exported dlogs, identities filtered (generator.rs:491-505)
    ks = mc.crs_scalars()
    assert np.array_equal(p["h"], o1.g1_fixed_mul(ks["h"]))
    nz = [i for i, v in enumerate(o1.fr_to_ints(ks["a"])) if v]
    assert np.array_equal(p["a"], o1.g1_fixed_mul(ks["a"][nz]))
    r, s = rng.randrange(R), rng.randrange(R)
    proof = mc.prove(r, s)
    assert len(proof) == 192                                                       # lib.rs:559
    assert proof == mc.expected_proof(r, s)      # verifies in the exponent (tests/mod.rs:287-370 style)
    # and a different witness / randomness gives a different proof
    assert proof != mc.prove(r, (s + 1) % R)
