"""Pin oracle-0 (Python big-int restatement) on the reference's golden vectors.

Every assertion below mirrors one in /root/reference/groth16/src/tests/mod.rs
(test_xordemo :91-373, zero_coeff_* :375-440) — the only known-answer vectors
the reference holds for generator + domain + multiexp + prover.
"""
import random

from oracle.oracle0 import bellman as B
from oracle.oracle0 import fields as F

E = B.DummyEngine
fr = E.fr
ALPHA, BETA, GAMMA, DELTA, TAU = 48577, 22580, 53332, 5481, 3673   # tests/mod.rs:95-99
R, S = 27134, 17146                                               # tests/mod.rs:274-275


def _params():
    return B.generate_parameters(E, B.xor_demo(None, None), 1, 1, ALPHA, BETA, GAMMA, DELTA, TAU)


def test_toy_field_constants():
    # dummy_engine.rs:297-320 and tests/mod.rs:126-134
    assert fr.pow(fr.ROOT_OF_UNITY, 1 << 10) == 1
    w8 = fr.pow(fr.ROOT_OF_UNITY, 1 << 7)
    assert fr.pow(w8, 8) == 1 and w8 == 20201
    assert fr.pow(5, 63) == 57751 and fr.inv(57751) == 12832 and fr.inv(2) == 32257


def test_xordemo_crs():
    p = _params()
    assert len(p.h) == 7                                          # :124
    t_at_tau = fr.sub(fr.pow(TAU, 8), 1)
    w8 = fr.pow(fr.ROOT_OF_UNITY, 1 << 7)
    tmp = 1
    for i in range(8):                                            # :146-154
        tmp = fr.mul(tmp, fr.sub(TAU, fr.pow(w8, i)))
    assert tmp == t_at_tau
    coeff = fr.mul(fr.inv(DELTA), t_at_tau)
    cur = 1
    for h in p.h:                                                 # :160-173
        assert h == fr.mul(cur, coeff)
        cur = fr.mul(cur, TAU)
    assert len(p.ic) == 2 and len(p.l) == 2 and len(p.a) == 4     # :176-182
    assert len(p.b_g1) == 2 and len(p.b_g2) == 2                  # :185-186
    u_i = [59158, 48317, 21767, 10402]                            # :216-227
    v_i = [0, 0, 60619, 30791]
    w_i = [0, 23320, 41193, 41193]
    assert p.a == u_i
    assert p.b_g1 == [v for v in v_i if v] and p.b_g2 == [v for v in v_i if v]
    for i in range(4):                                            # :241-262
        t = fr.add(fr.add(fr.mul(BETA, u_i[i]), fr.mul(ALPHA, v_i[i])), w_i[i])
        if i < 2:
            assert fr.mul(t, fr.inv(GAMMA)) == p.ic[i]
        else:
            assert fr.mul(t, fr.inv(DELTA)) == p.l[i - 2]
    assert (p.alpha_g1, p.beta_g1, p.beta_g2) == (ALPHA, BETA, BETA)     # :265-270
    assert (p.gamma_g2, p.delta_g1, p.delta_g2) == (GAMMA, DELTA, DELTA)


def test_xordemo_proof():
    p = _params()
    det = {}
    a, b, c = B.create_proof(E, B.xor_demo(True, False), p, R, S, det)
    u_i = [59158, 48317, 21767, 10402]
    v_i = [0, 0, 60619, 30791]
    exp_a = (DELTA * R + ALPHA + u_i[0] + u_i[1] + u_i[2]) % fr.q        # :293-301
    exp_b = (DELTA * S + BETA + v_i[0] + v_i[1] + v_i[2]) % fr.q         # :310-318
    assert a == exp_a and b == exp_b
    hco = [5040, 11763, 10755, 63633, 128, 9747, 8739]                   # :358
    assert det["h_coeffs"] == hco
    exp_c = (a * S + b * R - DELTA * R * S + p.l[0]) % fr.q              # :335-367
    for i, co in enumerate(hco):
        exp_c = (exp_c + p.h[i] * co) % fr.q
    assert c == exp_c
    assert (a, b, c) == (3269, 471, 8383)          # SURVEY.md App. A derived literals
    assert B.dummy_verify(p, (a, b, c), [1])                             # :372
    assert not B.dummy_verify(p, (a, b, (c + 1) % fr.q), [1])


def test_zero_coeff_regression():
    # tests/mod.rs:409-440
    for one_var in (True, False):
        circ = B.mult_with_zero_coeffs(5, 6, 30, one_var)
        pk = B.generate_parameters(E, circ, 1, 1, ALPHA, BETA, GAMMA, DELTA, TAU)
        pf = B.create_proof(E, circ, pk, R, S)
        assert B.dummy_verify(pk, pf, [])


def test_fft_properties_toy():
    # domain.rs:436-457 fft_composition and :460-498 parallel_fft_consistency
    rng = random.Random(1)
    for log_n in range(0, 8):
        n = 1 << log_n
        v = [rng.randrange(fr.q) for _ in range(n)]
        d = B.EvaluationDomain(fr, v)
        d.ifft(); d.fft()
        assert d.coeffs == v
        d.icoset_fft(); d.coset_fft()
        assert d.coeffs == v
        d.coset_fft(); d.icoset_fft()
        assert d.coeffs == v
        for log_cpus in range(0, log_n + 1):
            s = list(v)
            B.serial_fft(fr, s, d.omega, log_n)
            q = list(v)
            B.parallel_fft(fr, q, d.omega, log_n, log_cpus)
            assert s == q
        # natural order in / natural order out: out[k] = sum_j a[j] w^{jk}
        s = list(v)
        B.serial_fft(fr, s, d.omega, log_n)
        for k in range(min(n, 4)):
            assert s[k] == sum(v[j] * fr.pow(d.omega, j * k) for j in range(n)) % fr.q


def test_polynomial_arith_toy():
    # domain.rs:385-418: FFT multiplication == schoolbook
    rng = random.Random(2)
    for na in range(0, 12):
        for nb in range(0, 12):
            a = [rng.randrange(fr.q) for _ in range(na)]
            b = [rng.randrange(fr.q) for _ in range(nb)]
            naive = [0] * (na + nb)
            for i, x in enumerate(a):
                for j, y in enumerate(b):
                    naive[i + j] = (naive[i + j] + x * y) % fr.q
            A = B.EvaluationDomain(fr, a + [0] * (na + nb - na))
            Bd = B.EvaluationDomain(fr, b + [0] * (na + nb - nb))
            A.fft(); Bd.fft(); A.mul_assign(Bd); A.ifft()
            assert A.coeffs[:na + nb] == naive


def test_multiexp_matches_naive_bls12():
    # multiexp.rs:334-378 (test_with_bls12) at a size Python can afford
    rng = random.Random(3)
    n = 40
    bases = [F.G1.mul(F.G1_GEN, rng.randrange(1, F.FR_MODULUS)) for _ in range(n)]
    exps = [rng.randrange(F.FR_MODULUS) for _ in range(n)]
    exps[3], exps[7] = 0, 1
    fast = B.multiexp(F.FR, F.G1, bases, 0, None, exps)
    assert fast == B.naive_multiexp(F.G1, bases, exps)
    b2 = [F.G2.mul(F.G2_GEN, rng.randrange(1, F.FR_MODULUS)) for _ in range(8)]
    e2 = [rng.randrange(F.FR_MODULUS) for _ in range(8)]
    assert B.multiexp(F.FR, F.G2, b2, 0, None, e2) == B.naive_multiexp(F.G2, b2, e2)


def test_multiexp_source_semantics():
    # skip/identity/EOF behaviour (multiexp.rs:53-86, 242-265)
    g = E.g1
    # identity base is fine when its scalar is zero, an error otherwise
    assert B.multiexp(fr, g, [0, 5], 0, None, [0, 3]) == 15
    try:
        B.multiexp(fr, g, [0, 5], 0, None, [2, 3]); assert False
    except B.UnexpectedIdentity:
        pass
    # running out of bases
    try:
        B.multiexp(fr, g, [5], 0, None, [2, 3]); assert False
    except B.UnexpectedEof:
        pass
    # density: bases exist only for set bits; extra trailing bases ignored
    assert B.multiexp(fr, g, [9, 7, 11, 999], 1, [False, True, True], [4, 2, 3]) == (2 * 7 + 3 * 11) % fr.q
    assert B.window_size(31) == 3 and B.window_size(32) == 4
    assert B.window_size(1 << 20) == 14 and B.window_size(1 << 24) == 17
