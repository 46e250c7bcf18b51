"""Oracle-0: BLS12-381 pairing and the Groth16 verifier, in Python integers.

TEST INFRASTRUCTURE ONLY (see fields.py).  Restates what `groth16::verify_proof`
(/root/reference/groth16/src/verifier.rs:23-58) needs from the `pairing` / `bls12_381` crates
(not in /root/reference): a bilinear, non-degenerate pairing e: G1 x G2 -> Gt.  Any such
pairing decides the verification equation identically, so this one is written for clarity, not
to reproduce bls12_381's Gt values: Fp12 is Fp[w]/(w^12 - 2 w^6 + 2) (w^6 = 1 + u), G2 points
are moved to E(Fp12) through the sextic twist, and the Miller loop is the textbook double-and-add
over |z| with affine line functions; final exponentiation by (p^12 - 1)/r.
Self-checks (tests/test_oracle0_pairing.py): the twist lands on y^2 = x^3 + 4, bilinearity in
both arguments, non-degeneracy.

Also: `Proof::read` (groth16/src/lib.rs:47-99): ZCash compressed point decoding with the
on-curve, subgroup and not-identity checks.
"""
from . import fields as F

P = F.FP_MODULUS
R_ORDER = F.FR_MODULUS
ATE_LOOP = -F.BLS_X            # |z| = 0xd201000000010000

# w^12 = 2 w^6 - 2
_DEG = 12


class Fp12:
    __slots__ = ("c",)

    def __init__(self, coeffs):
        self.c = [x % P for x in coeffs] + [0] * (_DEG - len(coeffs))

    @staticmethod
    def one():
        return Fp12([1])

    @staticmethod
    def zero():
        return Fp12([0])

    @staticmethod
    def from_fp(a):
        return Fp12([a])

    def __eq__(self, o):
        return self.c == o.c

    def is_zero(self):
        return not any(self.c)

    def __add__(self, o):
        return Fp12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return Fp12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return Fp12([-a for a in self.c])

    def scale(self, k):
        return Fp12([a * k for a in self.c])

    def __mul__(self, o):
        t = [0] * (2 * _DEG - 1)
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for k in range(2 * _DEG - 2, _DEG - 1, -1):      # w^k = 2 w^(k-6) - 2 w^(k-12)
            v = t[k]
            if v:
                t[k - 6] += 2 * v
                t[k - 12] -= 2 * v
        return Fp12(t[:_DEG])

    def pow(self, e):
        res, base = Fp12.one(), self
        while e:
            if e & 1:
                res = res * base
            base = base * base
            e >>= 1
        return res

    def inv(self):
        """extended Euclid on polynomials over Fp against the modulus w^12 - 2 w^6 + 2"""
        lm, hm = [1] + [0] * _DEG, [0] * (_DEG + 1)
        low, high = self.c + [0], [2, 0, 0, 0, 0, 0, (-2) % P, 0, 0, 0, 0, 0, 1]

        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        def poly_div(a, b):             # quotient of a / b, rounding towards the leading terms
            dega, degb = deg(a), deg(b)
            temp, o = list(a), [0] * len(a)
            binv = pow(b[degb], -1, P)
            for i in range(dega - degb, -1, -1):
                q = temp[degb + i] * binv % P
                o[i] = (o[i] + q) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - q * b[c]) % P
            return o[: deg(o) + 1]

        while deg(low):
            r = poly_div(high, low)
            r += [0] * (_DEG + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(_DEG + 1):
                for j in range(_DEG + 1 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                    new[i + j] = (new[i + j] - low[i] * r[j]) % P
            lm, low, hm, high = nm, new, lm, low
        linv = pow(low[0], -1, P)
        return Fp12([x * linv for x in lm[:_DEG]])

    def __truediv__(self, o):
        return self * o.inv()


W = Fp12([0, 1])
_W2_INV = (W * W).inv()
_W3_INV = (W * W * W).inv()


def _fp2_to_fp12(a):
    # u = w^6 - 1:  a0 + a1 u = (a0 - a1) + a1 w^6
    return Fp12([a[0] - a[1], 0, 0, 0, 0, 0, a[1]])


def twist(q):
    """E'(Fp2): y^2 = x^3 + 4(u+1)  ->  E(Fp12): y^2 = x^3 + 4,  (x, y) -> (x / w^2, y / w^3)"""
    if q is None:
        return None
    return (_fp2_to_fp12(q[0]) * _W2_INV, _fp2_to_fp12(q[1]) * _W3_INV)


def cast_g1(p):
    return None if p is None else (Fp12.from_fp(p[0]), Fp12.from_fp(p[1]))


def _on_curve12(pt):
    x, y = pt
    return y * y - x * x * x == Fp12([4])


def _double(pt):
    x, y = pt
    m = (x * x).scale(3) / y.scale(2)
    nx = m * m - x.scale(2)
    return (nx, m * (x - nx) - y)


def _add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        return _double(p1) if y1 == y2 else None
    m = (y2 - y1) / (x2 - x1)
    nx = m * m - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not (x1 == x2):
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) / y1.scale(2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q12, p12):
    if q12 is None or p12 is None:
        return Fp12.one()
    r, f = q12, Fp12.one()
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        f = f * f * _linefunc(r, r, p12)
        r = _double(r)
        if (ATE_LOOP >> i) & 1:
            f = f * _linefunc(r, q12, p12)
            r = _add(r, q12)
    return f


FINAL_EXP = (P ** 12 - 1) // R_ORDER


def final_exponentiation(f):
    return f.pow(FINAL_EXP)


def pairing(p_g1, q_g2):
    """e(P, Q), P affine in G1 (or None), Q affine in G2 (or None)"""
    return final_exponentiation(miller_loop(twist(q_g2), cast_g1(p_g1)))


def multi_pairing(pairs):
    """prod e(P_i, Q_i) with a single final exponentiation (MultiMillerLoop, verifier.rs:46-52)"""
    f = Fp12.one()
    for p_g1, q_g2 in pairs:
        f = f * miller_loop(twist(q_g2), cast_g1(p_g1))
    return final_exponentiation(f)


# ---------------------------------------------------------------------------
# groth16/src/verifier.rs
# ---------------------------------------------------------------------------
def verify_proof(vk, proof, public_inputs):
    """vk: dict alpha_g1, beta_g2, gamma_g2, delta_g2, ic (affine tuples); proof: (A, B, C).
    verifier.rs:23-58: e(A,B) e(acc,-gamma) e(C,-delta) == e(alpha,beta)."""
    if len(public_inputs) + 1 != len(vk["ic"]):
        raise ValueError("InvalidVerifyingKey")
    acc = vk["ic"][0]
    for x, b in zip(public_inputs, vk["ic"][1:]):
        acc = F.G1.add(acc, F.G1.mul(b, x))
    a, b, c = proof
    lhs = multi_pairing([(a, b), (acc, F.G2.neg(vk["gamma_g2"])), (c, F.G2.neg(vk["delta_g2"]))])
    return lhs == pairing(vk["alpha_g1"], vk["beta_g2"])


# ---------------------------------------------------------------------------
# Proof::read (groth16/src/lib.rs:47-99): compressed ZCash encodings
# ---------------------------------------------------------------------------
def _fp_sqrt(a):
    r = pow(a, (P + 1) // 4, P)          # p = 3 mod 4
    return r if r * r % P == a % P else None


def _fp2_sqrt(a):
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        r = _fp_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = _fp_sqrt((-a0) % P)
        return None if r is None else (0, r)
    s = _fp_sqrt((a0 * a0 + a1 * a1) % P)
    if s is None:
        return None
    inv2 = pow(2, -1, P)
    for cand in ((a0 + s) * inv2 % P, (a0 - s) * inv2 % P):
        x0 = _fp_sqrt(cand)
        if x0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            if F.Fp2Ops.mul((x0, x1), (x0, x1)) == (a0, a1):
                return (x0, x1)
    return None


def g1_decompress(b):
    assert len(b) == 48
    if not b[0] & 0x80:
        raise ValueError("not compressed")
    if b[0] & 0x40:
        if any(b[1:]) or b[0] & 0x3F:
            raise ValueError("invalid infinity encoding")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        raise ValueError("x not canonical")
    y = _fp_sqrt((x * x * x + 4) % P)
    if y is None:
        raise ValueError("not on curve")
    if (y > (P - 1) // 2) != bool(b[0] & 0x20):
        y = P - y
    pt = (x, y)
    if F.G1.mul(pt, R_ORDER - 1) != F.G1.neg(pt):
        raise ValueError("not in the prime-order subgroup")
    return pt


def g2_decompress(b):
    assert len(b) == 96
    if not b[0] & 0x80:
        raise ValueError("not compressed")
    if b[0] & 0x40:
        if any(b[1:]) or b[0] & 0x3F:
            raise ValueError("invalid infinity encoding")
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    if x0 >= P or x1 >= P:
        raise ValueError("x not canonical")
    x = (x0, x1)
    rhs = F.Fp2Ops.add(F.Fp2Ops.mul(F.Fp2Ops.mul(x, x), x), F.G2_B)
    y = _fp2_sqrt(rhs)
    if y is None:
        raise ValueError("not on curve")
    if F._fp2_lexi_larger(y) != bool(b[0] & 0x20):
        y = F.Fp2Ops.neg(y)
    pt = (x, y)
    if F.G2.mul(pt, R_ORDER - 1) != F.G2.neg(pt):
        raise ValueError("not in the prime-order subgroup")
    return pt


def proof_read(b):
    """Proof::read: 192 bytes -> (A, B, C); rejects the point at infinity (lib.rs:59-69)."""
    if len(b) != 192:
        raise ValueError("proof must be 192 bytes")
    a, bb, c = g1_decompress(b[:48]), g2_decompress(b[48:144]), g1_decompress(b[144:])
    if a is None or bb is None or c is None:
        raise ValueError("point at infinity")
    return a, bb, c
