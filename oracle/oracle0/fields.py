"""Oracle-0: BLS12-381 parameters and big-int field / curve arithmetic.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (bellman_b200/, csrc/)
may import this package; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may.

The arithmetic that bellman's hot path sits on lives in third-party crates
that are NOT in /root/reference (Cargo.lock:105-108 bls12_381 0.8.0,
:310-313 ff 0.13.0, :364-367 group 0.13.0).  This file restates the public
BLS12-381 parameter set with plain Python integers so the limb-level C++
oracle (oracle/oracle1) and the CUDA kernels can be cross-checked against
"obviously correct" arithmetic.  Every constant is self-checked on import
(curve equations, subgroup order, root-of-unity order).
"""

# --- BLS12-381 parameters -------------------------------------------------
# curve parameter z (negative), r = z^4 - z^2 + 1, p = (z-1)^2 r / 3 + z
BLS_X = -0xD201000000010000
FR_MODULUS = BLS_X**4 - BLS_X**2 + 1
FP_MODULUS = (BLS_X - 1) ** 2 * FR_MODULUS // 3 + BLS_X
assert FR_MODULUS == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
assert FP_MODULUS == int(
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
    "1eabfffeb153ffffb9feffffffffaaab", 16)

FR_S = 32                      # 2-adicity of r-1   (ff::PrimeField::S)
FR_GENERATOR = 7               # MULTIPLICATIVE_GENERATOR of bls12_381::Scalar
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (FR_MODULUS - 1) >> FR_S, FR_MODULUS)
assert pow(FR_ROOT_OF_UNITY, 1 << FR_S, FR_MODULUS) == 1
assert pow(FR_ROOT_OF_UNITY, 1 << (FR_S - 1), FR_MODULUS) != 1
FR_NUM_BITS = 255

G1_B = 4                       # y^2 = x^3 + 4
G2_B = (4, 4)                  # y^2 = x^3 + 4(u+1)

G1_GEN = (
    int("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
        "6c55e83ff97a1aeffb3af00adb22c6bb", 16),
    int("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3ed"
        "d03cc744a2888ae40caa232946c5e7e1", 16),
)
G2_GEN = (
    (int("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d177"
         "0bac0326a805bbefd48056c8c121bdb8", 16),
     int("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
         "334cf11213945d57e5ac7d055d042b7e", 16)),
    (int("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c"
         "923ac9cc3baca289e193548608b82801", 16),
     int("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab"
         "3f370d275cec1da1aaa9075ff05f79be", 16)),
)


class PrimeField:
    """Integers mod q with the handful of ops the reference path calls
    (square, invert, pow_vartime, mul/add/sub_assign)."""

    def __init__(self, q, s=None, generator=None, root_of_unity=None, num_bits=None):
        self.q = q
        self.S = s
        self.GENERATOR = generator
        self.ROOT_OF_UNITY = root_of_unity
        self.NUM_BITS = num_bits if num_bits is not None else q.bit_length()
        self.ZERO = 0
        self.ONE = 1

    def add(self, a, b): return (a + b) % self.q
    def sub(self, a, b): return (a - b) % self.q
    def mul(self, a, b): return (a * b) % self.q
    def neg(self, a): return (-a) % self.q
    def inv(self, a):
        assert a % self.q != 0
        return pow(a, -1, self.q)
    def pow(self, a, e): return pow(a, e, self.q)


FR = PrimeField(FR_MODULUS, FR_S, FR_GENERATOR, FR_ROOT_OF_UNITY, FR_NUM_BITS)
P = FP_MODULUS


# --- Fp / Fp2 element helpers (Fp2 = Fp[u]/(u^2+1), elements are (c0, c1)) --
class FpOps:
    zero = 0
    one = 1
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def inv(a): return pow(a, -1, P)
    @staticmethod
    def is_zero(a): return a % P == 0
    @staticmethod
    def from_int(k): return k % P


class Fp2Ops:
    zero = (0, 0)
    one = (1, 0)
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)
    @staticmethod
    def inv(a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
        return (a[0] * n % P, (-a[1]) * n % P)
    @staticmethod
    def is_zero(a): return a[0] % P == 0 and a[1] % P == 0
    @staticmethod
    def from_int(k): return (k % P, 0)


class Curve:
    """Short Weierstrass y^2 = x^3 + b over Fp or Fp2.  Points are affine
    tuples (x, y) or None for the identity.  Slow and simple on purpose."""

    def __init__(self, F, b, gen, name):
        self.F, self.b, self.gen, self.name = F, b, gen, name

    def identity(self): return None
    def is_identity(self, pt): return pt is None

    def is_on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.is_zero(F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)))

    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    def add(self, p1, p2):
        F = self.F
        if p1 is None: return p2
        if p2 is None: return p1
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if F.is_zero(F.add(y1, y2)):
                return None
            lam = F.mul(F.mul(F.from_int(3), F.mul(x1, x1)), F.inv(F.add(y1, y1)))
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def double(self, pt): return self.add(pt, pt)

    def mul(self, pt, k):
        k %= FR_MODULUS
        acc = None
        add = pt
        while k:
            if k & 1:
                acc = self.add(acc, add)
            add = self.add(add, add)
            k >>= 1
        return acc


G1 = Curve(FpOps, G1_B, G1_GEN, "G1")
G2 = Curve(Fp2Ops, G2_B, G2_GEN, "G2")

# self-check the recalled generators: on curve and of order r
assert G1.is_on_curve(G1_GEN) and G1.mul(G1_GEN, FR_MODULUS - 1) == G1.neg(G1_GEN)
assert G2.is_on_curve(G2_GEN) and G2.mul(G2_GEN, FR_MODULUS - 1) == G2.neg(G2_GEN)


# --- ZCash encodings (groth16/src/lib.rs:39-45 uses GroupEncoding::to_bytes) --
def _fp_lexi_larger(y):          # y > -y  <=>  y > (p-1)/2
    return y > (P - 1) // 2


def _fp2_lexi_larger(y):         # compare c1 first, then c0
    if y[1] != 0:
        return y[1] > (P - 1) // 2
    return y[0] > (P - 1) // 2


def g1_compress(pt):
    if pt is None:
        return bytes([0xC0] + [0] * 47)
    b = bytearray(pt[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _fp_lexi_larger(pt[1]):
        b[0] |= 0x20
    return bytes(b)


def g2_compress(pt):
    if pt is None:
        return bytes([0xC0] + [0] * 95)
    x, y = pt
    b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _fp2_lexi_larger(y):
        b[0] |= 0x20
    return bytes(b)


def g1_uncompressed(pt):
    if pt is None:
        return bytes([0x40] + [0] * 95)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def g2_uncompressed(pt):
    if pt is None:
        return bytes([0x40] + [0] * 191)
    x, y = pt
    return (x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big")
            + y[1].to_bytes(48, "big") + y[0].to_bytes(48, "big"))
