"""Instruction-level emulation of the wide (2N-limb) product, the dedicated squaring and the
stand-alone Montgomery reduction of bellman_b200/csrc/mp.cuh.  Design check only."""
import random, sys
sys.path.insert(0, '.')
from oracle.oracle0.fields import FR_MODULUS, FP_MODULUS
M32 = 0xffffffff

class St: c = 0
def lo(a, b): return (a * b) & M32
def hi(a, b): return (a * b) >> 32
def addc(x, y, cin):
    s = x + y + cin
    St.c = s >> 32
    return s & M32

def wide_mul(a, b, N):
    """T[2N] = a*b.  A collects the products a[j]*b[i] with i+j even, B those with i+j odd
    (B is offset by one limb); inside a row the (lo,hi) pairs of either class are contiguous."""
    A = [0] * (2 * N + 1); B = [0] * (2 * N + 1)
    for i in range(N):
        # class "same parity as i" -> positions i+j even  (j = i%2, i%2+2, ...) goes to A
        for cls in (0, 1):
            acc = A if cls == 0 else B
            j0 = (i % 2) if cls == 0 else 1 - (i % 2)
            cin = 0
            pos = None
            for j in range(j0, N, 2):
                pos = i + j - cls          # index inside the accumulator (B is shifted by one limb)
                acc[pos] = addc(acc[pos], lo(a[j], b[i]), cin); cin = St.c
                acc[pos + 1] = addc(acc[pos + 1], hi(a[j], b[i]), cin); cin = St.c
            if pos is not None:
                acc[pos + 2] = addc(acc[pos + 2], 0, cin)
                assert St.c == 0
    # T = A + (B << 32)
    T = [0] * (2 * N)
    T[0] = A[0]; cin = 0
    for k in range(1, 2 * N):
        T[k] = addc(A[k], B[k - 1], cin); cin = St.c
    assert cin == 0 and A[2 * N] == 0 and B[2 * N - 1] + 0 == B[2 * N - 1]
    return T

def wide_sqr(a, N):
    """T[2N] = a^2: off-diagonal products once (same even/odd classes), doubled, plus diagonals."""
    A = [0] * (2 * N + 1); B = [0] * (2 * N + 1)
    for i in range(N):
        for cls in (0, 1):
            acc = A if cls == 0 else B
            # j > i with (i+j) % 2 == cls
            j0 = i + 2 if cls == 0 else i + 1
            cin = 0; pos = None
            for j in range(j0, N, 2):
                pos = i + j - cls
                acc[pos] = addc(acc[pos], lo(a[j], a[i]), cin); cin = St.c
                acc[pos + 1] = addc(acc[pos + 1], hi(a[j], a[i]), cin); cin = St.c
            if pos is not None:
                acc[pos + 2] = addc(acc[pos + 2], 0, cin)
                assert St.c == 0
    T = [0] * (2 * N)
    T[0] = A[0]; cin = 0
    for k in range(1, 2 * N):
        T[k] = addc(A[k], B[k - 1], cin); cin = St.c
    assert cin == 0
    # double: shift left by one bit
    for k in range(2 * N - 1, 0, -1):
        T[k] = ((T[k] << 1) | (T[k - 1] >> 31)) & M32
    T[0] = (T[0] << 1) & M32
    # add diagonals a[i]^2 at limbs (2i, 2i+1): one carry chain
    cin = 0
    for i in range(N):
        T[2 * i] = addc(T[2 * i], lo(a[i], a[i]), cin); cin = St.c
        T[2 * i + 1] = addc(T[2 * i + 1], hi(a[i], a[i]), cin); cin = St.c
    assert cin == 0
    return T

def redc_low(t, p, inv, N):
    """t (N limbs, any value < 2^(32N)) -> t * 2^(-32N) mod p, result < p + 1 before the final
    conditional subtraction: the merged CIOS routine of mp.cuh with b = 1 (the a*b rows vanish)."""
    X = [0] * (N + 1); Y = [0] * (N + 1)
    for k in range(0, N, 2):
        X[k] = t[k]; X[k + 1] = 0
        Y[k] = t[k + 1]; Y[k + 1] = 0
    def cmad_even(E, x, s):
        cin = 0
        for k in range(0, N, 2):
            E[k] = addc(E[k], lo(x[k], s), cin); cin = St.c
            E[k + 1] = addc(E[k + 1], hi(x[k], s), cin); cin = St.c
        E[N] = addc(E[N], 0, cin); assert St.c == 0
    def cmad_odd(O, x, s):
        cin = 0
        for k in range(0, N, 2):
            O[k] = addc(O[k], lo(x[k + 1], s), cin); cin = St.c
            O[k + 1] = addc(O[k + 1], hi(x[k + 1], s), cin); cin = St.c
        O[N] = addc(O[N], 0, cin); assert St.c == 0
    E, O = X, Y
    for i in range(N):
        if i > 0:
            assert E[0] == 0
            O[0] = addc(O[0], E[1], 0); cin = St.c
            for k in range(N):                       # E := (E >> 64) + carry
                src = E[k + 2] if k + 2 <= N else 0
                E[k] = addc(src, 0, cin); cin = St.c
            E[N] = cin
            E, O = O, E
        m = (E[0] * inv) & M32
        cmad_odd(O, p, m)
        cmad_even(E, p, m)
    assert E[0] == 0
    res = [0] * N; cin = 0
    for k in range(N):
        res[k] = addc(O[k], E[k + 1], cin); cin = St.c
    assert O[N] + cin == 0
    return res

def limbs(v, n): return [(v >> (32 * i)) & M32 for i in range(n)]
def val(l): return sum(x << (32 * i) for i, x in enumerate(l))

for q, N in ((FR_MODULUS, 8), (FP_MODULUS, 12)):
    inv = (-pow(q, -1, 1 << 32)) % (1 << 32)
    R = 1 << (32 * N); Rinv = pow(R, -1, q)
    rng = random.Random(2)
    edge = [0, 1, q - 1, q - 2, R - 1, (1 << 32) - 1, R >> 1, 2 * q - 1 if 2 * q - 1 < R else q]
    cases = [(x, y) for x in edge for y in edge] + [(rng.randrange(R), rng.randrange(R)) for _ in range(1500)]
    for x, y in cases:
        assert val(wide_mul(limbs(x, N), limbs(y, N), N)) == x * y
        assert val(wide_sqr(limbs(x, N), N)) == x * x
    for x, _ in cases:
        r = val(redc_low(limbs(x, N), limbs(q, N), inv, N))
        assert r <= q and (r - x * Rinv) % q == 0, hex(x)
    # full reduction of T < q*R : REDC(T) = T_hi + redc_low(T_lo) mod q
    for _ in range(1500):
        T = rng.randrange(q * R)
        r = val(redc_low(limbs(T % R, N), limbs(q, N), inv, N)) + (T >> (32 * N))
        assert r < 2 * q + 1 and (r - T * Rinv) % q == 0
    print("ok", N)
