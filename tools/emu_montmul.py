"""Instruction-level emulation of the even/odd-accumulator Montgomery product used by
bellman_b200/csrc/mp.cuh (mad.lo.cc / madc.hi.cc chains).  Design check only."""
import random, sys
sys.path.insert(0, '.')
from oracle.oracle0.fields import FR_MODULUS, FP_MODULUS
M32 = 0xffffffff

class CC:
    def __init__(self): self.c = 0
cc = CC()
def lo(a,b): return (a*b) & M32
def hi(a,b): return (a*b) >> 32
def add_cc(x, y, cin=0):
    s = x + y + cin
    cc.c = s >> 32
    return s & M32

def montmul(a, b, p, inv, N):
    # a, b, p: limb lists
    X = [0]*(N+1); Y = [0]*(N+1)
    def row_first(E, O, x, s):
        for k in range(0, N, 2):
            E[k] = lo(x[k], s); E[k+1] = hi(x[k], s)
            O[k] = lo(x[k+1], s); O[k+1] = hi(x[k+1], s)
        E[N] = 0; O[N] = 0
    def cmad_even(E, x, s):   # E += even products of x*s  (fresh chain)
        cin = 0
        for k in range(0, N, 2):
            E[k] = add_cc(E[k], lo(x[k], s), cin); cin = cc.c
            E[k+1] = add_cc(E[k+1], hi(x[k], s), cin); cin = cc.c
        E[N] = add_cc(E[N], 0, cin); assert cc.c == 0
    def cmad_odd(O, x, s, cin=0):    # O += odd products of x*s (O is offset one limb)
        for k in range(0, N, 2):
            O[k] = add_cc(O[k], lo(x[k+1], s), cin); cin = cc.c
            O[k+1] = add_cc(O[k+1], hi(x[k+1], s), cin); cin = cc.c
        O[N] = add_cc(O[N], 0, cin); assert cc.c == 0
    def madc_rshift(Xs, x, s, cin):  # in place: newY[k] = Xs[k+2] + odd products + carry
        for k in range(0, N, 2):
            src0 = Xs[k+2] if k+2 <= N else 0
            src1 = Xs[k+3] if k+3 <= N else 0
            Xs[k] = add_cc(src0, lo(x[k+1], s), cin); cin = cc.c
            Xs[k+1] = add_cc(src1, hi(x[k+1], s), cin); cin = cc.c
        Xs[N] = cin
    E, O = X, Y
    for i in range(N):
        if i == 0:
            row_first(E, O, a, b[0])
        else:
            # previous: E* (aligned limb0, E*[0]==0), O*.  new even := O, new odd := E>>64
            assert E[0] == 0
            O[0] = add_cc(O[0], E[1]); c1 = cc.c
            madc_rshift(E, a, b[i], c1)
            E, O = O, E
            cmad_even(E, a, b[i])
        m = (E[0] * inv) & M32
        cmad_odd(O, p, m)
        cmad_even(E, p, m)
    assert E[0] == 0
    # result = O + (E >> 32)
    res = [0]*N; cin = 0
    for k in range(N):
        res[k] = add_cc(O[k], E[k+1], cin); cin = cc.c
    top = O[N] + cin
    assert top == 0, top
    return res

def limbs(v, N): return [(v >> (32*i)) & M32 for i in range(N)]
def val(l): return sum(x << (32*i) for i, x in enumerate(l))

for q, N in ((FR_MODULUS, 8), (FP_MODULUS, 12)):
    inv = (-pow(q, -1, 1 << 32)) % (1 << 32)
    Rinv = pow(1 << (32*N), -1, q)
    rng = random.Random(1)
    edge = [0, 1, q-1, q-2, (1 << (32*N-1)) % q, 2**32 - 1, (2**(32*N) - 1) % q]
    cases = [(x, y) for x in edge for y in edge] + [(rng.randrange(q), rng.randrange(q)) for _ in range(3000)]
    for x, y in cases:
        r = val(montmul(limbs(x, N), limbs(y, N), limbs(q, N), inv, N))
        assert r < 2*q
        if r >= q: r -= q
        assert r == x * y * Rinv % q, (hex(x), hex(y))
    print("ok", N)
