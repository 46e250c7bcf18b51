"""Oracle-0: restatement of bellman's Groth16 hot path with Python integers.

TEST INFRASTRUCTURE ONLY (see fields.py header).  Generic over an "engine"
so that the toy DummyEngine instantiation (groth16/src/tests/dummy_engine.rs)
reproduces the reference's only known-answer test
(groth16/src/tests/mod.rs:91-373), and the BLS12-381 instantiation pins the
limb-level oracle (oracle/oracle1) and the CUDA kernels on small sizes.

Each function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import math

from . import fields as F


# ---------------------------------------------------------------------------
# Engines
# ---------------------------------------------------------------------------
class AdditiveGroup:
    """DummyEngine groups: G1 = G2 = Fr under addition, scalar-mul is field
    multiplication (dummy_engine.rs:336-378; Group impl for Fr)."""

    def __init__(self, fr):
        self.fr = fr
    def identity(self): return 0
    def is_identity(self, a): return a % self.fr.q == 0
    def add(self, a, b): return (a + b) % self.fr.q
    def double(self, a): return (2 * a) % self.fr.q
    def mul(self, a, k): return (a * k) % self.fr.q
    def neg(self, a): return (-a) % self.fr.q


class DummyEngine:
    # dummy_engine.rs:15 (modulus 64513), :297-320 (S, generator, root)
    fr = F.PrimeField(64513, s=10, generator=5, root_of_unity=57751, num_bits=16)
    g1 = AdditiveGroup(fr)
    g2 = AdditiveGroup(fr)


class Bls12:
    fr = F.FR
    g1 = F.G1
    g2 = F.G2


# ---------------------------------------------------------------------------
# src/domain.rs
# ---------------------------------------------------------------------------
class PolynomialDegreeTooLarge(Exception):
    pass


class UnexpectedIdentity(Exception):
    pass


class UnexpectedEof(Exception):
    pass


def bitreverse(n, l):                       # domain.rs:273-280
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def serial_fft(fr, a, omega, log_n):        # domain.rs:272-314
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[rk], a[k] = a[k], a[rk]
    m = 1
    for _ in range(log_n):
        w_m = fr.pow(omega, n // (2 * m))
        k = 0
        while k < n:
            w = 1
            for j in range(m):
                t = fr.mul(a[k + j + m], w)
                a[k + j + m] = fr.sub(a[k + j], t)
                a[k + j] = fr.add(a[k + j], t)
                w = fr.mul(w, w_m)
            k += 2 * m
        m *= 2


def parallel_fft(fr, a, omega, log_n, log_cpus):   # domain.rs:316-372
    assert log_n >= log_cpus
    num_cpus = 1 << log_cpus
    log_new_n = log_n - log_cpus
    tmp = [[0] * (1 << log_new_n) for _ in range(num_cpus)]
    new_omega = fr.pow(omega, num_cpus)
    for j in range(num_cpus):
        omega_j = fr.pow(omega, j)
        omega_step = fr.pow(omega, j << log_new_n)
        elt = 1
        for i in range(1 << log_new_n):
            for s in range(num_cpus):
                idx = (i + (s << log_new_n)) % (1 << log_n)
                tmp[j][i] = fr.add(tmp[j][i], fr.mul(a[idx], elt))
                elt = fr.mul(elt, omega_step)
            elt = fr.mul(elt, omega_j)
        serial_fft(fr, tmp[j], new_omega, log_new_n)
    mask = (1 << log_cpus) - 1
    for idx in range(len(a)):
        a[idx] = tmp[idx & mask][idx >> log_cpus]


class EvaluationDomain:
    def __init__(self, fr, coeffs):          # from_coeffs, domain.rs:47-79
        self.fr = fr
        m, exp = 1, 0
        while m < len(coeffs):
            m *= 2
            exp += 1
            if exp >= fr.S:
                raise PolynomialDegreeTooLarge()
        omega = fr.ROOT_OF_UNITY
        for _ in range(exp, fr.S):
            omega = fr.mul(omega, omega)
        self.coeffs = list(coeffs) + [0] * (m - len(coeffs))
        self.exp = exp
        self.omega = omega
        self.omegainv = fr.inv(omega)
        self.geninv = fr.inv(fr.GENERATOR)
        self.minv = fr.inv(m % fr.q)

    def fft(self, log_cpus=0):               # domain.rs:81-83, 261-269
        self._best_fft(self.omega, log_cpus)

    def _best_fft(self, omega, log_cpus):
        if self.exp <= log_cpus:
            serial_fft(self.fr, self.coeffs, omega, self.exp)
        else:
            parallel_fft(self.fr, self.coeffs, omega, self.exp, log_cpus)

    def ifft(self, log_cpus=0):              # domain.rs:85-99
        self._best_fft(self.omegainv, log_cpus)
        self.coeffs = [self.fr.mul(v, self.minv) for v in self.coeffs]

    def distribute_powers(self, g):          # domain.rs:101-113
        u = 1
        for i in range(len(self.coeffs)):
            self.coeffs[i] = self.fr.mul(self.coeffs[i], u)
            u = self.fr.mul(u, g)

    def coset_fft(self, log_cpus=0):         # domain.rs:115-118
        self.distribute_powers(self.fr.GENERATOR)
        self.fft(log_cpus)

    def icoset_fft(self, log_cpus=0):        # domain.rs:120-125
        self.ifft(log_cpus)
        self.distribute_powers(self.geninv)

    def z(self, tau):                        # domain.rs:129-134
        return self.fr.sub(self.fr.pow(tau, len(self.coeffs)), 1)

    def divide_by_z_on_coset(self):          # domain.rs:139-151
        i = self.fr.inv(self.z(self.fr.GENERATOR))
        self.coeffs = [self.fr.mul(v, i) for v in self.coeffs]

    def mul_assign(self, other):             # domain.rs:154-170
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = [self.fr.mul(a, b) for a, b in zip(self.coeffs, other.coeffs)]

    def sub_assign(self, other):             # domain.rs:173-189
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = [self.fr.sub(a, b) for a, b in zip(self.coeffs, other.coeffs)]


# ---------------------------------------------------------------------------
# src/multiexp.rs
# ---------------------------------------------------------------------------
def window_size(n):                          # multiexp.rs:318-322
    return 3 if n < 32 else int(math.ceil(math.log(n)))


class Source:                                # multiexp.rs:45-86
    def __init__(self, group, bases, offset):
        self.g, self.bases, self.pos = group, bases, offset
    def next(self):
        if len(self.bases) <= self.pos:
            raise UnexpectedEof()
        if self.g.is_identity(self.bases[self.pos]):
            raise UnexpectedIdentity()
        b = self.bases[self.pos]
        self.pos += 1
        return b
    def skip(self, amt):
        if len(self.bases) <= self.pos:
            raise UnexpectedEof()
        self.pos += amt


def multiexp(engine_fr, group, bases, offset, density, exponents, c=None):
    """multiexp + multiexp_inner (multiexp.rs:210-332).  `density` is None
    (FullDensity) or a list of bools with len == len(exponents)."""
    n = len(exponents)
    if c is None:
        c = window_size(n)
    if density is not None:
        assert len(density) == n             # multiexp.rs:324-329
    num_bits = engine_fr.NUM_BITS
    mask = (1 << c) - 1
    parts = []
    for chunk, lo in enumerate(range(0, num_bits, c)):      # :288-293
        acc = group.identity()
        src = Source(group, bases, offset)
        buckets = [group.identity()] * ((1 << c) - 1)
        handle_trivial = chunk == 0
        for i, e in enumerate(exponents):                   # :242-265
            if density is not None and not density[i]:
                continue
            e %= engine_fr.q
            if e == 0:
                src.skip(1)
            elif e == 1:
                if handle_trivial:
                    acc = group.add(acc, src.next())
                else:
                    src.skip(1)
            else:
                d = (e >> lo) & mask
                if d != 0:
                    buckets[d - 1] = group.add(buckets[d - 1], src.next())
                else:
                    src.skip(1)
        running = group.identity()                          # :271-275
        for b in reversed(buckets):
            running = group.add(running, b)
            acc = group.add(acc, running)
        parts.append(acc)
    acc = group.identity()                                  # :295-300
    for part in reversed(parts):
        for _ in range(c):
            acc = group.double(acc)
        acc = group.add(acc, part)
    return acc


def naive_multiexp(group, bases, exponents):  # multiexp.rs:336-349 (test)
    acc = group.identity()
    for b, e in zip(bases, exponents):
        acc = group.add(acc, group.mul(b, e))
    return acc


# ---------------------------------------------------------------------------
# src/lib.rs circuit surface (only what the prover/generator need)
# ---------------------------------------------------------------------------
class Var:
    __slots__ = ("is_input", "idx")
    def __init__(self, is_input, idx):
        self.is_input, self.idx = is_input, idx


ONE = Var(True, 0)                            # ConstraintSystem::one(), lib.rs


class LC:
    """LinearCombination: ordered (Variable, coeff) terms (lib.rs:190-299).
    Terms are NOT merged, matching `lc + (coeff, var)` which pushes."""
    def __init__(self, terms=None):
        self.terms = list(terms or [])
    def add(self, var, coeff=1):
        return LC(self.terms + [(var, coeff)])
    def sub(self, var, coeff=1):
        return LC(self.terms + [(var, -coeff)])


class ProvingAssignment:                      # prover.rs:57-162
    def __init__(self, fr):
        self.fr = fr
        self.a_aux_density, self.b_input_density, self.b_aux_density = [], [], []
        self.a, self.b, self.c = [], [], []
        self.input_assignment, self.aux_assignment = [], []

    def alloc(self, f):                       # prover.rs:76-89
        self.aux_assignment.append(f() % self.fr.q)
        self.a_aux_density.append(False)
        self.b_aux_density.append(False)
        return Var(False, len(self.aux_assignment) - 1)

    def alloc_input(self, f):                 # prover.rs:91-103
        self.input_assignment.append(f() % self.fr.q)
        self.b_input_density.append(False)
        return Var(True, len(self.input_assignment) - 1)

    def _eval(self, lc, input_density, aux_density):      # prover.rs:19-55
        acc = 0
        for var, coeff in lc.terms:
            coeff %= self.fr.q
            if coeff != 0:
                if var.is_input:
                    tmp = self.input_assignment[var.idx]
                    if input_density is not None:
                        input_density[var.idx] = True
                else:
                    tmp = self.aux_assignment[var.idx]
                    if aux_density is not None:
                        aux_density[var.idx] = True
                acc = (acc + tmp * coeff) % self.fr.q
        return acc

    def enforce(self, a, b, c):               # prover.rs:105-145
        self.a.append(self._eval(a, None, self.a_aux_density))
        self.b.append(self._eval(b, self.b_input_density, self.b_aux_density))
        self.c.append(self._eval(c, None, None))


class KeypairAssembly:                        # generator.rs:43-155
    def __init__(self, fr):
        self.fr = fr
        self.num_inputs = self.num_aux = self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self, f):
        self.num_aux += 1
        self.at_aux.append([]); self.bt_aux.append([]); self.ct_aux.append([])
        return Var(False, self.num_aux - 1)

    def alloc_input(self, f):
        self.num_inputs += 1
        self.at_inputs.append([]); self.bt_inputs.append([]); self.ct_inputs.append([])
        return Var(True, self.num_inputs - 1)

    def enforce(self, a, b, c):
        def ev(lc, inputs, aux):              # generator.rs:104-116
            for var, coeff in lc.terms:
                (inputs if var.is_input else aux)[var.idx].append(
                    (coeff % self.fr.q, self.num_constraints))
        ev(a, self.at_inputs, self.at_aux)
        ev(b, self.bt_inputs, self.bt_aux)
        ev(c, self.ct_inputs, self.ct_aux)
        self.num_constraints += 1


# ---------------------------------------------------------------------------
# groth16/src/generator.rs
# ---------------------------------------------------------------------------
class Parameters:
    pass


def generate_parameters(E, circuit, g1, g2, alpha, beta, gamma, delta, tau):
    """generator.rs:159-507.  wNAF scalar-mul is replaced by plain scalar-mul
    (same group element)."""
    fr = E.fr
    asm = KeypairAssembly(fr)
    asm.alloc_input(lambda: 1)                               # :188
    circuit(asm)                                             # :191
    for i in range(asm.num_inputs):                          # :195-202
        asm.enforce(LC().add(Var(True, i)), LC(), LC())
    dom = EvaluationDomain(fr, [0] * asm.num_constraints)    # :205-206
    m = len(dom.coeffs)
    gamma_inverse = fr.inv(gamma)                            # :228-243
    delta_inverse = fr.inv(delta)
    cur = 1
    for i in range(m):                                       # :249-264
        dom.coeffs[i] = cur
        cur = fr.mul(cur, tau)
    coeff = fr.mul(dom.z(tau), delta_inverse)                # :267-268
    h = [E.g1.mul(g1, fr.mul(dom.coeffs[i], coeff)) for i in range(m - 1)]   # :271-296
    dom.ifft()                                               # :300
    lag = dom.coeffs

    def eval_at_tau(p):                                      # :376-389
        acc = 0
        for coeff_, index in p:
            acc = fr.add(acc, fr.mul(lag[index], coeff_))
        return acc

    def ev(at, bt, ct, inv):                                 # :310-426
        a, b1, b2, ext = [], [], [], []
        for at_i, bt_i, ct_i in zip(at, bt, ct):
            a_t, b_t, c_t = eval_at_tau(at_i), eval_at_tau(bt_i), eval_at_tau(ct_i)
            a.append(E.g1.mul(g1, a_t) if a_t != 0 else E.g1.identity())
            b1.append(E.g1.mul(g1, b_t) if b_t != 0 else E.g1.identity())
            b2.append(E.g2.mul(g2, b_t) if b_t != 0 else E.g2.identity())
            e = fr.mul(fr.add(fr.add(fr.mul(a_t, beta), fr.mul(b_t, alpha)), c_t), inv)
            ext.append(E.g1.mul(g1, e))
        return a, b1, b2, ext

    a_in, b1_in, b2_in, ic = ev(asm.at_inputs, asm.bt_inputs, asm.ct_inputs, gamma_inverse)
    a_aux, b1_aux, b2_aux, l = ev(asm.at_aux, asm.bt_aux, asm.ct_aux, delta_inverse)
    for e in l:                                              # :466-470
        if E.g1.is_identity(e):
            raise RuntimeError("UnconstrainedVariable")
    prm = Parameters()
    prm.alpha_g1 = E.g1.mul(g1, alpha); prm.beta_g1 = E.g1.mul(g1, beta)
    prm.beta_g2 = E.g2.mul(g2, beta); prm.gamma_g2 = E.g2.mul(g2, gamma)
    prm.delta_g1 = E.g1.mul(g1, delta); prm.delta_g2 = E.g2.mul(g2, delta)
    prm.ic, prm.h, prm.l = ic, h, l
    prm.a = [e for e in a_in + a_aux if not E.g1.is_identity(e)]       # :491-505
    prm.b_g1 = [e for e in b1_in + b1_aux if not E.g1.is_identity(e)]
    prm.b_g2 = [e for e in b2_in + b2_aux if not E.g2.is_identity(e)]
    prm.num_inputs, prm.num_aux, prm.m = asm.num_inputs, asm.num_aux, m
    return prm


# ---------------------------------------------------------------------------
# groth16/src/prover.rs
# ---------------------------------------------------------------------------
def synthesize_witness(E, circuit):           # prover.rs:193-215
    pa = ProvingAssignment(E.fr)
    pa.alloc_input(lambda: 1)
    circuit(pa)
    for i in range(len(pa.input_assignment)):
        pa.enforce(LC().add(Var(True, i)), LC(), LC())
    return pa


def h_coefficients(fr, a, b, c):              # prover.rs:221-242
    A, B, C = EvaluationDomain(fr, a), EvaluationDomain(fr, b), EvaluationDomain(fr, c)
    A.ifft(); A.coset_fft()
    B.ifft(); B.coset_fft()
    C.ifft(); C.coset_fft()
    A.mul_assign(B)
    A.sub_assign(C)
    A.divide_by_z_on_coset()
    A.icoset_fft()
    return A.coeffs[:-1]


def create_proof(E, circuit, prm, r, s, details=None):
    """prover.rs:182-361.  Returns (A, B, C) as group elements."""
    fr = E.fr
    pa = synthesize_witness(E, circuit)
    hco = h_coefficients(fr, pa.a, pa.b, pa.c)
    h = multiexp(fr, E.g1, prm.h, 0, None, hco)                              # :244
    inp, aux = pa.input_assignment, pa.aux_assignment
    l = multiexp(fr, E.g1, prm.l, 0, None, aux)                              # :263-268
    a_inputs = multiexp(fr, E.g1, prm.a, 0, None, inp)                       # :275-280
    a_aux = multiexp(fr, E.g1, prm.a, len(inp), pa.a_aux_density, aux)       # :281-286
    b_in_total = sum(pa.b_input_density)                                     # :288-291
    b_g1_inputs = multiexp(fr, E.g1, prm.b_g1, 0, pa.b_input_density, inp)   # :296-301
    b_g1_aux = multiexp(fr, E.g1, prm.b_g1, b_in_total, pa.b_aux_density, aux)
    b_g2_inputs = multiexp(fr, E.g2, prm.b_g2, 0, pa.b_input_density, inp)   # :312-318
    b_g2_aux = multiexp(fr, E.g2, prm.b_g2, b_in_total, pa.b_aux_density, aux)
    if E.g1.is_identity(prm.delta_g1) or E.g2.is_identity(prm.delta_g2):     # :320-324
        raise UnexpectedIdentity()
    g_a = E.g1.add(E.g1.mul(prm.delta_g1, r), prm.alpha_g1)                  # :326-327
    g_b = E.g2.add(E.g2.mul(prm.delta_g2, s), prm.beta_g2)                   # :328-329
    rs = fr.mul(r, s)
    g_c = E.g1.mul(prm.delta_g1, rs)                                         # :335-337
    g_c = E.g1.add(g_c, E.g1.mul(prm.alpha_g1, s))
    g_c = E.g1.add(g_c, E.g1.mul(prm.beta_g1, r))
    a_answer = E.g1.add(a_inputs, a_aux)                                     # :339-343
    g_a = E.g1.add(g_a, a_answer)
    g_c = E.g1.add(g_c, E.g1.mul(a_answer, s))
    b1_answer = E.g1.add(b_g1_inputs, b_g1_aux)                              # :345-354
    b2_answer = E.g2.add(b_g2_inputs, b_g2_aux)
    g_b = E.g2.add(g_b, b2_answer)
    g_c = E.g1.add(g_c, E.g1.mul(b1_answer, r))
    g_c = E.g1.add(g_c, h)
    g_c = E.g1.add(g_c, l)
    if details is not None:
        details.update(h_coeffs=hco, witness=pa, h=h, l=l, a_answer=a_answer,
                       b1_answer=b1_answer, b2_answer=b2_answer)
    return g_a, g_b, g_c


def proof_bytes(proof):                       # groth16/src/lib.rs:39-45 (BLS12-381 only)
    a, b, c = proof
    return F.g1_compress(a) + F.g2_compress(b) + F.g1_compress(c)


# ---------------------------------------------------------------------------
# Circuits used by the reference's tests
# ---------------------------------------------------------------------------
def xor_demo(a, b):                           # groth16/src/tests/mod.rs:13-89
    def synth(cs):
        a_var = cs.alloc(lambda: 1 if a else 0)
        cs.enforce(LC().add(ONE).sub(a_var), LC().add(a_var), LC())
        b_var = cs.alloc(lambda: 1 if b else 0)
        cs.enforce(LC().add(ONE).sub(b_var), LC().add(b_var), LC())
        c_var = cs.alloc_input(lambda: 1 if (bool(a) ^ bool(b)) else 0)
        cs.enforce(LC().add(a_var).add(a_var), LC().add(b_var),
                   LC().add(a_var).add(b_var).sub(c_var))
    return synth


def mult_with_zero_coeffs(a, b, c, one_var):  # groth16/src/tests/mod.rs:375-407
    def synth(cs):
        av = cs.alloc(lambda: a)
        bv = cs.alloc(lambda: b)
        cv = cs.alloc(lambda: c)
        if one_var:
            cs.enforce(LC().add(av), LC().add(ONE, 0).add(bv), LC().add(cv))
        else:
            cs.enforce(LC().add(av), LC().add(av, 0).add(bv), LC().add(cv))
    return synth


def mimc_hash(fr, xl, xr, constants):         # groth16/tests/common/mod.rs:20-35
    for c in constants:
        t = fr.add(xl, c)
        t2 = fr.mul(fr.mul(t, t), t)
        xl, xr = fr.add(t2, xr), xl
    return xl


def mimc_circuit(fr, xl, xr, constants):      # groth16/tests/common/mod.rs:48-129
    rounds = len(constants)
    def synth(cs):
        xl_v, xr_v = xl, xr
        xl_var = cs.alloc(lambda: xl_v)
        xr_var = cs.alloc(lambda: xr_v)
        for i in range(rounds):
            ci = constants[i]
            tmp_v = fr.mul(fr.add(xl_v, ci), fr.add(xl_v, ci))
            tmp = cs.alloc(lambda: tmp_v)
            cs.enforce(LC().add(xl_var).add(ONE, ci), LC().add(xl_var).add(ONE, ci),
                       LC().add(tmp))
            new_v = fr.add(fr.mul(fr.add(xl_v, ci), tmp_v), xr_v)
            if i == rounds - 1:
                new_var = cs.alloc_input(lambda: new_v)
            else:
                new_var = cs.alloc(lambda: new_v)
            cs.enforce(LC().add(tmp), LC().add(xl_var).add(ONE, ci),
                       LC().add(new_var).sub(xr_var))
            xr_var, xr_v = xl_var, xl_v
            xl_var, xl_v = new_var, new_v
    return synth


# ---------------------------------------------------------------------------
# Trapdoor validity check: with the toxic waste known the
# Groth16 verification equation can be checked in the exponent, pairing-free:
#   e(A,B) = e(alpha,beta) e(acc,gamma) e(C,delta)
#   <=>  a*b = alpha*beta + acc*gamma + c*delta   (dlogs w.r.t. g1,g2)
# For DummyEngine this is literally verifier.rs:46-52 (pairing = product).
# ---------------------------------------------------------------------------
def dummy_verify(prm, proof, public_inputs):
    fr = DummyEngine.fr
    if len(public_inputs) + 1 != len(prm.ic):
        return False
    acc = prm.ic[0]
    for x, b in zip(public_inputs, prm.ic[1:]):
        acc = fr.add(acc, fr.mul(b, x))
    a, b, c = proof
    lhs = fr.mul(a, b)
    rhs = fr.add(fr.add(fr.mul(prm.alpha_g1, prm.beta_g2), fr.mul(acc, prm.gamma_g2)),
                 fr.mul(c, prm.delta_g2))
    return lhs == rhs
