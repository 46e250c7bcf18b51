"""groth16::generate_parameters on the device.

Mirrors /root/reference/groth16/src/generator.rs:159-507 from the point where the circuit has
been synthesised into a KeypairAssembly (:43-155, :188-191):

  * the powers of tau (:249-264) and the H-query scalars tau^i t(tau)/delta (:267-296) are
    produced by the domain kernels (`distribute_powers`, scale);
  * the Lagrange coefficients L_j(tau) come from one inverse NTT (:300);
  * every group element -- h, a, b_g1, b_g2, ic/l and the verifying key -- is a fixed-base
    multiplication on the device (`bb_fixed_base_mul`; the reference uses a wNAF table per thread
    and batch normalisation, :209-226,271-296,310-415);
  * the per-variable QAP evaluations (:376-415) are sparse dot products with L_j(tau); they are
    O(non-zeros) field operations and stay on the host, as integers.

g1 and g2 are the standard generators scaled by `g1_scalar` / `g2_scalar` (generate_random_parameters
draws random generators, :21-40; a random multiple of a fixed generator is the same distribution).
The result maps the names of groth16::Parameters (vk_g1 = alpha, beta, delta; vk_g2 = beta, gamma,
delta; ic, h, l, a, b_g1, b_g2) to Montgomery-limb arrays, the layout `Parameters(worker, p)` and
`params_io.write_parameters` take.
"""
import numpy as np

from . import (EvaluationDomain, G1, G2, FORM_CANONICAL, SynthesisError, UnexpectedIdentity, fixed_base_mul)

FR_MODULUS = EvaluationDomain.FR_MODULUS
_R = 1 << 256
_R_INV = pow(_R, -1, FR_MODULUS)


class UnconstrainedVariable(SynthesisError):
    """SynthesisError::UnconstrainedVariable (generator.rs:466-470)"""


class KeypairAssembly:
    """generator.rs:43-155: per variable, the (coefficient, constraint index) pairs of the A, B and
    C matrices.  Variables are ("input", i) or ("aux", i); linear combinations are lists of
    (variable, coefficient)."""

    def __init__(self):
        self.num_inputs = self.num_aux = self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self):                                          # :58-75
        self.num_aux += 1
        for m in (self.at_aux, self.bt_aux, self.ct_aux):
            m.append([])
        return ("aux", self.num_aux - 1)

    def alloc_input(self):                                    # :77-94
        self.num_inputs += 1
        for m in (self.at_inputs, self.bt_inputs, self.ct_inputs):
            m.append([])
        return ("input", self.num_inputs - 1)

    def enforce(self, a, b, c):                               # :96-138
        for lc, inputs, aux in ((a, self.at_inputs, self.at_aux), (b, self.bt_inputs, self.bt_aux), (c, self.ct_inputs, self.ct_aux)):
            for (kind, idx), coeff in lc:
                (inputs if kind == "input" else aux)[idx].append((coeff % FR_MODULUS, self.num_constraints))
        self.num_constraints += 1


def _to_limbs(values):
    out = np.zeros((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def _mont_to_ints(arr):
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [(int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) * _R_INV % FR_MODULUS for r in a]


def generate_parameters(worker, assembly, alpha, beta, gamma, delta, tau, g1_scalar=1, g2_scalar=1):
    """`assembly`: the KeypairAssembly as `circuit.synthesize` leaves it (ONE already allocated as
    input 0, generator.rs:188-191).  alpha..tau: integers mod r."""
    q = FR_MODULUS
    import copy
    asm = copy.deepcopy(assembly)                             # the caller's assembly is left as synthesis made it
    for i in range(asm.num_inputs):                           # x_i * 0 = 0, :195-202
        asm.enforce([(("input", i), 1)], [], [])
    dom = EvaluationDomain.from_coeffs(worker, np.zeros((asm.num_constraints, 4), dtype=np.uint64))   # :205-206
    m = dom.coeffs.shape[0]
    if gamma % q == 0 or delta % q == 0:                      # :228-243
        raise UnexpectedIdentity("gamma or delta is not invertible")
    gamma_inverse, delta_inverse = pow(gamma, -1, q), pow(delta, -1, q)
    g1s, g2s = g1_scalar % q, g2_scalar % q

    def g1_mul(ints):
        return fixed_base_mul(worker, G1, _to_limbs([v * g1s % q for v in ints]), FORM_CANONICAL)

    def g2_mul(ints):
        return fixed_base_mul(worker, G2, _to_limbs([v * g2s % q for v in ints]), FORM_CANONICAL)

    # powers of tau (:249-264), then the H query g1^(tau^i t(tau)/delta) for i < m-1 (:267-296)
    dom.coeffs[:] = EvaluationDomain._to_mont(1)
    dom.distribute_powers(tau)
    hq = EvaluationDomain(worker, dom.coeffs.copy(), dom.exp)
    hq._pointwise(2, k=dom.z(tau) * delta_inverse * g1s % q)
    h = fixed_base_mul(worker, G1, hq.coeffs[: m - 1])        # Montgomery scalars straight from the kernel
    dom.ifft()                                                # Lagrange coefficients at tau, :300
    lag = _mont_to_ints(dom.coeffs)

    def eval_at_tau(terms):                                   # :376-389
        acc = 0
        for coeff, index in terms:
            acc += lag[index] * coeff
        return acc % q

    def evaluate(at, bt, ct, inv):                            # :310-415
        a_t = [eval_at_tau(t) for t in at]
        b_t = [eval_at_tau(t) for t in bt]
        c_t = [eval_at_tau(t) for t in ct]
        ext = [(a * beta + b * alpha + c) % q * inv % q for a, b, c in zip(a_t, b_t, c_t)]
        return a_t, b_t, ext

    a_in, b_in, ic_k = evaluate(asm.at_inputs, asm.bt_inputs, asm.ct_inputs, gamma_inverse)
    a_aux, b_aux, l_k = evaluate(asm.at_aux, asm.bt_aux, asm.ct_aux, delta_inverse)
    if any(v == 0 for v in l_k):                              # an identity in L, :466-470
        raise UnconstrainedVariable("auxiliary variable %d is unconstrained" % l_k.index(0))
    a_k = [v for v in a_in + a_aux if v]                      # identities are filtered out, :491-505
    b_k = [v for v in b_in + b_aux if v]
    return dict(
        vk_g1=g1_mul([alpha, beta, delta]), vk_g2=g2_mul([beta, gamma, delta]),
        ic=g1_mul(ic_k), h=h, l=g1_mul(l_k), a=g1_mul(a_k), b_g1=g1_mul(b_k), b_g2=g2_mul(b_k),
        num_inputs=asm.num_inputs, num_aux=asm.num_aux, m=m)
