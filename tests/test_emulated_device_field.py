"""The DEVICE arithmetic path of bellman_b200/csrc/{mp,field,curve}.cuh, executed on the CPU.

tests/native/emu_field.cpp is compiled with g++ -DBB_EMULATE_PTX: the PTX carry-chain primitives
(mad.lo.cc / madc.hi.cc / addc / subc ...) are modelled with an explicit carry flag, everything
above them is the template code the kernels instantiate -- the merged two-accumulator Montgomery
product, wide_mul / wide_sqr / redc_wide (dedicated Fp squaring), and the lazily reduced Fp2
product.  Checked against Python integers (oracle0 constants) on edge values and random draws.
The GPU parity tests remain the gate for what nvcc makes of the same code.
"""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.oracle0 import fields as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_MOD, P_MOD = F.FR_MODULUS, F.FP_MODULUS


# (BB_FP_WIDE_SQR, BB_FP2_LAZY): the shipped default (0, 0) and the alternative arithmetic variants
@pytest.fixture(scope="module", params=[(0, 0), (1, 0), (1, 1)], ids=["default", "wide-sqr", "wide-sqr+lazy-fp2"])
def emu(request, tmp_path_factory):
    sqr, lazy = request.param
    out = tmp_path_factory.mktemp("emu") / f"libemu_field_{sqr}{lazy}.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-DBB_EMULATE_PTX", f"-DBB_FP_WIDE_SQR={sqr}", f"-DBB_FP2_LAZY={lazy}",
                    "-I", os.path.join(ROOT, "bellman_b200/csrc"),
                    "-shared", "-fPIC", os.path.join(ROOT, "tests/native/emu_field.cpp"), "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def pack(vals, nl):
    return (ctypes.c_uint32 * (nl * len(vals)))(*[(v >> (32 * k)) & 0xFFFFFFFF for v in vals for k in range(nl)])


def unpack(buf, nl):
    w = list(buf)
    return [sum(w[i * nl + k] << (32 * k) for k in range(nl)) for i in range(len(w) // nl)]


def cases(q, nl, n_random, seed):
    rng = random.Random(seed)
    R = 1 << (32 * nl)
    edge = [0, 1, 2, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2, R % q, (R * R) % q, (1 << 32) - 1, (1 << 64) - 1,
            q - (1 << 32), q >> 1, (1 << (q.bit_length() - 1))]
    edge = [e % q for e in edge]
    pairs = [(x, y) for x in edge for y in edge]
    pairs += [(rng.randrange(q), rng.randrange(q)) for _ in range(n_random)]
    return pairs


def expect(op, x, y, q, Rinv):
    return [(x + y) % q, (x - y) % q, x * y * Rinv % q, x * x * Rinv % q, (-x) % q, 2 * x % q][op]


@pytest.mark.parametrize("name,q,nl", [("fr", R_MOD, 8), ("fp", P_MOD, 12)])
def test_prime_field_device_path(emu, name, q, nl):
    Rinv = pow(1 << (32 * nl), -1, q)
    pairs = cases(q, nl, 3000, 11)
    a, b = pack([p[0] for p in pairs], nl), pack([p[1] for p in pairs], nl)
    fn = getattr(emu, f"emu_{name}_op")
    for op in range(6):
        out = (ctypes.c_uint32 * (nl * len(pairs)))()
        fn(op, a, b, out, ctypes.c_size_t(len(pairs)))
        got = unpack(out, nl)
        for (x, y), g in zip(pairs, got):
            assert g == expect(op, x, y, q, Rinv), (name, op, hex(x), hex(y))


def test_fp2_device_path(emu):
    q, nl = P_MOD, 12
    Rinv = pow(1 << 384, -1, q)
    rng = random.Random(5)
    edge = [0, 1, q - 1, q - 2, (q - 1) // 2, (1 << 384) % q, (1 << 380)]
    quads = [(a0, a1, b0, b1) for a0 in edge for a1 in edge for b0 in edge[:4] for b1 in edge[2:6]]
    quads += [tuple(rng.randrange(q) for _ in range(4)) for _ in range(3000)]
    a = pack([v for t in quads for v in t[:2]], nl)
    b = pack([v for t in quads for v in t[2:]], nl)
    for op in range(6):
        out = (ctypes.c_uint32 * (24 * len(quads)))()
        emu.emu_fp2_op(op, a, b, out, ctypes.c_size_t(len(quads)))
        got = unpack(out, nl)
        for i, (a0, a1, b0, b1) in enumerate(quads):
            if op == 2:      # Montgomery product in Fp[u]/(u^2+1)
                e = ((a0 * b0 - a1 * b1) * Rinv % q, (a0 * b1 + a1 * b0) * Rinv % q)
            elif op == 3:
                e = ((a0 * a0 - a1 * a1) * Rinv % q, 2 * a0 * a1 * Rinv % q)
            else:
                e = (expect(op, a0, b0, q, Rinv), expect(op, a1, b1, q, Rinv))
            assert (got[2 * i], got[2 * i + 1]) == e, (op, i)


def test_wide_primitives(emu):
    rng = random.Random(9)
    for nl, q, sfx in ((12, P_MOD, "12"), (8, R_MOD, "8")):
        R = 1 << (32 * nl)
        Rinv = pow(R, -1, q)
        vals = [0, 1, R - 1, q, q - 1, 2 * q - 1, R >> 1, (1 << 32) - 1] + [rng.randrange(R) for _ in range(400)]
        for x in vals:
            for y in (vals[rng.randrange(len(vals))], R - 1, 0):
                out = (ctypes.c_uint32 * (2 * nl))()
                getattr(emu, "emu_wide_mul" + sfx)(pack([x], nl), pack([y], nl), out)
                assert unpack(out, 2 * nl)[0] == x * y
            out = (ctypes.c_uint32 * (2 * nl))()
            getattr(emu, "emu_wide_sqr" + sfx)(pack([x], nl), out)
            assert unpack(out, 2 * nl)[0] == x * x
        # redc_wide: any T < 2 q^2
        for t in [0, 1, q * q, 2 * q * q - 1, q * R // 4, R - 1, R] + [rng.randrange(2 * q * q) for _ in range(2000)]:
            out = (ctypes.c_uint32 * nl)()
            getattr(emu, "emu_redc_wide" + sfx)(pack([t], 2 * nl), out)
            assert unpack(out, nl)[0] == t * Rinv % q, hex(t)


def test_point_sums_through_device_field(emu):
    """XYZZ mixed additions (incl. doubling, inverse and identity operands) over the emulated field"""
    rng = random.Random(3)
    Rp = (1 << 384) % P_MOD

    def m(v):
        return v * Rp % P_MOD

    g1 = [F.G1.mul(F.G1_GEN, rng.randrange(1, R_MOD)) for _ in range(6)]
    seq1 = [g1[0], g1[1], g1[1], None, g1[2], F.G1.neg(g1[2]), g1[3], g1[3], g1[3], g1[4], g1[5]]
    want = None
    for p in seq1:
        want = F.G1.add(want, p)
    flat = [m(c) for p in seq1 for c in (p if p is not None else (0, 0))]
    out = (ctypes.c_uint32 * 24)()
    emu.emu_g1_sum(pack(flat, 12), ctypes.c_size_t(len(seq1)), out)
    assert tuple(unpack(out, 12)) == (m(want[0]), m(want[1]))
    # acc == P then + P (doubling branch), then - 2P (identity), from the identity again
    seq = [g1[0], g1[0], F.G1.neg(F.G1.add(g1[0], g1[0])), g1[1]]
    flat = [m(c) for p in seq for c in p]
    emu.emu_g1_sum(pack(flat, 12), ctypes.c_size_t(len(seq)), out)
    assert tuple(unpack(out, 12)) == (m(g1[1][0]), m(g1[1][1]))

    g2 = [F.G2.mul(F.G2_GEN, rng.randrange(1, R_MOD)) for _ in range(4)]
    seq2 = [g2[0], g2[1], g2[1], None, g2[2], F.G2.neg(g2[2]), g2[3], g2[3], g2[3]]
    want = None
    for p in seq2:
        want = F.G2.add(want, p)
    flat = []
    for p in seq2:
        if p is None:
            flat += [0, 0, 0, 0]
        else:
            flat += [m(p[0][0]), m(p[0][1]), m(p[1][0]), m(p[1][1])]
    out = (ctypes.c_uint32 * 48)()
    emu.emu_g2_sum(pack(flat, 12), ctypes.c_size_t(len(seq2)), out)
    assert tuple(unpack(out, 12)) == (m(want[0][0]), m(want[0][1]), m(want[1][0]), m(want[1][1]))
