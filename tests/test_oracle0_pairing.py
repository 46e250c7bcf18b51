"""The pairing-based verifier of the oracle: `verify_proof`
(groth16/src/verifier.rs:23-58) over a Python-integer BLS12-381 pairing, and `Proof::read`
(groth16/src/lib.rs:47-99).  With it the harness can assert what groth16/tests/mimc.rs:88-92
asserts -- prove, write, read, verify -- without knowing the toxic waste."""
import random

import pytest

from oracle import o1
from oracle.oracle0 import fields as F
from oracle.oracle0 import pairing as PR

R = F.FR_MODULUS


def test_pairing_is_bilinear_and_nondegenerate():
    q12 = PR.twist(F.G2_GEN)
    assert PR._on_curve12(q12) and PR._on_curve12(PR.cast_g1(F.G1_GEN))
    a = PR.Fp12(list(range(3, 15)))
    assert a * a.inv() == PR.Fp12.one()
    e = PR.pairing(F.G1_GEN, F.G2_GEN)
    assert not e == PR.Fp12.one()
    assert e.pow(R) == PR.Fp12.one()
    k1, k2 = 0x1234567, 0x7654321
    assert PR.pairing(F.G1.mul(F.G1_GEN, k1), F.G2.mul(F.G2_GEN, k2)) == e.pow(k1 * k2 % R)
    assert PR.pairing(None, F.G2_GEN) == PR.Fp12.one() and PR.pairing(F.G1_GEN, None) == PR.Fp12.one()
    assert PR.multi_pairing([(F.G1_GEN, F.G2_GEN), (F.G1.neg(F.G1_GEN), F.G2_GEN)]) == PR.Fp12.one()


def test_point_decoding_round_trips():
    rng = random.Random(3)
    for _ in range(4):
        p = F.G1.mul(F.G1_GEN, rng.randrange(1, R))
        q = F.G2.mul(F.G2_GEN, rng.randrange(1, R))
        assert PR.g1_decompress(F.g1_compress(p)) == p
        assert PR.g2_decompress(F.g2_compress(q)) == q
    assert PR.g1_decompress(F.g1_compress(None)) is None and PR.g2_decompress(F.g2_compress(None)) is None
    bad = bytearray(F.g1_compress(F.G1_GEN)); bad[47] ^= 1
    with pytest.raises(ValueError):
        PR.g1_decompress(bytes(bad))                      # x+1 is (almost surely) off the curve or off the subgroup
    with pytest.raises(ValueError):
        PR.proof_read(F.g1_compress(None) + F.g2_compress(F.G2_GEN) + F.g1_compress(F.G1_GEN))   # lib.rs:59-69


def _vk_from_params(p):
    g1 = o1.g1_to_affine_ints(p["vk_g1"])      # alpha, beta, delta
    g2 = o1.g2_to_affine_ints(p["vk_g2"])      # beta, gamma, delta
    return dict(alpha_g1=g1[0], beta_g2=g2[0], gamma_g2=g2[1], delta_g2=g2[2], ic=o1.g1_to_affine_ints(p["ic"]))


def test_mimc_prove_write_read_verify():
    """groth16/tests/mimc.rs:23-103 with the CPU oracle as prover and the pairing verifier."""
    rng = random.Random(9)
    mc = o1.Mimc(322, seed=21)
    mc.set_toxic([rng.randrange(1, R) for _ in range(5)])
    mc.generate()
    p = mc.export_params()
    vk = _vk_from_params(p)
    image = o1.fr_to_ints(mc.witness()["inputs"])[1]
    r, s = rng.randrange(R), rng.randrange(R)
    proof_bytes = mc.prove(r, s)
    assert proof_bytes == mc.expected_proof(r, s)
    proof = PR.proof_read(proof_bytes)
    assert PR.verify_proof(vk, proof, [image])
    assert not PR.verify_proof(vk, proof, [(image + 1) % R])                 # wrong public input
    a, b, c = proof
    assert not PR.verify_proof(vk, (a, b, F.G1.add(c, F.G1_GEN)), [image])   # tampered proof
    with pytest.raises(ValueError):
        PR.verify_proof(vk, proof, [])                                       # InvalidVerifyingKey, verifier.rs:28-30
