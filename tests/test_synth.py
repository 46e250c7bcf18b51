"""Host-side logic of the product: the synthetic MiMC-chain witness generator used by bench.py
(bellman_b200/csrc/synth.cu) must record exactly what the oracle's restatement of
ProvingAssignment records (groth16/src/prover.rs:73-145,193-215) for the same seed."""
import ctypes as C

import numpy as np
import pytest

import bellman_b200 as bb
from oracle import o1


@pytest.mark.parametrize("rounds", [1, 2, 322, 4095])
def test_synth_witness_matches_oracle(rounds):
    lib = bb.load_library()
    shape = np.zeros(7, np.uint64)
    assert lib.bb_synth_mimc_shape(C.c_size_t(rounds), shape.ctypes.data_as(C.c_void_p)) == 0
    ni, na, n, m = (int(x) for x in shape[:4])
    mc = o1.Mimc(rounds, seed=1234 + rounds)
    assert (ni, na, n, m) == (mc.num_inputs, mc.num_aux, mc.num_constraints, mc.m)
    assert [int(x) for x in shape[4:]] == [mc.a_aux_total, mc.b_in_total, mc.b_aux_total]
    a, b, c = (np.zeros((n, 4), np.uint64) for _ in range(3))
    inputs, aux = np.zeros((ni, 4), np.uint64), np.zeros((na, 4), np.uint64)
    words = (na + 63) // 64
    ad, bd, bi = np.zeros(words, np.uint64), np.zeros(words, np.uint64), np.zeros(1, np.uint64)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    assert lib.bb_synth_mimc_witness(C.c_size_t(rounds), C.c_uint64(1234 + rounds), p(a), p(b), p(c), p(inputs), p(aux),
                                     p(ad), p(bi), p(bd)) == 0
    w = mc.witness()
    for name, got in (("a", a), ("b", b), ("c", c), ("inputs", inputs), ("aux", aux)):
        assert np.array_equal(got, w[name]), name
    assert np.array_equal(ad, bb.pack_density(w["a_aux_density"])[0])
    assert np.array_equal(bd, bb.pack_density(w["b_aux_density"])[0])
    assert int(bi[0]) == int(bb.pack_density(w["b_input_density"])[0][0])
