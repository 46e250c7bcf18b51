"""MSMs over resident window multiples (bb_bases_precompute / option msm_precompute) in the host-thread
emulation of the product sources (see test_emulated_pipeline.py): same points, same proofs, same
error semantics as the per-window path."""
import numpy as np
import pytest

import bellman_b200 as bb
from oracle import o1

import test_gpu_parity as G
import test_emulated_pipeline as E
from test_emulated_pipeline import worker                       # noqa: F401  (the emulated-library fixture)


@pytest.fixture()
def precompute(worker):
    worker.set_option("msm_precompute", 1)
    yield worker
    worker.set_option("msm_precompute", 0)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 33, 1000])
def test_emulated_precompute_g1(precompute, n):
    G.test_multiexp_g1_matches_oracle(precompute, n)


@pytest.mark.parametrize("n", [3, 40, 300])
def test_emulated_precompute_g2(precompute, n):
    G.test_multiexp_g2_matches_oracle(precompute, n)


def test_emulated_precompute_variants(precompute):
    E.test_emulated_multiexp_windows(precompute)
    E.test_emulated_multiexp_density_fast_paths_and_skew(precompute)
    G.test_multiexp_error_semantics(precompute)
    # explicit table construction, then queries of different densities / offsets over the same bases
    n = 600
    pts = o1.g1_fixed_mul(o1.fr_random(91, n))
    bases = bb.Bases(precompute, bb.G1, pts).precompute()
    rng = np.random.default_rng(9)
    for off, m in ((0, n), (5, 200), (599, 1)):
        dens = rng.random(m + 50) < 0.8
        dens[np.cumsum(dens) > n - off] = False              # never run past the end
        ex = o1.fr_random(92 + off, m + 50)
        rc, want = o1.multiexp(1, pts, off, dens.astype(np.uint8), ex)
        assert rc == 0
        got = bb.multiexp(precompute, (bases, off), bb.DensityTracker(dens), ex).wait()
        assert np.array_equal(got, want)


def test_emulated_precompute_prove_and_shards(precompute):
    """a 2^7-constraint MiMC proof and its 2-way shards over per-window-slot tables (the 322-round proof and the 8-way
    shards run under the one-bucket-set form below)"""
    import random
    rng = random.Random(72)
    mc = o1.Mimc(60, seed=4)
    mc.set_toxic([rng.randrange(1, o1.FR_MODULUS) for _ in range(5)])
    mc.generate()
    params = bb.Parameters(precompute, mc.export_params())
    asg = G._assignment(mc.witness())
    r, s = rng.randrange(o1.FR_MODULUS), rng.randrange(o1.FR_MODULUS)
    proof = bb.create_proof(asg, params, r, s)
    assert proof == mc.prove(r, s) == mc.expected_proof(r, s)
    parts = []
    for k in range(2):
        pk = bb.Parameters(precompute, mc.export_params(), shard_index=k, shard_count=2)
        parts.append(bb.prove_partials(asg, pk))
        pk.free()
    assert bb.finalize(params, parts, r, s) == proof


@pytest.mark.parametrize("variant", [33])
def test_emulated_accumulate_variants(worker, variant):
    """msm_acc_variant: round 1's launch-bound / prefetch variants of the accumulate kernel were measured (no gain) and removed;
    the key is still accepted and changes nothing"""
    worker.set_option("msm_acc_variant", variant)
    try:
        for pre in (0, 1):
            worker.set_option("msm_precompute", pre)
            G.test_multiexp_g1_matches_oracle(worker, 1000)
            G.test_multiexp_g2_matches_oracle(worker, 40)
            G.test_multiexp_error_semantics(worker)
    finally:
        worker.set_option("msm_acc_variant", 0)
        worker.set_option("msm_precompute", 0)


# ---- msm_precompute = 2: one bucket set for all windows ------------------------------------------------
@pytest.fixture()
def unified(worker):
    worker.set_option("msm_precompute", 2)
    yield worker
    worker.set_option("msm_precompute", 0)
    worker.set_option("msm_unified_rows_log", 3)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 33, 1000])
def test_emulated_unified_g1(unified, n):
    G.test_multiexp_g1_matches_oracle(unified, n)


@pytest.mark.parametrize("n", [3, 40, 300])
def test_emulated_unified_g2(unified, n):
    G.test_multiexp_g2_matches_oracle(unified, n)


@pytest.mark.parametrize("rounds", [1, 2, 4, 6])
def test_emulated_unified_forced_rounds(unified, rounds):
    unified.set_option("msm_affine_rounds", rounds)
    try:
        G.test_multiexp_g1_matches_oracle(unified, 1000)
        G.test_multiexp_g2_matches_oracle(unified, 40)
        G.test_multiexp_error_semantics(unified)
    finally:
        unified.set_option("msm_affine_rounds", -1)


def test_emulated_unified_variants(unified):
    E.test_emulated_multiexp_windows(unified)
    E.test_emulated_multiexp_density_fast_paths_and_skew(unified)
    G.test_multiexp_error_semantics(unified)


def test_emulated_tables_come_and_go(worker):
    """bb_bases_precompute / bb_bases_drop_table / the option: whichever way the table of a base vector appears or
    disappears between two MSMs, the point is the same"""
    n = 300
    pts = o1.g1_fixed_mul(o1.fr_random(95, n))
    ex = o1.fr_random(96, n)
    rc, want = o1.multiexp(1, pts, 0, None, ex)
    bases = bb.Bases(worker, bb.G1, pts)
    try:
        for prepare in (lambda: None, bases.precompute, bases.drop_table,
                        lambda: worker.set_option("msm_precompute", 2), bases.drop_table,           # rebuilt on first use
                        lambda: worker.set_option("msm_precompute_groups", 2),                       # G2 only: this G1 table is left alone
                        bases.drop_table, lambda: worker.set_option("msm_precompute", 1), bases.drop_table):
            prepare()
            assert np.array_equal(bb.multiexp(worker, (bases, 0), bb.FullDensity, ex).wait(), want)
    finally:
        worker.set_option("msm_precompute", 0)
        worker.set_option("msm_precompute_groups", 3)
        bases.free()


def test_emulated_unified_deep_rounds(unified):
    G.unified_deep_rounds_case(unified, 1 << 13)            # the rounds chosen by the fill: 8 of them


def test_emulated_unified_prove_and_shards(unified):
    E.test_emulated_prove_mimc322_and_shards(unified)
    G.test_prove_begin_end_with_coset_evaluations(unified)


def test_emulated_autotune(worker):
    G.autotune_case(worker, 60)
    assert bb.load_library().bb_tuning_name(99) is None
