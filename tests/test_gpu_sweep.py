"""Parity of the Neal-8 sweep kernels with the CPU oracle on the same seeded inputs.
Integer outputs (mismatch counts, histograms, allocations under an injected uniform tape, centres)
are bit-exact; log-likelihoods agree to 1e-12 relative (fp64); sigma draws agree with the reference's
bisection to its own bracket width (1e-9 in u)."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem, oracle_state_full, rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # n, p, m, K_true  (ragged p, n not a multiple of the tile, K across tile edges)
    (257, 13, 3, 2),
    (1000, 64, 4, 10),
    (515, 100, 5, 33),
    (300, 257, 2, 5),
]


@pytest.mark.parametrize("n,p,m,k", SHAPES)
def test_ll_block_matches_oracle(n, p, m, k):
    pb = Problem(n, p, m, k, seed=n + p)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    LL, mm = ch.ll_block(K)
    LLo, mmo = orc.ll_block(pb.od, cen, sig)
    assert np.array_equal(mm, mmo)  # integer mismatch counts: bit-exact
    assert np.max(rel_err(LL, LLo)) < 1e-12
    assert abs(ch.loglik() - orc.loglik(pb.od, c, cen, sig)) <= 1e-12 * abs(orc.loglik(pb.od, c, cen, sig))
    ch.close()


def test_ll_block_mixed_attribute_sizes():
    rng = np.random.default_rng(2)
    attr = rng.integers(2, 7, 40)
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    pb = Problem(400, 40, 3, 4, seed=3)
    X, lab, cent, at = ham_mix_gen(400, 40, attr, 4, seed=3)
    pb.X, pb.labels, pb.cent, pb.attr = X, lab, cent, at
    pb.od = orc.OracleData(X, at, pb.gamma, pb.v, pb.w)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    LL, mm = ch.ll_block(K)
    LLo, mmo = orc.ll_block(pb.od, cen, sig)
    assert np.array_equal(mm, mmo)
    assert np.max(rel_err(LL, LLo)) < 1e-12
    ch.close()


def _scan_case(pb, mode, seed, m_aux=3, pool=97, L=None, iters=1):
    K, c, cen, sig = oracle_state_full(pb, mode=mode, seed=seed, L=L, iters=iters, m_aux=m_aux)
    pc, ps = orc.draw_pool(pb.od, pool, seed + 1, o=orc.opts(stable_hig=1))
    rng = np.random.default_rng(seed)
    tape = (rng.integers(0, 2**53, size=pb.n * (m_aux + 1)).astype(np.float64) + 0.5) / 2.0**53
    ref = orc.neal8_scan(pb.od, m_aux, c, cen, sig, pc, ps, tape, o=orc.opts(counted=1))
    ch = pb.chain(m=m_aux)
    ch.set_state(K, c, cen, sig)
    ch.set_pool(pc, ps)
    ch.neal8_scan(tape)
    got = ch.snapshot()
    st = ch.stats()
    ch.close()
    return ref, got, st


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_scan_quiet_regime_bit_exact(seed):
    pb = Problem(3000, 64, 4, 10, seed=seed)
    ref, got, st = _scan_case(pb, "truth", seed)
    assert got["K"] == ref["K"]
    assert np.array_equal(got["c_i"], ref["c"])
    assert np.array_equal(got["centers"], ref["center"])
    assert np.array_equal(got["sigmas"], ref["sigma"])


@pytest.mark.parametrize("seed", [4, 5, 6])
def test_scan_burn_in_regime_bit_exact(seed):
    # random labels + diffuse data: many moves, births, deaths and singleton replacements (cases 1-4)
    pb = Problem(1500, 24, 3, 6, seed=seed, s=1.5)
    ref, got, st = _scan_case(pb, "random", seed, L=12, iters=1)
    assert st["scan_events"] > 100
    assert got["K"] == ref["K"]
    assert np.array_equal(got["c_i"], ref["c"])
    assert np.array_equal(got["centers"], ref["center"])
    assert np.array_equal(got["sigmas"], ref["sigma"])


def test_scan_with_many_singletons():
    # every observation starts in its own cluster (zoo_simulator.R L=101 analogue): cases 2 and 4 dominate
    pb = Problem(120, 16, 3, 3, seed=7, s=0.8)
    n = pb.n
    c0 = np.arange(n, dtype=np.int32)
    pc, ps = orc.draw_pool(pb.od, 200, 3, o=orc.opts(stable_hig=1))
    cen, sig = pc[:n].copy(), ps[:n].copy()
    rng = np.random.default_rng(11)
    tape = (rng.integers(0, 2**53, size=n * 4).astype(np.float64) + 0.5) / 2.0**53
    ref = orc.neal8_scan(pb.od, 3, c0, cen, sig, pc, ps, tape, o=orc.opts(counted=0), kcap=256)
    ch = pb.chain(m=3, c_i=c0, max_clusters=200)
    ch.set_state(n, c0, cen, sig)
    ch.set_pool(pc, ps)
    ch.neal8_scan(tape)
    got = ch.snapshot()
    ch.close()
    assert got["K"] == ref["K"]
    assert np.array_equal(got["c_i"], ref["c"])
    assert np.array_equal(got["sigmas"], ref["sigma"])


def test_histogram_and_update_phi_match_oracle():
    pb = Problem(2000, 37, 5, 7, seed=21)
    K, c, cen, sig = oracle_state_full(pb, mode="random", seed=21, L=9, iters=2)
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    H, cnt = ch.histogram(K)
    Ho, cnto = orc.histogram(pb.od, K, c, H.shape[2])
    assert np.array_equal(H, Ho) and np.array_equal(cnt, cnto)  # integer: bit-exact
    rng = np.random.default_rng(5)
    uc = rng.random((K, pb.p))
    us = rng.random((K, pb.p))
    # oracle consumes, per cluster in label order, p centre uniforms then p sigma uniforms
    tape = np.concatenate([np.concatenate([uc[k], us[k]]) for k in range(K)])
    ref = orc.update_phi(pb.od, c, cen, sig, tape, o=orc.opts(stable_hig=1, sigma_inverse_cdf=1))
    assert ref["consumed"] == tape.size
    ch.update_phi(uc, us)
    got = ch.snapshot()
    assert np.array_equal(got["centers"], ref["center"])  # Rcpp::sample incl. revsort tie order: bit-exact
    u_got, u_ref = np.exp(-1.0 / got["sigmas"]), np.exp(-1.0 / ref["sigma"])
    assert np.max(np.abs(u_got - u_ref)) <= 1e-9  # reference bisection bracket (hyperg.cpp:263)
    ref2 = orc.update_phi(pb.od, c, cen, sig, tape, o=orc.opts(stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0))
    assert np.max(rel_err(got["sigmas"], ref2["sigma"])) < 1e-9
    ch.close()


def test_update_phi_tie_heavy_centres():
    # small clusters with 6-level attributes: most levels are unseen => equal probabilities => the
    # revsort tie permutation decides the draw
    pb = Problem(60, 20, 6, 12, seed=4, s=2.0)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    sig = sig * 4.0  # flatter conditionals: ties carry real mass
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    rng = np.random.default_rng(8)
    uc, us = rng.random((K, pb.p)), rng.random((K, pb.p))
    tape = np.concatenate([np.concatenate([uc[k], us[k]]) for k in range(K)])
    ref = orc.update_phi(pb.od, c, cen, sig, tape, o=orc.opts(stable_hig=1, sigma_inverse_cdf=1))
    ch.update_phi(uc, us)
    got = ch.snapshot()
    assert np.array_equal(got["centers"], ref["center"])
    ch.close()


def test_full_size_parity_config_c2():
    # BASELINE.json config 2: n=1e4, p=64, m=4, K_true=10, injected uniform stream, bit-exact allocations
    for seed in (1, 2, 3):
        pb = Problem(10000, 64, 4, 10, seed=seed)
        ref, got, st = _scan_case(pb, "truth", seed, pool=997)
        assert got["K"] == ref["K"]
        assert np.array_equal(got["c_i"], ref["c"])


def _assert_scan_equal(ref, got):
    assert got["K"] == ref["K"]
    assert np.array_equal(got["c_i"], ref["c"])
    assert np.array_equal(got["centers"], ref["center"])
    assert np.array_equal(got["sigmas"], ref["sigma"])


@pytest.mark.parametrize("seed", [31, 32])
def test_scan_more_than_64_entries(seed):
    # K + m_aux > 64: the 8-entries-per-lane evaluation path; diffuse data so that draws are real decisions
    pb = Problem(2800, 32, 4, 70, seed=seed, s=1.1)
    ref, got, st = _scan_case(pb, "truth", seed, pool=211)
    assert ref["K"] + 3 > 64 and st["scan_events"] > 20
    _assert_scan_equal(ref, got)


@pytest.mark.parametrize("seed", [41, 42])
def test_scan_and_update_phi_with_12_levels(seed):
    # more than 7 levels: C-form likelihood kernel, global-atomic histogram, general centre draw
    pb = Problem(1500, 40, 12, 6, seed=seed, s=1.0)
    ref, got, st = _scan_case(pb, "random", seed, L=9, iters=1)
    assert st["scan_events"] > 50
    _assert_scan_equal(ref, got)
    K, c, cen, sig = oracle_state_full(pb, mode="random", seed=seed, L=9, iters=2)
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    LL, mm = ch.ll_block(K)
    LLo, mmo = orc.ll_block(pb.od, cen, sig)
    assert np.array_equal(mm, mmo) and np.max(rel_err(LL, LLo)) < 1e-12
    H, cnt = ch.histogram(K)
    Ho, cnto = orc.histogram(pb.od, K, c, H.shape[2])
    assert np.array_equal(H, Ho) and np.array_equal(cnt, cnto)
    rng = np.random.default_rng(seed)
    uc, us = rng.random((K, pb.p)), rng.random((K, pb.p))
    tape = np.concatenate([np.concatenate([uc[k], us[k]]) for k in range(K)])
    refp = orc.update_phi(pb.od, c, cen, sig, tape, o=orc.opts(stable_hig=1, sigma_inverse_cdf=1))
    ch.update_phi(uc, us)
    assert np.array_equal(ch.snapshot()["centers"], refp["center"])
    ch.close()


def test_scan_tiny_problem_one_aux():
    pb = Problem(40, 8, 3, 2, seed=51, s=1.0)
    ref, got, st = _scan_case(pb, "random", 51, m_aux=1, pool=17, L=3)
    _assert_scan_equal(ref, got)


def test_scan_births_beyond_the_materialised_columns():
    # every observation its own cluster and only 4 spare LL columns (max_clusters = n + 4): the clusters opened by
    # cases 3 and 4 soon get slots >= ldl, whose likelihoods are evaluated on the fly instead of being read
    pb = Problem(100, 16, 3, 3, seed=61, s=0.8)
    n = pb.n
    c0 = np.arange(n, dtype=np.int32)
    pc, ps = orc.draw_pool(pb.od, 150, 62, o=orc.opts(stable_hig=1))
    cen, sig = pc[:n].copy(), ps[:n].copy()
    rng = np.random.default_rng(61)
    tape = (rng.integers(0, 2**53, size=n * 4).astype(np.float64) + 0.5) / 2.0**53
    ref = orc.neal8_scan(pb.od, 3, c0, cen, sig, pc, ps, tape, o=orc.opts(counted=0), kcap=256)
    ch = pb.chain(m=3, c_i=c0, max_clusters=n + 4)
    ch.set_state(n, c0, cen, sig)
    ch.set_pool(pc, ps)
    ch.neal8_scan(tape)
    got = ch.snapshot()
    st = ch.stats()
    ch.close()
    assert st["scan_events"] > 30
    assert got["K"] == ref["K"]
    assert np.array_equal(got["c_i"], ref["c"])
    assert np.array_equal(got["sigmas"], ref["sigma"])


@pytest.mark.parametrize("L", [1, 2, 7, 50, 101, 1000])
def test_initial_assignment_under_injected_uniforms(L):
    # a4 sample_initial_assignment (common_functions.cpp:174-183): Rcpp::sample(L, n, true) - 1 = (int)(L u + 1) - 1
    from split_and_merge_gibbs_sampling_b200.api import initial_assignment
    rng = np.random.default_rng(L)
    n = 5000
    u = (rng.integers(0, 2**53, size=n).astype(np.float64) + 0.5) / 2.0**53
    # both ends of R's unif_rand range (its fixup keeps draws inside [2.33e-10, 1 - 2.33e-10]) and a bin edge
    u[:4] = [2.328306437080797e-10, 1.0 - 2.328306437080797e-10, 0.5, 1.0 / L if L > 1 else 0.25]
    got = initial_assignment(u, L)
    ref = orc.initial_assignment(L, u)
    assert np.array_equal(got, ref)
    assert got.min() >= 0 and got.max() <= L - 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_config_c2_full_size_from_a_random_start(seed):
    # BASELINE.json config 2 (n=1e4, p=64, m=4, K_true=10) in the regime where the draws are real decisions: the state
    # is one oracle iteration away from L=20 random labels, so the pass moves thousands of observations, opens and
    # closes clusters; allocations, K and the parameters carried by the births are bit-exact under the injected stream
    pb = Problem(10000, 64, 4, 10, seed=seed)
    ref, got, st = _scan_case(pb, "random", seed, pool=997, L=20, iters=1)
    assert st["scan_events"] > 1000
    _assert_scan_equal(ref, got)


@pytest.mark.parametrize("mode,seed", [("truth", 1), ("random", 2)])
def test_config_c3_shape_reduced_n(mode, seed):
    # BASELINE.json config 3 (MNIST-shaped: p=784, m=2) at n=2500: pp=784 is 12 full 64-attribute tiles + a 16-wide one
    pb = Problem(2500, 784, 2, 12, seed=seed, s=0.9)
    K, c, cen, sig = oracle_state_full(pb, mode=mode, seed=seed, L=15, iters=1)
    ch = pb.chain()
    ch.set_state(K, c, cen, sig)
    LL, mm = ch.ll_block(K)
    LLo, mmo = orc.ll_block(pb.od, cen, sig)
    assert np.array_equal(mm, mmo)
    assert np.max(rel_err(LL, LLo)) < 1e-12
    H, cnt = ch.histogram(K)
    Ho, cnto = orc.histogram(pb.od, K, c, H.shape[2])
    assert np.array_equal(cnt, cnto) and np.array_equal(H, Ho)
    ch.close()
    ref, got, st = _scan_case(pb, mode, seed, pool=301, L=15, iters=1)
    if mode == "random":
        assert st["scan_events"] > 200
    _assert_scan_equal(ref, got)


@pytest.mark.parametrize("mode,seed", [("truth", 3), ("random", 4)])
def test_config_c4_shape_reduced_n(mode, seed):
    # BASELINE.json config 4 (p=256, m=5, K~100, 3 auxiliaries) at n=6000: K + m_aux > 64 (8 entries per lane in the
    # exact evaluation), LL block 2 column groups wide
    pb = Problem(6000, 256, 5, 100, seed=seed, s=0.9)
    ref, got, st = _scan_case(pb, mode, seed, pool=499, L=100, iters=1)
    assert got["K"] > 64
    if mode == "random":
        assert st["scan_events"] > 500
    _assert_scan_equal(ref, got)
