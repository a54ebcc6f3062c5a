"""Behavioural checks of the oracle's restated third-party pieces (Rcpp sample, R revsort, rbeta,
rhig) and of its two execution modes."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import oracle_lib as orc
from helpers import Problem

L = orc.lib()


def test_revsort_descending_and_permutation():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5, 6, 17, 64):
        a = rng.random(n)
        a0 = a.copy()
        ib = np.arange(1, n + 1, dtype=np.int32)
        L.orc_revsort(orc.P(a), orc.P(ib, orc.ip), n)
        assert np.all(np.diff(a) <= 0)
        assert np.array_equal(a0[ib - 1], a)
        assert sorted(ib) == list(range(1, n + 1))


def test_revsort_two_element_tie_puts_second_first():
    a = np.array([0.5, 0.5])
    ib = np.array([1, 2], dtype=np.int32)
    L.orc_revsort(orc.P(a), orc.P(ib, orc.ip), 2)
    assert list(ib) == [2, 1]


def test_sample_probs_walks_descending_order():
    probs = np.array([0.1, 0.6, 0.3])
    fl = np.zeros(2, dtype=np.int32)
    assert L.orc_sample_probs_one(orc.P(probs), 3, 0.59, orc.P(fl, orc.ip)) == 1
    assert L.orc_sample_probs_one(orc.P(probs), 3, 0.61, orc.P(fl, orc.ip)) == 2
    assert L.orc_sample_probs_one(orc.P(probs), 3, 0.91, orc.P(fl, orc.ip)) == 0
    # unnormalised input is re-normalised (Rcpp FixProb)
    p2 = probs * 7
    assert L.orc_sample_probs_one(orc.P(p2), 3, 0.61, orc.P(fl, orc.ip)) == 2
    # zero-probability entries are never drawn, not even by the fall-through of the loop
    p3 = np.array([0.0, 1.0, 0.0])
    assert L.orc_sample_probs_one(orc.P(p3), 3, 0.999999, orc.P(fl, orc.ip)) == 1
    # invalid probabilities are an error, like Rcpp::stop
    p4 = np.array([0.2, np.nan])
    assert L.orc_sample_probs_one(orc.P(p4), 2, 0.5, orc.P(fl, orc.ip)) == -1


def test_rbeta_matches_scipy_distribution():
    for (a, b) in [(1.25, 5.0), (0.6, 0.8), (31.25, 75.0), (3.0, 1.5)]:
        out = np.empty(20000)
        L.orc_rbeta_many(a, b, out.size, 7, orc.P(out))
        assert stats.kstest(out, stats.beta(a, b).cdf).pvalue > 1e-3


def test_rhig_two_branches_same_law():
    # the Beta-rejection branch and the inverse-CDF (bisection) branch sample the same distribution
    for (v, w, m) in [(6, 0.25, 2), (6, 0.25, 5), (26, 30.25, 2)]:
        assert L.orc_rhig_beta_branch(v, w, m) == 1
        a = np.empty(6000)
        b = np.empty(6000)
        o1 = orc.opts(sigma_inverse_cdf=0)
        o2 = orc.opts(sigma_inverse_cdf=1)
        assert L.orc_rhig_many(v, w, m, a.size, 3, C.byref(o1), orc.P(a)) == 0
        assert L.orc_rhig_many(v, w, m, b.size, 4, C.byref(o2), orc.P(b)) == 0
        assert stats.ks_2samp(a, b).pvalue > 1e-3


def test_counted_mode_is_bit_identical_to_faithful():
    pb = Problem(300, 12, 4, 4, seed=5)
    kw = dict(m_aux=3, iterations=6, L=6, c_init=None, burnin=2, t=3, r=3, neal8=True, split_merge_=True, seed=99,
              pool_size=50)
    a = orc.run_chain(pb.od, o=orc.opts(counted=0, stable_hig=1), **kw)
    b = orc.run_chain(pb.od, o=orc.opts(counted=1, stable_hig=1), **kw)
    assert np.array_equal(a["c_i"], b["c_i"])
    assert np.array_equal(a["total_cls"], b["total_cls"])
    assert np.array_equal(a["loglikelihood"], b["loglikelihood"])
    assert np.array_equal(a["accepted"], b["accepted"])


def test_stable_hig_matches_faithful_chain_on_small_clusters():
    # where the reference's 2F1 is finite, the log-incomplete-beta form reproduces the same chain
    pb = Problem(120, 8, 3, 3, seed=8)
    kw = dict(m_aux=3, iterations=8, L=4, c_init=None, burnin=2, t=2, r=2, neal8=True, split_merge_=True, seed=5,
              pool_size=40)
    a = orc.run_chain(pb.od, o=orc.opts(stable_hig=0), **kw)
    b = orc.run_chain(pb.od, o=orc.opts(stable_hig=1), **kw)
    assert np.array_equal(a["c_i"], b["c_i"])
    assert np.allclose(a["loglikelihood"], b["loglikelihood"], rtol=1e-12)


def test_faithful_reference_throws_on_large_clusters():
    # 2F1 overflows for ~10^3-member clusters; the reference then throws inside split-merge
    pb = Problem(4000, 16, 5, 2, seed=2)
    with pytest.raises(orc.OracleError):
        for seed in range(1, 6):
            orc.run_chain(pb.od, 3, 2, 2, pb.labels, 0, 1, 1, False, True, seed, o=orc.opts(counted=1, stable_hig=0),
                          pool_size=10)


def test_validate_state_error_on_empty_initial_label():
    pb = Problem(30, 6, 3, 2, seed=1)
    c0 = np.zeros(30, dtype=np.int32)
    c0[:5] = 2  # label 1 is empty -> not contiguous
    with pytest.raises(orc.OracleError, match="State validation failed"):
        orc.run_chain(pb.od, 3, 1, 3, c0, 0, 1, 1, True, False, 1, pool_size=10)


def test_zoo_posterior_is_sane():
    import os
    from split_and_merge_gibbs_sampling_b200.synth import zoo_dataset
    path = os.path.join(os.path.dirname(__file__), "golden", "zoo.data")
    X, attr, v, w, g, gt = zoo_dataset(path)
    assert X.shape == (101, 16)
    assert list(attr) == [2] * 12 + [6] + [2] * 3
    assert list(np.bincount(gt)[1:]) == [41, 20, 5, 13, 4, 8, 10]
    d = orc.OracleData(X, attr, g, v, w)
    from sklearn.metrics import adjusted_rand_score
    r = orc.run_chain(d, 3, 150, 20, None, 250, 10, 10, True, True, seed=5)
    ks = r["total_cls"]
    assert 4 <= np.median(ks) <= 10
    assert adjusted_rand_score(gt, r["final_ass"]) > 0.5
