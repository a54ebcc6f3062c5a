"""End-to-end: the run_markov_chain mirror (code/launcher.cpp:7-174) on the zoo data and on a synthetic
mixture; posterior summaries agree with the oracle within Monte-Carlo error; argument errors surface."""
import os

import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem

pytestmark = pytest.mark.gpu


def _zoo():
    from split_and_merge_gibbs_sampling_b200.synth import zoo_dataset
    return zoo_dataset(os.path.join(os.path.dirname(__file__), "golden", "zoo.data"))


def test_run_markov_chain_result_shape_matches_reference_list():
    from split_and_merge_gibbs_sampling_b200 import run_markov_chain
    X, attr, v, w, g, gt = _zoo()
    res = run_markov_chain(X, attr, g, v, w, m=3, iterations=20, L=5, burnin=10, t=3, r=3, neal8=True, split_merge=True,
                           seed=7)
    assert set(res) >= {"total_cls", "c_i", "centers", "sigmas", "loglikelihood", "final_ass", "time", "accepted"}
    assert len(res["total_cls"]) == 20 and len(res["c_i"]) == 20 and res["loglikelihood"].shape == (20,)
    for it in range(20):
        K = res["total_cls"][it]
        assert len(res["centers"][it]) == K and len(res["sigmas"][it]) == K
        c = res["c_i"][it]
        assert c.min() == 0 and c.max() == K - 1 and len(np.unique(c)) == K  # validate_state invariant
        for k in range(K):
            assert np.all(res["centers"][it][k] >= 1) and np.all(res["centers"][it][k] <= attr)
            assert np.all(res["sigmas"][it][k] > 0)
    # log-likelihood of the snapshot is reproducible by the oracle from the same state (1e-12 rel)
    d = orc.OracleData(X, attr, g, v, w)
    it = 19
    ll = orc.loglik(d, res["c_i"][it], np.array(res["centers"][it]), np.array(res["sigmas"][it]))
    assert abs(ll - res["loglikelihood"][it]) <= 1e-12 * abs(ll)
    assert np.array_equal(res["final_ass"], res["c_i"][-1])


def test_result_views_outlive_the_result_dict():
    """The allocation trace / centres / sigmas come back as views of the library's buffers (no copy); the buffers must
    live as long as any of the arrays does, and two calls must not share storage."""
    import gc
    from split_and_merge_gibbs_sampling_b200 import run_markov_chain
    X, attr, v, w, g, gt = _zoo()
    kw = dict(m=3, iterations=15, L=5, burnin=5, t=3, r=3, neal8=True, split_merge=True, seed=11)
    res = run_markov_chain(X, attr, g, v, w, **kw)
    keep_c, keep_s = res["c_i"][7], res["sigmas"][7][0]
    ref_c, ref_s = keep_c.copy(), keep_s.copy()
    del res
    gc.collect()
    other = run_markov_chain(X, attr, g, v, w, **dict(kw, seed=12))  # would recycle freed storage
    assert np.array_equal(keep_c, ref_c) and np.array_equal(keep_s, ref_s)
    again = run_markov_chain(X, attr, g, v, w, **kw)
    assert np.array_equal(again["c_i"][7], ref_c) and np.array_equal(again["sigmas"][7][0], ref_s)
    assert not np.array_equal(other["loglikelihood"], again["loglikelihood"])


def test_zoo_posterior_matches_oracle_within_mc_error():
    from sklearn.metrics import adjusted_rand_score
    from split_and_merge_gibbs_sampling_b200 import run_markov_chain
    X, attr, v, w, g, gt = _zoo()
    d = orc.OracleData(X, attr, g, v, w)
    k_gpu, k_cpu, ari_gpu, ari_cpu, ll_gpu, ll_cpu = [], [], [], [], [], []
    for seed in range(1, 6):
        res = run_markov_chain(X, attr, g, v, w, m=3, iterations=300, L=10, burnin=300, t=10, r=10, neal8=True,
                               split_merge=True, seed=seed, c_i=np.arange(101) % 10)
        k_gpu.append(np.mean(res["total_cls"]))
        ari_gpu.append(adjusted_rand_score(gt, res["final_ass"]))
        ll_gpu.append(np.mean(res["loglikelihood"]))
        r = orc.run_chain(d, 3, 300, 10, np.arange(101) % 10, 300, 10, 10, True, True, seed=seed)
        k_cpu.append(np.mean(r["total_cls"]))
        ari_cpu.append(adjusted_rand_score(gt, r["final_ass"]))
        ll_cpu.append(np.mean(r["loglikelihood"]))
    # means over 5 seeds agree within 3 standard errors (plus a small floor)
    def close(a, b, floor):
        se = np.sqrt(np.var(a, ddof=1) / len(a) + np.var(b, ddof=1) / len(b))
        return abs(np.mean(a) - np.mean(b)) <= 3 * se + floor
    assert close(k_gpu, k_cpu, 0.5), (k_gpu, k_cpu)
    assert close(ll_gpu, ll_cpu, 8.0), (ll_gpu, ll_cpu)
    assert np.mean(ari_gpu) > 0.6 and np.mean(ari_cpu) > 0.6


def test_synthetic_mixture_recovers_truth():
    from sklearn.metrics import adjusted_rand_score
    from split_and_merge_gibbs_sampling_b200 import Chain
    pb = Problem(5000, 64, 4, 10, seed=3)
    ch = Chain(pb.X.astype(np.float64), pb.attr, 1.0, pb.v, pb.w, m=3, L=10, neal8=True, split_merge=True, seed=5,
               compact_init=True)
    ch.step(30)
    s = ch.snapshot()
    assert adjusted_rand_score(pb.labels, s["c_i"]) > 0.95
    assert 8 <= s["K"] <= 14
    st = ch.stats()
    assert st["sweeps"] == 30 and st["launches"] > 0
    ch.close()


def test_argument_errors():
    from split_and_merge_gibbs_sampling_b200 import Chain, SmgError
    X = np.array([[1, 2], [2, 1], [1, 1], [2, 2]], dtype=np.float64)
    with pytest.raises(SmgError, match="codes"):
        Chain(np.array([[1, 3], [2, 1], [1, 1], [2, 2]], dtype=np.float64), [2, 2], 1.0, [6, 6], [0.25, 0.25])
    with pytest.raises(SmgError, match="v\\[j\\] must be > 0"):
        Chain(X, [2, 2], 1.0, [0.0, 6], [0.25, 0.25])
    Chain(X, [2, 2], 1.0, [0.5, 6], [0.25, 0.25]).close()  # v_j <= 1 is accepted, as in the reference (hyperg.cpp:359-376)
    with pytest.raises(SmgError, match="State validation failed"):
        Chain(X, [2, 2], 1.0, [6, 6], [0.25, 0.25], c_i=[0, 0, 2, 2])
    ch = Chain(X, [2, 2], 1.0, [6, 6], [0.25, 0.25], c_i=[5, 5, 6, 6])  # any base, shifted by min (launcher.cpp:34-37)
    assert ch.snapshot()["K"] == 2
    ch.close()


def test_run_markov_chain_equals_stepwise_chain_on_a_large_matrix():
    """>= 2^20 entries: the fp64 matrix is packed by host threads (not by the device ingest kernel) and the
    snapshots leave through the asynchronous pinned ring; every kept iteration must equal what the step-wise
    handle (uint8 upload, synchronous snapshot) produces from the same seed."""
    from split_and_merge_gibbs_sampling_b200 import run_markov_chain
    pb = Problem(4200, 256, 5, 8, seed=21)
    kw = dict(m=3, L=8, t=3, r=3, neal8=True, split_merge=True)
    res = run_markov_chain(np.asfortranarray(pb.X.astype(np.float64)), pb.attr, pb.gamma, pb.v, pb.w, iterations=7,
                           burnin=2, c_i=pb.labels, seed=33, thinning=2, **kw)
    from split_and_merge_gibbs_sampling_b200 import Chain
    ch = Chain(pb.X, pb.attr, pb.gamma, pb.v, pb.w, c_i=pb.labels, seed=33, thinning=2, data_u8=True, **kw)
    ch.step(2 * 2)  # burn-in * thinning
    for it in range(7):
        ch.step(1)
        s = ch.snapshot()
        assert res["total_cls"][it] == s["K"]
        assert np.array_equal(res["c_i"][it], s["c_i"])
        assert res["loglikelihood"][it] == s["loglikelihood"]
        assert np.array_equal(np.array(res["centers"][it]), s["centers"])
        assert np.array_equal(np.array(res["sigmas"][it]), s["sigmas"])
        assert res["accepted"][it] == s["accepted"]
        ch.step(1)
    ch.close()
    bad = np.asfortranarray(pb.X.astype(np.float64))
    bad[17, 3] = 9.0  # beyond attrisize
    from split_and_merge_gibbs_sampling_b200 import SmgError
    with pytest.raises(SmgError):
        run_markov_chain(bad, pb.attr, pb.gamma, pb.v, pb.w, iterations=1, burnin=0, c_i=pb.labels, seed=1, **kw)


def test_posterior_summaries_match_oracle_at_the_metric_shape():
    """north_star: 'matching posterior summaries' -- K-posterior, ARI, log-likelihood and the PSM of whole chains
    (launcher.cpp:85-154: Neal-8 + update_phi + split-merge every iteration) against the counted oracle at the metric's
    attribute shape (p=256, 5 levels, 3 auxiliaries) over 5 seeds, from L=20 random labels."""
    from sklearn.metrics import adjusted_rand_score
    from split_and_merge_gibbs_sampling_b200 import run_markov_chain
    n, p, kt = 3000, 256, 12
    pb = Problem(n, p, 5, kt, seed=77, s=1.3)
    Xd = pb.X.astype(np.float64)
    o = orc.opts(counted=1, stable_hig=1)
    its, burn = 30, 30
    sub = np.arange(0, n, 7)  # PSM on a row subsample: 429 x 429
    G = {"K": [], "ll": [], "ari": []}
    Cc = {"K": [], "ll": [], "ari": []}
    psm_g = np.zeros((sub.size, sub.size))
    psm_c = np.zeros((sub.size, sub.size))
    for seed in range(1, 6):
        res = run_markov_chain(Xd, pb.attr, pb.gamma, pb.v, pb.w, m=3, iterations=its, L=20, burnin=burn, t=5, r=5,
                               neal8=True, split_merge=True, seed=seed)
        G["K"].append(np.mean(res["total_cls"]))
        G["ll"].append(np.mean(res["loglikelihood"]))
        G["ari"].append(adjusted_rand_score(pb.labels, res["final_ass"]))
        cg = np.asarray(res["c_i"])[:, sub]
        psm_g += (cg[:, :, None] == cg[:, None, :]).mean(0)
        r = orc.run_chain(pb.od, 3, its, 20, None, burn, 5, 5, True, True, seed=seed, o=o)
        Cc["K"].append(np.mean(r["total_cls"]))
        Cc["ll"].append(np.mean(r["loglikelihood"]))
        Cc["ari"].append(adjusted_rand_score(pb.labels, r["final_ass"]))
        cc = np.asarray(r["c_i"])[:, sub]
        psm_c += (cc[:, :, None] == cc[:, None, :]).mean(0)

    def close(a, b, floor):
        se = np.sqrt(np.var(a, ddof=1) / len(a) + np.var(b, ddof=1) / len(b))
        return abs(np.mean(a) - np.mean(b)) <= 3 * se + floor
    assert close(G["K"], Cc["K"], 0.5), (G["K"], Cc["K"])
    assert close(G["ll"], Cc["ll"], 1e-3 * abs(np.mean(Cc["ll"]))), (G["ll"], Cc["ll"])
    assert close(G["ari"], Cc["ari"], 0.02), (G["ari"], Cc["ari"])
    assert np.mean(G["ari"]) > 0.9 and np.mean(Cc["ari"]) > 0.9
    # PSM averaged over the seeds: mean absolute difference of the co-clustering probabilities
    assert np.mean(np.abs(psm_g - psm_c)) / 5 < 0.02
