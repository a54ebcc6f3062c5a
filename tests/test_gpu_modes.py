"""Driver options of run_markov_chain (code/launcher.cpp:85-154): n8_step_size, sam_step_size, neal8 / split_merge
switches, capacity errors.  In every mode the snapshot must satisfy the reference's validate_state invariant and its
log-likelihood must be reproducible by the oracle from the snapshot itself (1e-12 relative)."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem

pytestmark = pytest.mark.gpu


def _check_snapshot(pb, s):
    K, c = s["K"], s["c_i"]
    assert c.min() == 0 and c.max() == K - 1 and len(np.unique(c)) == K
    ll = orc.loglik(pb.od, c, s["centers"], s["sigmas"])
    assert abs(ll - s["loglikelihood"]) <= 1e-12 * abs(ll)


@pytest.mark.parametrize("kw", [dict(n8_step_size=2, sam_step_size=3), dict(n8_step_size=3, sam_step_size=1),
                                dict(neal8=False, split_merge=True), dict(neal8=True, split_merge=False),
                                dict(t=0, r=0), dict(t=4, r=1), dict(t=1, r=5)])
def test_step_sizes_and_switches(kw):
    pb = Problem(1800, 40, 4, 5, seed=71, s=0.9)
    ch = pb.chain(L=7, c_i=None, compact_init=True, seed=72, **kw)
    for _ in range(9):
        ch.step(1)
        _check_snapshot(pb, ch.snapshot())
    ch.close()


def test_step_many_iterations_equals_one_at_a_time():
    pb = Problem(1500, 32, 4, 5, seed=73)
    a = pb.chain(L=6, c_i=None, compact_init=True, seed=74, n8_step_size=2)
    b = pb.chain(L=6, c_i=None, compact_init=True, seed=74, n8_step_size=2)
    a.step(7)
    for _ in range(7):
        b.step(1)
    sa, sb = a.snapshot(), b.snapshot()
    assert sa["K"] == sb["K"] and np.array_equal(sa["c_i"], sb["c_i"]) and sa["loglikelihood"] == sb["loglikelihood"]
    assert np.array_equal(sa["sigmas"], sb["sigmas"])
    a.close()
    b.close()


def test_cluster_capacity_error_is_reported():
    from split_and_merge_gibbs_sampling_b200 import SmgError
    # every observation its own cluster, capacity far below n
    pb = Problem(64, 8, 3, 3, seed=75, s=1.5)
    with pytest.raises(SmgError):
        pb.chain(c_i=np.arange(64, dtype=np.int32), max_clusters=16)


def test_long_run_crosses_the_pool_refresh():
    """1100 iterations: the auxiliary pool is re-drawn at iteration 1000 (launcher.cpp:123-129).  The chain must stay
    valid, and a second chain from the same seed stepped in different call sizes must agree bit for bit (the aux
    columns of the next pass are prefetched on a side stream; the refresh must not race them)."""
    pb = Problem(700, 24, 4, 4, seed=81, s=0.9)
    a = pb.chain(L=5, c_i=None, compact_init=True, seed=82)
    b = pb.chain(L=5, c_i=None, compact_init=True, seed=82)
    a.step(1100)
    for k in (999, 1, 1, 99):
        b.step(k)
    sa, sb = a.snapshot(), b.snapshot()
    _check_snapshot(pb, sa)
    assert sa["K"] == sb["K"] and np.array_equal(sa["c_i"], sb["c_i"]) and sa["loglikelihood"] == sb["loglikelihood"]
    assert np.array_equal(sa["sigmas"], sb["sigmas"])
    a.close()
    b.close()


def test_incremental_histogram_equals_recount():
    """The cluster histograms behind update_phi are maintained incrementally (only rows whose label changed since
    the last sweep are moved).  After sweeps with births, deaths, relabelling and accepted split-merge proposals
    the table must equal the oracle's recount of the current labels (integers: bit-exact)."""
    pb = Problem(900, 40, 5, 6, seed=83, s=0.8)
    ch = pb.chain(L=12, c_i=None, compact_init=True, seed=84)
    seen_K = set()
    for k in (1, 1, 3, 20, 75):
        ch.step(k)
        s = ch.snapshot(with_phi=False)
        seen_K.add(int(s["K"]))
        H, cnt = ch.histogram(s["K"])
        Ho, cnto = orc.histogram(pb.od, s["K"], s["c_i"].astype(np.int32), H.shape[2])
        assert np.array_equal(H, Ho) and np.array_equal(cnt, cnto)
        ch.step(1)  # the sweep after an on-demand histogram call starts from an up-to-date table
    assert len(seen_K) > 1  # the run did change the number of clusters
    ch.close()


def test_checkpoint_resume_is_bit_exact():
    """A chain resumed from (iteration, snapshot) on a fresh handle continues exactly like the uninterrupted one,
    also across the pool refresh at iteration 1000 (the reference has no checkpointing)."""
    pb = Problem(600, 24, 4, 4, seed=91, s=0.9)
    kw = dict(L=5, c_i=None, compact_init=True, seed=92)
    a = pb.chain(**kw)
    a.step(997)
    ck = a.checkpoint()
    assert ck["iteration"] == 997
    a.step(40)
    b = pb.chain(**kw)
    b.resume(ck)
    b.step(40)
    sa, sb = a.snapshot(), b.snapshot()
    assert sa["K"] == sb["K"] and np.array_equal(sa["c_i"], sb["c_i"]) and sa["loglikelihood"] == sb["loglikelihood"]
    assert np.array_equal(sa["centers"], sb["centers"]) and np.array_equal(sa["sigmas"], sb["sigmas"])
    a.close()
    b.close()


def test_validate_state_and_device_generator():
    from split_and_merge_gibbs_sampling_b200 import Chain, synth_generate
    n, p, m, kt, s = 6000, 48, 5, 6, 0.5
    X, lab, cen, attr = synth_generate(n, p, m, kt, s=s, seed=5)
    assert X.min() >= 1 and X.max() <= m and np.array_equal(np.bincount(lab), np.full(kt, n // kt))
    match = (X == cen[lab]).mean()
    pm = 1.0 / (1.0 + (m - 1) * np.exp(-1.0 / s))
    assert abs(match - pm) < 4 * np.sqrt(pm * (1 - pm) / (n * p))  # match rate of the spec (data_generation.R)
    other = X[X != cen[lab]]
    assert other.size > 0 and len(np.unique(other)) == m
    ch = Chain(X, attr, 1.0, np.full(p, 6.0), np.full(p, 0.25), m=3, L=kt, seed=3, compact_init=True, data_u8=True)
    for _ in range(6):
        ch.step(1)
        ch.validate_state()  # common_functions.cpp:146-172 on the device state
    from sklearn.metrics import adjusted_rand_score
    assert adjusted_rand_score(lab, ch.snapshot(with_phi=False)["c_i"]) > 0.95
    ch.close()


def _run_trace(pb, mode, iters, overlap=True, **kw):
    """Steps a chain one iteration at a time under the given split-merge device path; returns the snapshots."""
    import os
    os.environ["SMG_SM_MODE"] = mode
    os.environ["SMG_NO_K1_OVERLAP"] = "0" if overlap else "1"
    try:
        ch = pb.chain(**kw)
        out = []
        for _ in range(iters):
            ch.step(1)
            s = ch.snapshot()
            _check_snapshot(pb, s)
            out.append(s)
        st = ch.stats()
        ch.close()
    finally:
        os.environ.pop("SMG_SM_MODE", None)
        os.environ.pop("SMG_NO_K1_OVERLAP", None)
    return out, st


def test_split_merge_device_paths_give_the_same_chain():
    """Cluster kernel, cooperative kernel and the sequence of launches consume the same Philox draws and take the same
    decisions: the chains coincide (allocations, K, acceptance flags, centres).  The start is over-merged (everything in
    one cluster) so that splits ARE accepted, and log-likelihoods are checked against the oracle after every iteration:
    the likelihood block evaluated beside the proposal must have been patched after each acceptance."""
    pb = Problem(2400, 32, 4, 6, seed=91, s=0.6)
    kw = dict(c_i=np.zeros(2400, dtype=np.int32), seed=92, t=4, r=3)
    ref, st_ref = _run_trace(pb, "multi", 40, overlap=False, **kw)
    assert st_ref["sm_accepted"] >= 2  # the patch path is exercised
    for mode, overlap in (("cluster", True), ("coop", True), ("multi", True), ("cluster", False)):
        got, st = _run_trace(pb, mode, 40, overlap=overlap, **kw)
        assert st["sm_accepted"] == st_ref["sm_accepted"], mode
        for a, b in zip(ref, got):
            assert a["K"] == b["K"] and a["accepted"] == b["accepted"], mode
            assert np.array_equal(a["c_i"], b["c_i"]), mode
            assert np.array_equal(a["centers"], b["centers"]), mode
            assert np.max(np.abs(a["sigmas"] - b["sigmas"]) / b["sigmas"]) < 1e-12, mode
            assert abs(a["loglikelihood"] - b["loglikelihood"]) <= 1e-12 * abs(b["loglikelihood"]), mode


def test_merge_acceptance_is_patched_too():
    """Over-split start (every true cluster cut in two at random): some merge gets accepted (seeds are tried until the
    reference path accepts one while K decreases); the overlapped likelihood block must have been patched -- same chain,
    log-likelihoods checked against the oracle after every iteration."""
    for seed in range(94, 104):
        pb = Problem(320, 16, 3, 3, seed=seed, s=0.4)
        rng = np.random.default_rng(seed)
        c0 = (pb.labels * 2 + rng.integers(0, 2, size=320)).astype(np.int32)
        kw = dict(c_i=c0, seed=seed, t=2, r=2)
        ref, st_ref = _run_trace(pb, "multi", 80, overlap=False, **kw)
        merged = any(b["accepted"] and b["K"] < a["K"] for a, b in zip(ref[:-1], ref[1:]))
        if not merged:
            continue
        got, st = _run_trace(pb, "cluster", 80, overlap=True, **kw)
        assert st["sm_accepted"] == st_ref["sm_accepted"]
        for a, b in zip(ref, got):
            assert a["K"] == b["K"] and a["accepted"] == b["accepted"]
            assert np.array_equal(a["c_i"], b["c_i"])
        return
    pytest.fail("no seed produced an accepted merge")


def test_pool_free_auxiliary_components():
    """aux_mode='philox' (north_star item 1: auxiliary columns from per-observation Philox streams): every auxiliary
    component is a fresh prior draw (launcher.cpp:67-77 law: centre ~ U{1..m_j}, sigma ~ HIG(v_j,w_j,m_j)) -- the column
    values are the Hamming log-likelihood of those parameters (1e-12), centres are uniform, sigmas follow the HIG prior
    (KS), entries differ between observations and between passes, and no n*m*p pool is allocated."""
    from scipy import special, stats
    pb = Problem(1500, 48, 4, 5, seed=91, s=0.9)
    ch = pb.chain(L=6, c_i=None, compact_init=True, seed=5, aux_mode="philox")
    cnt = 1500 * 3
    ll, cen, sig = ch.aux_free(cnt)
    # column value = -sum_j [x != c]/sigma - sum_j log(1 + (m-1) exp(-1/sigma))  of observation w // 3
    X = pb.X[np.arange(cnt) // 3].astype(np.float64)
    ref = -np.sum((X != cen) / sig, axis=1) - np.sum(np.log1p((pb.attr[None, :] - 1.0) * np.exp(-1.0 / sig)), axis=1)
    assert np.max(np.abs(ll - ref) / np.abs(ref)) < 1e-12
    assert cen.min() == 1 and cen.max() == 4
    freq = np.bincount(cen.astype(int).ravel(), minlength=5)[1:] / cen.size
    assert np.max(np.abs(freq - 0.25)) < 0.01
    v, w, m = 6.0, 0.25, 4.0
    u = np.exp(-1.0 / sig.ravel()[:20000])
    a, b, xmax = w + 1.0, v - 1.0, (m - 1.0) / m

    def cdf(uu):
        x = uu * (m - 1) / (1 + uu * (m - 1))
        return special.betainc(a, b, x) / special.betainc(a, b, xmax)
    assert stats.kstest(u, cdf).pvalue > 1e-3
    assert len(np.unique(sig[:, 0])) == cnt  # a fresh draw per (observation, component)
    ch.step(1)
    ll2, cen2, sig2 = ch.aux_free(cnt)
    assert not np.array_equal(sig2, sig)  # and per pass
    for _ in range(6):
        ch.step(1)
        _check_snapshot(pb, ch.snapshot())
    with pytest.raises(Exception):
        ch.set_pool(np.ones((4, pb.p)), np.ones((4, pb.p)))
    ch.close()
    # same seed, same chain
    a1 = pb.chain(L=6, c_i=None, compact_init=True, seed=5, aux_mode="philox")
    a2 = pb.chain(L=6, c_i=None, compact_init=True, seed=5, aux_mode="philox")
    a1.step(5)
    a2.step(5)
    s1, s2 = a1.snapshot(), a2.snapshot()
    assert np.array_equal(s1["c_i"], s2["c_i"]) and s1["loglikelihood"] == s2["loglikelihood"]
    a1.close()
    a2.close()


def test_pool_free_and_pool_modes_agree_in_distribution():
    """Same target distribution: K, log-likelihood and ARI of chains with pool-free auxiliary components against the
    stored-pool mode over 6 seeds (noisy data, so that births from auxiliary components do happen)."""
    from sklearn.metrics import adjusted_rand_score
    pb = Problem(1000, 32, 4, 4, seed=93, s=1.0)
    out = {"pool": {"K": [], "ll": [], "ari": [], "births": 0}, "philox": {"K": [], "ll": [], "ari": [], "births": 0}}
    for mode in out:
        for seed in range(1, 7):
            ch = pb.chain(L=8, c_i=None, compact_init=True, seed=seed, aux_mode=mode, t=3, r=3)
            ch.step(40)
            Ks, lls = [], []
            for _ in range(80):
                ch.step(1)
                s = ch.snapshot(with_phi=False)
                Ks.append(s["K"])
                lls.append(s["loglikelihood"])
            out[mode]["K"].append(np.mean(Ks))
            out[mode]["ll"].append(np.mean(lls))
            out[mode]["ari"].append(adjusted_rand_score(pb.labels, s["c_i"]))
            out[mode]["births"] += ch.stats()["births"]
            ch.close()

    def close(a, b, floor):
        se = np.sqrt(np.var(a, ddof=1) / len(a) + np.var(b, ddof=1) / len(b))
        return abs(np.mean(a) - np.mean(b)) <= 3 * se + floor
    assert close(out["pool"]["K"], out["philox"]["K"], 0.75), out
    assert close(out["pool"]["ll"], out["philox"]["ll"], 15.0), out
    assert close(out["pool"]["ari"], out["philox"]["ari"], 0.03), out


def test_pool_free_births_carry_the_parameters_their_column_was_evaluated_with():
    """A cluster opened from a pool-free auxiliary component must get exactly the (centre, sigma) its column value was
    computed from (they are re-derived from the Philox key inside the scan, never stored)."""
    pb = Problem(400, 8, 3, 3, seed=95, s=2.0)
    ch = pb.chain(L=2, c_i=None, compact_init=True, seed=9, aux_mode="philox", m=3)
    born = 0
    for it in range(12):
        before = ch.snapshot()
        ll, cen, sig = ch.aux_free(pb.n * 3)
        ch.neal8_scan(None)
        after = ch.snapshot()
        old_rows = {tuple(r) for r in before["sigmas"]}
        aux_rows = {tuple(r): k for k, r in enumerate(sig)}
        for k in range(after["K"]):
            key = tuple(after["sigmas"][k])
            if key in old_rows:
                continue
            assert key in aux_rows, "a new cluster's sigmas are not those of any auxiliary component of the pass"
            assert np.array_equal(after["centers"][k], cen[aux_rows[key]])
            born += 1
        ch.step(1)
    assert born >= 5
    ch.close()
