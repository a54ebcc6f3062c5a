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
