"""BASELINE.json configs 3 and 4 at FULL size (C3: n = 6e4, p = 784, 2 levels -- 13 attribute tiles with a partial last
one; C4: n = 1e6, p = 256, 5 levels, K ~ 100 -- the wide evaluation path and 64-bit index arithmetic at scale).  The
oracle cannot replay an allocation pass at these sizes in seconds, so the gates are size-independent properties: integer
mismatch counts and the likelihood block of the current state (O(n K p), exact / 1e-12), the reference's validate_state
invariant after every iteration, the log-likelihood the device reports against the oracle's for the snapshot it returns
(1e-12), and the truth being kept (ARI)."""
import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu


def _run(n, p, cats, kt, s, gamma, v, w, sweeps, check_block):
    from sklearn.metrics import adjusted_rand_score
    from split_and_merge_gibbs_sampling_b200 import Chain
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(n, p, cats, kt, s=s, seed=1)
    vv, ww = np.full(p, v), np.full(p, w)
    od = orc.OracleData(X, attr, gamma, vv, ww)
    ch = Chain(X, attr, gamma, vv, ww, m=3, L=kt, t=10, r=10, neal8=True, split_merge=True, seed=1, c_i=labels, data_u8=True,
               max_clusters=min(250, kt + 60))
    for _ in range(sweeps):
        ch.step(1)
        ch.validate_state()
    sn = ch.snapshot()
    # reference log-likelihood: per-observation values (common_functions.cpp:355-401 in the compact form) added with
    # math.fsum -- at n = 1e6 a plain running sum (the oracle's, the reference's) carries ~1e-11 of rounding itself, so
    # the 1e-12 gate is held against the exactly rounded sum and the oracle's own value is checked at 1e-10
    import math
    cen, sig, c = sn["centers"], sn["sigmas"], sn["c_i"]
    isg = 1.0 / sig
    sden = np.log1p((attr[None, :] - 1.0) * np.exp(-1.0 / sig)).sum(1)
    parts = []
    for a in range(0, n, 100000):
        cc = c[a:a + 100000]
        mism = X[a:a + 100000] != cen[cc].astype(np.uint8)
        parts.append(-(mism * isg[cc]).sum(1) - sden[cc])
    ll = math.fsum(np.concatenate(parts).tolist())
    assert abs(ll - sn["loglikelihood"]) <= 1e-12 * abs(ll)
    assert abs(orc.loglik(od, c, cen, sig) - ll) <= 1e-10 * abs(ll)
    assert abs(sn["K"] - kt) <= 3
    assert adjusted_rand_score(labels, sn["c_i"]) > 0.99
    if check_block:
        LL, mm = ch.ll_block(sn["K"])
        LLo, mmo = orc.ll_block(od, sn["centers"], sn["sigmas"])
        assert np.array_equal(mm, mmo)
        assert np.max(np.abs(LL - LLo) / np.abs(LLo)) < 1e-12
    st = ch.stats()
    ch.close()
    return st


def test_config_c3_full_size():
    st = _run(60000, 784, 2, 20, 0.6, 0.1514657, 3.0, 0.5, sweeps=3, check_block=True)
    assert st["sweeps"] == 3 and st["sm_proposals"] == 3


def test_config_c4_full_size():
    st = _run(1000000, 256, 5, 100, 0.5, 1.0, 6.0, 0.25, sweeps=2, check_block=False)
    assert st["sweeps"] == 2 and st["sm_proposals"] == 2
