"""Host-side diagnostics: ESS/IAT, ARI, Binder point estimate from PSM counts."""
import numpy as np

from split_and_merge_gibbs_sampling_b200 import diagnostics as dg


def test_ari_matches_sklearn():
    from sklearn.metrics import adjusted_rand_score
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.integers(0, 5, 200)
        b = np.where(rng.random(200) < 0.7, a, rng.integers(0, 5, 200))
        assert abs(dg.adjusted_rand_index(a, b) - adjusted_rand_score(a, b)) < 1e-12
    assert dg.adjusted_rand_index(a, a) == 1.0


def test_iat_of_ar1():
    rng = np.random.default_rng(1)
    phi, n = 0.8, 200000
    x = np.zeros(n)
    e = rng.normal(size=n)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    true_iat = (1 + phi) / (1 - phi)  # 9
    assert abs(dg.iat(x) - true_iat) / true_iat < 0.1
    assert abs(dg.ess(x) - n / true_iat) / (n / true_iat) < 0.1
    assert abs(dg.iat(rng.normal(size=50000)) - 1.0) < 0.1


def test_binder_estimate_picks_the_consensus():
    truth = np.repeat(np.arange(3), 10)
    rng = np.random.default_rng(2)
    draws = []
    for _ in range(40):
        c = truth.copy()
        flip = rng.integers(0, 30, 2)
        c[flip] = rng.integers(0, 3, 2)
        draws.append(c)
    psm = sum((c[:, None] == c[None, :]).astype(np.int64) for c in draws)
    best, losses = dg.binder_point_estimate(psm, len(draws), draws + [truth])
    assert losses[-1] <= losses.min() + 1e-9 or dg.adjusted_rand_index(draws[best], truth) > 0.9
