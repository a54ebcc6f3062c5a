"""Device special functions against the golden values and the oracle (fp64, tolerances stated)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def test_hig_inverse_cdf_golden():
    from split_and_merge_gibbs_sampling_b200 import hig_inv_u
    g = G["hig_inv_u"]
    u = hig_inv_u([r["omega"] for r in g], [r["v"] for r in g], [r["w"] for r in g], [float(r["m"]) for r in g])
    ref = np.array([r["u"] for r in g])
    assert np.max(np.abs(u - ref)) < 1e-13


def test_hig_inverse_cdf_vs_reference_bisection():
    # hyperg.cpp:221-287 stops at a bracket of 1e-9: the device root must lie within that width
    from split_and_merge_gibbs_sampling_b200 import hig_inv_u
    rng = np.random.default_rng(3)
    L = orc.lib()
    cases = [(6, 0.25, 2), (6, 0.25, 5), (3, 0.5, 6), (8, 18.25, 2), (40, 25.25, 4), (206, 120.25, 5), (1.5, 0.25, 3)]
    for (v, w, m) in cases:
        om = rng.random(64)
        om[:4] = [1e-6, 1e-3, 0.999, 0.5]
        got = hig_inv_u(om, v, w, float(m))
        ref = np.empty_like(om)
        o = orc.opts(stable_hig=1, sigma_inverse_cdf=1)
        assert L.orc_rhig_u_from_omega(v, w, m, om.size, orc.P(om), C.byref(o), orc.P(ref)) == 0
        assert np.max(np.abs(got - ref)) <= 1e-9
        o2 = orc.opts(stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0)
        assert L.orc_rhig_u_from_omega(v, w, m, om.size, orc.P(om), C.byref(o2), orc.P(ref)) == 0
        assert np.max(np.abs(got - ref) / np.maximum(ref, 1e-300)) <= 1e-10


def test_hig_large_cluster_parameters():
    from split_and_merge_gibbs_sampling_b200 import hig_inv_u
    L = orc.lib()
    rng = np.random.default_rng(5)
    for (v, w, m) in [(1306, 700.25, 5), (6506, 3500.25, 5), (65006, 35000.25, 5), (30006, 20000.25, 2)]:
        om = rng.random(16)
        got = hig_inv_u(om, v, w, float(m))
        ref = np.empty_like(om)
        o2 = orc.opts(stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0)
        assert L.orc_rhig_u_from_omega(v, w, m, om.size, orc.P(om), C.byref(o2), orc.P(ref)) == 0
        assert np.all(np.isfinite(got))
        assert np.max(np.abs(got - ref) / ref) <= 1e-9


def test_logdensity_hig_golden_and_oracle():
    from split_and_merge_gibbs_sampling_b200 import logdensity_hig
    g = G["logdensity_hig"]
    got = logdensity_hig([r["s"] for r in g], [r["v"] for r in g], [r["w"] for r in g], [float(r["m"]) for r in g])
    ref = np.array([r["val"] for r in g])
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-12
    rng = np.random.default_rng(9)
    L = orc.lib()
    s = rng.uniform(0.1, 3.0, 200)
    v = rng.uniform(1.5, 3000, 200)
    w = rng.uniform(0.0, 2000, 200)
    m = rng.integers(2, 7, 200).astype(np.float64)
    got = logdensity_hig(s, v, w, m)
    ref = np.array([L.orc_logdensity_hig(a, b, c, d, 1, None) for a, b, c, d in zip(s, v, w, m)])
    # the density is a difference of lgamma-sized addends ((v+w) log(v+w)): 1e-12 relative to the largest addend
    scale = np.maximum(np.maximum(1.0, np.abs(ref)), (v + w) * np.log(v + w))
    assert np.max(np.abs(got - ref) / scale) < 1e-12


def test_production_sigma_sampler_has_the_hig_law():
    """The Beta-rejection sampler (hyperg.cpp:359-368 branch) is checked in distribution:
    KS against the exact CDF in u-space, I_x(w+1,v-1)/I_xmax(w+1,v-1), x = u(m-1)/(1+u(m-1))."""
    from scipy import special, stats
    from split_and_merge_gibbs_sampling_b200 import rhig_u
    cases = [(6, 0.25, 2), (6, 0.25, 5), (3, 0.5, 6), (1.5, 0.25, 3), (8, 18.25, 2), (10, 40.25, 2), (1306, 700.25, 5),
             (40, 25.25, 4)]
    for (v, w, m) in cases:
        u = rhig_u(20000, v, w, m, seed=3)
        assert np.all((u > 0) & (u < 1))
        a, b, xmax = w + 1.0, v - 1.0, (m - 1.0) / m

        def cdf(uu):
            x = uu * (m - 1) / (1 + uu * (m - 1))
            return special.betainc(a, b, x) / special.betainc(a, b, xmax)
        assert stats.kstest(u, cdf).pvalue > 1e-3, (v, w, m)
    # draws of different seeds / sites are different streams
    assert not np.array_equal(rhig_u(100, 6, 0.25, 2, seed=1), rhig_u(100, 6, 0.25, 2, seed=2))


def test_hig_with_v_at_most_one():
    """v_j <= 1: the reference's Beta(w+1, v-1) proposal does not exist (qbeta -> NaN, hyperg.cpp:359), every draw takes
    the inverse-CDF branch (:370-376).  Device inverse CDF against 40-digit quadrature of the CDF, against the oracle's
    bisection in the reference's own 2F1 form (stable_hig=0) and in the continued-fraction form; log-density against both;
    the production sampler in distribution."""
    import mpmath as mp
    from scipy import stats
    from split_and_merge_gibbs_sampling_b200 import hig_inv_u, logdensity_hig, rhig_u
    mp.mp.dps = 40
    L = orc.lib()
    rng = np.random.default_rng(17)
    for (v, w, m) in [(1.0, 0.25, 5), (0.8, 0.25, 5), (0.3, 0.0, 2), (0.5, 2.25, 4), (1.0, 3.0, 7)]:
        om = rng.random(24)
        om[:3] = [1e-6, 0.5, 0.999]
        got = hig_inv_u(om, v, w, float(m))
        assert np.all((got > 0) & (got < 1))
        f = lambda t: t ** w * (1 + (m - 1) * t) ** (-(v + w))  # density in u = exp(-1/sigma), hyperg.cpp:111
        tot = mp.quad(f, [0, 1])
        cdf = np.array([float(mp.quad(f, [0, float(g)]) / tot) for g in got])
        assert np.max(np.abs(cdf - om)) < 1e-12, (v, w, m)
        for stable in (0, 1):
            ref = np.empty_like(om)
            o = orc.opts(stable_hig=stable, sigma_inverse_cdf=1)
            assert L.orc_rhig_u_from_omega(v, w, m, om.size, orc.P(om), C.byref(o), orc.P(ref)) == 0
            assert np.max(np.abs(got - ref)) <= 1e-9, (v, w, m, stable)
        s = rng.uniform(0.1, 3.0, 16)
        gd = logdensity_hig(s, np.full(16, v), np.full(16, w), np.full(16, float(m)))
        for stable in (0, 1):
            ref = np.array([L.orc_logdensity_hig(a, v, w, float(m), stable, None) for a in s])
            assert np.max(np.abs(gd - ref) / np.maximum(1.0, np.abs(ref))) < 1e-12, (v, w, m, stable)
        u = rhig_u(8000, v, w, m, seed=5)
        grid = np.linspace(0.0, 1.0, 2001)
        cg = np.array([float(mp.quad(f, [0, float(g)]) / tot) if g > 0 else 0.0 for g in grid[::20]])
        assert stats.kstest(u, lambda x: np.interp(x, grid[::20], cg)).pvalue > 1e-3, (v, w, m)


def test_chain_with_v_at_most_one_runs_and_matches_the_oracle_update():
    """Whole iterations with v_j = 0.8 (prior draws by inversion only), and update_phi under an injected tape against the
    oracle (centres bit-exact, sigma within the reference's bisection bracket)."""
    from helpers import Problem
    pb = Problem(500, 12, 4, 3, seed=55, s=0.8, v=0.8, w=0.25)
    ch = pb.chain(L=4, c_i=None, compact_init=True, seed=3)
    for _ in range(8):
        ch.step(1)
        s = ch.snapshot()
        ch.validate_state()
        ll = orc.loglik(pb.od, s["c_i"], s["centers"], s["sigmas"])
        assert abs(ll - s["loglikelihood"]) <= 1e-12 * abs(ll)
    K, c = s["K"], s["c_i"]
    cen, sig = s["centers"], s["sigmas"]
    rng = np.random.default_rng(5)
    uc, us = rng.random((K, pb.p)), rng.random((K, pb.p))
    tape = np.concatenate([np.concatenate([uc[k], us[k]]) for k in range(K)])
    ref = orc.update_phi(pb.od, c, cen, sig, tape, o=orc.opts(stable_hig=1, sigma_inverse_cdf=1))
    ch.set_state(K, c, cen, sig)
    ch.update_phi(uc, us)
    got = ch.snapshot()
    assert np.array_equal(got["centers"], ref["center"])
    assert np.max(np.abs(np.exp(-1.0 / got["sigmas"]) - np.exp(-1.0 / ref["sigma"]))) <= 1e-9
    ch.close()
