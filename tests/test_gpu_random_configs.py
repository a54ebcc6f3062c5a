"""Randomised small configurations: the Neal-8 pass under an injected tape must reproduce the oracle's allocations
bit for bit whatever the shape (tiny n, one attribute, many levels, many aux components, K from 1 to n)."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem, oracle_state_full

pytestmark = pytest.mark.gpu


def _one(case_seed):
    rng = np.random.default_rng(case_seed)
    n = int(rng.choice([2, 3, 5, 17, 64, 130, 333]))
    p = int(rng.choice([1, 2, 7, 16, 33]))
    m = int(rng.choice([2, 3, 5, 8, 9, 20]))
    k = int(min(n, rng.choice([1, 2, 4, 9])))
    m_aux = int(rng.choice([1, 2, 3, 6]))
    s = float(rng.choice([0.4, 1.0, 3.0]))
    pb = Problem(n, p, m, k, seed=1000 + case_seed, s=s)
    mode = "truth" if rng.random() < 0.5 else "random"
    L = int(min(n, rng.integers(1, 8)))
    try:
        K, c, cen, sig = oracle_state_full(pb, mode=mode, seed=case_seed, L=L, iters=1, m_aux=m_aux)
    except orc.OracleError:
        return None  # e.g. an empty initial label: the reference stops too
    pool = int(rng.integers(1, 40))
    pc, ps = orc.draw_pool(pb.od, pool, case_seed + 1, o=orc.opts(stable_hig=1))
    tape = (rng.integers(0, 2**53, size=n * (m_aux + 1)).astype(np.float64) + 0.5) / 2.0**53
    ref = orc.neal8_scan(pb.od, m_aux, c, cen, sig, pc, ps, tape, o=orc.opts(counted=1), kcap=512)
    ch = pb.chain(m=m_aux, c_i=c, max_clusters=min(250, max(16, n + 4)), pool_size=max(pool, n * m_aux))
    ch.set_state(K, c, cen, sig)
    ch.set_pool(pc, ps)
    ch.neal8_scan(tape)
    got = ch.snapshot()
    ch.validate_state()
    ch.close()
    desc = dict(n=n, p=p, m=m, k=k, m_aux=m_aux, s=s, mode=mode, L=L, pool=pool)
    assert got["K"] == ref["K"], desc
    assert np.array_equal(got["c_i"], ref["c"]), desc
    assert np.array_equal(got["centers"], ref["center"]), desc
    assert np.array_equal(got["sigmas"], ref["sigma"]), desc
    return desc


def test_random_small_configurations():
    done = 0
    for case_seed in range(60):
        if _one(case_seed) is not None:
            done += 1
    assert done >= 40


def test_random_configurations_full_iterations_stay_valid():
    """Whole iterations (pass + update_phi + split-merge + log-likelihood) on odd shapes: the device state must pass
    validate_state after every iteration and the reported log-likelihood must be the oracle's for that state."""
    ran = 0
    for case_seed in range(40):
        rng = np.random.default_rng(5000 + case_seed)
        n = int(rng.choice([2, 3, 6, 25, 90, 400, 1100]))
        p = int(rng.choice([1, 3, 16, 20, 48]))
        m = int(rng.choice([2, 4, 7, 8, 13]))
        k = int(min(n, rng.choice([1, 2, 5])))
        m_aux = int(rng.choice([1, 3, 5]))
        pb = Problem(n, p, m, k, seed=2000 + case_seed, s=float(rng.choice([0.5, 1.5])))
        kw = dict(m=m_aux, L=int(min(n, rng.integers(1, 6))), c_i=None, compact_init=True, seed=case_seed,
                  t=int(rng.integers(0, 4)), r=int(rng.integers(0, 4)), max_clusters=min(250, max(16, n + 4)))
        ch = pb.chain(**kw)
        for _ in range(5):
            ch.step(1)
            ch.validate_state()
            s = ch.snapshot()
            ll = orc.loglik(pb.od, s["c_i"], s["centers"], s["sigmas"])
            assert abs(ll - s["loglikelihood"]) <= 1e-12 * max(abs(ll), 1.0), (kw, n, p, m)
        ch.close()
        ran += 1
    assert ran == 40
