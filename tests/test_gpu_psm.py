"""PSM accumulation on the tensor cores (tcgen05 u8 x u8 -> s32 over one-hot allocations) against numpy
one-hot Z Z^T: integer counts, bit-exact (SURVEY 8(d) parity gate)."""
import numpy as np
import pytest

from helpers import Problem

pytestmark = pytest.mark.gpu


def numpy_psm(labels):
    n = labels.shape[1]
    out = np.zeros((n, n), dtype=np.int64)
    for c in labels:
        out += (c[:, None] == c[None, :])
    return out.astype(np.int32)


@pytest.mark.parametrize("n,T,K,cap", [(300, 7, 9, 64), (1000, 70, 50, 32), (515, 5, 150, 8), (129, 3, 100, 2), (64, 1, 1, 4)])
def test_psm_matches_numpy(n, T, K, cap):
    from split_and_merge_gibbs_sampling_b200 import Psm
    rng = np.random.default_rng(n + T)
    labels = rng.integers(0, K, size=(T, n)).astype(np.int32)
    P = Psm(n, capacity_sweeps=cap)
    for c in labels:
        P.push(c)
    got = P.read()
    assert np.array_equal(got, numpy_psm(labels))
    assert P.info()["sweeps"] == T
    P.close()


def test_psm_device_reference_agrees():
    from split_and_merge_gibbs_sampling_b200 import Psm
    rng = np.random.default_rng(5)
    n, T = 700, 11
    labels = rng.integers(0, 40, size=(T, n)).astype(np.int32)
    P = Psm(n, capacity_sweeps=16)
    for c in labels:
        P.push(c)
    ref = P.reference()
    assert np.array_equal(P.read(), ref)
    P.close()


def test_psm_from_chain_allocations():
    from split_and_merge_gibbs_sampling_b200 import Psm
    pb = Problem(900, 32, 4, 6, seed=3)
    ch = pb.chain(L=6, c_i=None, compact_init=True, seed=4)
    P = Psm(pb.n, capacity_sweeps=4)
    kept = []
    for _ in range(9):
        ch.step(1)
        P.push_chain(ch)
        kept.append(ch.snapshot(with_phi=False)["c_i"].copy())
    got = P.read()
    assert np.array_equal(got, numpy_psm(np.stack(kept)))
    assert np.all(np.diag(got) == 9)
    ch.close()
    P.close()


def test_point_estimate_ari_and_ess_on_device_match_host():
    """Binder / VI point estimate from the device PSM, ARI and IAT/ESS kernels against the host numpy
    diagnostics (zoo_simulator.R:205-215,339-344)."""
    from split_and_merge_gibbs_sampling_b200 import Comm, Psm, adjusted_rand_index, trace_ess
    from split_and_merge_gibbs_sampling_b200 import diagnostics as dg
    rng = np.random.default_rng(12)
    n, T = 700, 40
    base = rng.integers(0, 6, n)
    draws = np.empty((T, n), dtype=np.int32)
    for t in range(T):
        c = base.copy()
        flip = rng.random(n) < 0.15
        c[flip] = rng.integers(0, 8, flip.sum())
        draws[t] = c
    P = Psm(n, capacity_sweeps=16)
    for t in range(T):
        P.push(draws[t])
    M = P.read().astype(np.int64)
    cand = draws[:11]
    b, v = P.point_estimate(cand, T)
    same = (cand[:, :, None] == cand[:, None, :])
    ref_b = np.array([np.triu(np.abs(T * s.astype(np.int64) - M), 1).sum() for s in same])
    assert np.array_equal(b, ref_b)  # exact integers
    idx, losses = dg.binder_point_estimate(M, T, cand)
    assert idx == int(np.argmin(b)) and np.allclose(losses, b / T, rtol=1e-12)
    a_i = same.sum(2)
    b_i = (same * M[None]).sum(2)
    s_i = M.sum(1)[None]
    ref_v = (np.log2(a_i) + np.log2(s_i / T) - 2 * np.log2(b_i / T)).sum(1)
    assert np.max(np.abs(v - ref_v) / np.abs(ref_v)) < 1e-12
    # row blocks add up (what the reduce-scattered matrix needs)
    b1, v1 = P.point_estimate(cand, T, 0, 300)
    b2, v2 = P.point_estimate(cand, T, 300, 400)
    assert np.array_equal(b1 + b2, b) and np.allclose(v1 + v2, v, rtol=1e-12)
    res = Comm().point_estimate(P, cand, T)
    assert res["best_binder"] == int(np.argmin(b)) and res["best_vi"] == int(np.argmin(v))
    assert np.allclose(res["vi_lower_bound"], v / n, rtol=1e-12)
    P.close()
    for _ in range(5):
        a, bb = rng.integers(0, 9, 5000), rng.integers(0, 7, 5000)
        bb[:2500] = a[:2500] % 7
        assert abs(adjusted_rand_index(a, bb) - dg.adjusted_rand_index(a, bb)) < 1e-12
    assert adjusted_rand_index(base, base) == 1.0
    x = np.cumsum(rng.normal(size=(3, 2000)), axis=1) * 0.05 + rng.normal(size=(3, 2000))
    iat, ess = trace_ess(x)
    for r in range(3):
        assert abs(iat[r] - dg.iat(x[r])) <= 1e-9 * dg.iat(x[r])
        assert abs(ess[r] - dg.ess(x[r])) <= 1e-9 * dg.ess(x[r])


@pytest.mark.parametrize("n,T,K,cap", [(512, 9, 12, 4), (1000, 70, 50, 32), (320, 5, 150, 8)])
def test_psm_distributed_form_on_one_rank(n, T, K, cap):
    """smg_chains_psm_distribute with a one-rank communicator: the epilogue that adds tiles into the owner's memory
    (atomics through the peer table) and the distributed mirror give the same counts as numpy."""
    from split_and_merge_gibbs_sampling_b200 import Comm, Psm
    rng = np.random.default_rng(n + T)
    labels = rng.integers(0, K, size=(T, n)).astype(np.int32)
    C, P = Comm(), Psm(n, capacity_sweeps=cap)
    C.distribute_psm(P)
    for c in labels:
        P.push(c)
    r0, nr, ms, bus = C.reduce_psm(P, "reduce_scatter")
    assert (r0, nr) == (0, n)
    assert np.array_equal(P.read(), numpy_psm(labels))
    P.close()
    C.close()
