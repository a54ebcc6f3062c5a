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
