"""Synthetic Hamming-mixture data (spec: code/old_code/data_generation.R:1-101, `ham_mix_gen`)
and the zoo fixture preparation (realdata_analysis/zoo_simulator.R:18-37).

True centres c_kj ~ U{1..m_j}; P(x_ij = c_kj) = 1/(1+(m_j-1) exp(-1/s_kj)), the other levels
uniform.  A counter-based numpy generator (Philox, seed stated) replaces R's set.seed(10091995).
"""
import numpy as np


def ham_mix_gen(n, p, m, k_true, s=0.5, seed=1, sizes=None):
    """Returns (X uint8 n x p with codes 1..m_j, labels int32 n, centres uint8 k x p, attrisize int32 p).

    `m` may be an int (every attribute has m levels) or a length-p sequence; `sizes` are the
    component sizes (default: as equal as possible); rows are shuffled."""
    rng = np.random.Generator(np.random.Philox(seed))
    attr = np.full(p, m, dtype=np.int32) if np.isscalar(m) else np.asarray(m, dtype=np.int32)
    if sizes is None:
        sizes = np.full(k_true, n // k_true)
        sizes[: n - sizes.sum()] += 1
    sizes = np.asarray(sizes)
    assert sizes.sum() == n
    cent = (rng.integers(0, 1 << 30, size=(k_true, p)) % attr[None, :] + 1).astype(np.uint8)
    labels = np.repeat(np.arange(k_true, dtype=np.int32), sizes)
    rng.shuffle(labels)
    sk = np.broadcast_to(np.asarray(s, dtype=np.float64), (k_true, p))
    pmatch = 1.0 / (1.0 + (attr[None, :] - 1) * np.exp(-1.0 / sk))
    X = np.empty((n, p), dtype=np.uint8)
    step = max(1, (1 << 24) // max(p, 1))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        lab = labels[lo:hi]
        c = cent[lab]
        u = rng.random((hi - lo, p))
        keep = u < pmatch[lab]
        # a uniformly chosen *other* level: shift by 1..m_j-1 (mod m_j)
        shift = (rng.integers(0, 1 << 30, size=(hi - lo, p)) % (attr[None, :] - 1)) + 1
        other = ((c.astype(np.int64) - 1 + shift) % attr[None, :] + 1).astype(np.uint8)
        X[lo:hi] = np.where(keep, c, other)
    return X, labels, cent, attr


def zoo_dataset(path):
    """UCI zoo as prepared by zoo_simulator.R:18-37: +1 shift, legs recoded to 1..6.
    Returns (X float64 101 x 16, attrisize, v, w, gamma, ground truth classes)."""
    rows = [ln.strip().split(",") for ln in open(path) if ln.strip()]
    raw = np.array([[int(x) for x in r[1:]] for r in rows])
    gt = raw[:, 16]
    zoo = raw[:, :16] + 1
    legs = zoo[:, 12].copy()
    rec = np.ones_like(legs)
    for src, dst in ((3, 2), (5, 3), (6, 4), (7, 5), (9, 6)):
        rec[legs == src] = dst
    zoo[:, 12] = rec
    attr = np.array([len(np.unique(zoo[:, j])) for j in range(16)], dtype=np.int32)
    v = np.array([6.0] * 12 + [3.0] + [6.0] * 3)
    w = np.array([0.25] * 12 + [0.5] + [0.25] * 3)
    return zoo.astype(np.float64), attr, v, w, 0.68, gt
