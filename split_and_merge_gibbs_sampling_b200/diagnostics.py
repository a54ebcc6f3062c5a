"""Chain diagnostics and posterior summaries the reference's R scripts compute after the run
(realdata_analysis/zoo_simulator.R:205-215,339-344: LaplacesDemon::ESS / IAT, mcclust::arandi,
mcclust.ext::comp.psm + minVI).  Host-side numpy on the returned traces; the PSM itself comes from the
tensor-core accumulator (`Psm`).  SURVEY 8(f) rows 1 and 4 -- outside the hot path, kept small."""
import numpy as np


def autocovariance(x):
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    x = x - x.mean()
    f = np.fft.rfft(x, 2 * n)
    return np.fft.irfft(f * np.conj(f))[:n] / n


def iat(x):
    """Integrated autocorrelation time, Geyer's initial positive sequence estimator."""
    g = autocovariance(x)
    if g[0] <= 0:
        return 1.0
    rho = g / g[0]
    tau = -1.0
    for k in range(0, rho.size - 1, 2):
        pair = rho[k] + rho[k + 1]
        if pair <= 0:
            break
        tau += 2.0 * pair
    return float(max(tau, 1.0))


def ess(x):
    """Effective sample size n / IAT."""
    x = np.asarray(x)
    return float(x.size / iat(x))


def adjusted_rand_index(a, b):
    """Hubert-Arabie adjusted Rand index of two labelings (mcclust::arandi)."""
    a, b = np.asarray(a), np.asarray(b)
    _, ia = np.unique(a, return_inverse=True)
    _, ib = np.unique(b, return_inverse=True)
    ct = np.zeros((ia.max() + 1, ib.max() + 1), dtype=np.int64)
    np.add.at(ct, (ia, ib), 1)
    comb = lambda z: z * (z - 1) / 2.0
    sij = comb(ct).sum()
    sa, sb = comb(ct.sum(1)).sum(), comb(ct.sum(0)).sum()
    tot = comb(a.size)
    exp = sa * sb / tot if tot else 0.0
    mx = 0.5 * (sa + sb)
    return float((sij - exp) / (mx - exp)) if mx != exp else 1.0


def binder_point_estimate(psm_counts, draws, candidates):
    """Among the candidate allocations (e.g. the kept draws), the one minimising the posterior expected Binder
    loss  sum_{i<j} |[c_i = c_j] - psm_ij|  (the 'draws' variant of mcclust's minbinder).  psm_counts are the
    integer co-clustering counts over `draws` kept sweeps.  Returns (index, losses)."""
    P = np.asarray(psm_counts, dtype=np.float64) / float(draws)
    losses = []
    for c in candidates:
        c = np.asarray(c)
        same = (c[:, None] == c[None, :])
        losses.append(float(np.abs(same - P).sum() / 2.0))
    return int(np.argmin(losses)), np.array(losses)
