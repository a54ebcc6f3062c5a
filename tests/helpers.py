"""Shared builders for parity tests: synthetic problems and oracle-generated states."""
import numpy as np

import oracle_lib as orc
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen


class Problem:
    def __init__(self, n, p, m, k_true, seed=1, s=0.5, gamma=1.0, v=6.0, w=0.25):
        self.X, self.labels, self.cent, self.attr = ham_mix_gen(n, p, m, k_true, s=s, seed=seed)
        self.n, self.p = n, p
        self.gamma = gamma
        self.v = np.full(p, v)
        self.w = np.full(p, w)
        self.od = orc.OracleData(self.X, self.attr, gamma, self.v, self.w)

    def chain(self, **kw):
        from split_and_merge_gibbs_sampling_b200 import Chain
        kw.setdefault("m", 3)
        kw.setdefault("c_i", self.labels)
        return Chain(self.X.astype(np.float64), self.attr, self.gamma, self.v, self.w, **kw)


def oracle_state(pb, mode="truth", seed=11, L=None, iters=1, m_aux=3, pool_size=64):
    """A (K, c, center, sigma) state produced by the oracle itself.
    mode 'truth': start from the true labels (quiet regime); 'random': start from L random labels (burn-in)."""
    o = orc.opts(counted=1, stable_hig=1)
    c_init = pb.labels if mode == "truth" else None
    L = L or int(pb.labels.max() + 1)
    r = orc.run_chain(pb.od, m_aux, iters, L, c_init, 0, 2, 2, True, False, seed, o=o, pool_size=pool_size, kcap=512)
    c = r["final_ass"].copy()
    K = int(c.max() + 1)
    # parameters of the last kept iteration
    import ctypes as C
    total = np.zeros(max(iters, 1), dtype=np.int32)
    return K, c, r


def oracle_state_full(pb, mode="truth", seed=11, L=None, iters=1, m_aux=3, pool_size=64):
    """Same as oracle_state but also returns centres/sigmas (from the oracle's last snapshot)."""
    import ctypes as C
    o = orc.opts(counted=1, stable_hig=1)
    c_init = pb.labels if mode == "truth" else None
    L = L or int(pb.labels.max() + 1)
    kcap = 512
    d = pb.od
    total = np.zeros(max(iters, 1), dtype=np.int32)
    ll = np.zeros(max(iters, 1))
    acc = np.zeros(max(iters, 1), dtype=np.int32)
    fin = np.zeros(d.n, dtype=np.int32)
    lc = np.zeros((kcap, d.p))
    ls = np.zeros((kcap, d.p))
    secs = C.c_double()
    diag = np.zeros(4, dtype=np.int64)
    e = C.create_string_buffer(512)
    ci = None if c_init is None else np.ascontiguousarray(c_init, dtype=np.int32)
    rc = orc.lib().orc_run_chain(d.ref, m_aux, iters, L, orc.P(ci, orc.ip), 0, 2, 2, 1, 0, 1, 1, 1, seed, C.byref(o),
                                 pool_size, orc.P(total, orc.ip), None, orc.P(ll), orc.P(acc, orc.ip), orc.P(fin, orc.ip),
                                 kcap, orc.P(lc), orc.P(ls), C.byref(secs), orc.P(diag, orc.llp), e, 512)
    if rc:
        raise orc.OracleError(e.value.decode())
    K = int(total[iters - 1])
    return K, fin.copy(), lc[:K].copy(), ls[:K].copy()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
