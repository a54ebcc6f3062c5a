"""ctypes access to the CPU oracle (oracle/_build/liboracle.so) -- test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)
llp = C.POINTER(C.c_longlong)


class Opts(C.Structure):
    _fields_ = [("counted", C.c_int), ("stable_hig", C.c_int), ("sigma_inverse_cdf", C.c_int), ("bisect_tol", C.c_double),
                ("bisect_max", C.c_int), ("validate", C.c_int), ("det_i1", C.c_int)]


def opts(counted=0, stable_hig=0, sigma_inverse_cdf=0, bisect_tol=1e-9, bisect_max=150, validate=1, det_i1=-1):
    return Opts(counted, stable_hig, sigma_inverse_cdf, bisect_tol, bisect_max, validate, det_i1)


class Data(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("X", dp), ("attrisize", ip), ("gamma", C.c_double), ("v", dp), ("w", dp)]


def build():
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "smg_oracle.hpp")]
    if (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_dhamming.restype = C.c_double
        L.orc_dhamming.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_hyp2f1.restype = C.c_double
        L.orc_hyp2f1.argtypes = [C.c_double] * 4 + [ip]
        L.orc_norm_const2.restype = C.c_double
        L.orc_norm_const2.argtypes = [C.c_double] * 3 + [C.c_int, ip]
        L.orc_logdensity_hig.restype = C.c_double
        L.orc_logdensity_hig.argtypes = [C.c_double] * 4 + [C.c_int, ip]
        L.orc_lF_conK2.restype = C.c_double
        L.orc_lF_conK2.argtypes = [C.c_double] * 5 + [C.c_int]
        L.orc_bisec_hyper2.restype = C.c_double
        L.orc_bisec_hyper2.argtypes = [C.c_double] * 4 + [C.POINTER(Opts), ip]
        L.orc_pbeta.restype = C.c_double
        L.orc_pbeta.argtypes = [C.c_double] * 3
        L.orc_log_ibeta.restype = C.c_double
        L.orc_log_ibeta.argtypes = [C.c_double] * 3
        L.orc_rhig_beta_branch.restype = C.c_int
        L.orc_rhig_beta_branch.argtypes = [C.c_double] * 3
        L.orc_revsort.argtypes = [dp, ip, C.c_int]
        L.orc_sample_probs_one.restype = C.c_int
        L.orc_sample_probs_one.argtypes = [dp, C.c_int, C.c_double, ip]
        L.orc_initial_assignment.argtypes = [C.c_int, C.c_int, dp, ip]
        L.orc_rhig_many.argtypes = [C.c_double] * 3 + [C.c_long, C.c_ulonglong, C.POINTER(Opts), dp]
        L.orc_rhig_u_from_omega.argtypes = [C.c_double] * 3 + [C.c_long, dp, C.POINTER(Opts), dp]
        L.orc_rbeta_many.argtypes = [C.c_double, C.c_double, C.c_long, C.c_ulonglong, dp]
        L.orc_loglik.restype = C.c_double
        L.orc_loglik.argtypes = [C.POINTER(Data), C.c_int, ip, dp, dp]
        L.orc_ll_block.argtypes = [C.POINTER(Data), C.c_int, dp, dp, dp, ip]
        L.orc_neal8_scan.argtypes = [C.POINTER(Data), C.c_int, C.c_int, ip, ip, dp, dp, C.c_long, dp, dp, dp, C.c_long,
                                     C.POINTER(Opts), llp, C.c_char_p, C.c_int]
        L.orc_histogram.argtypes = [C.POINTER(Data), C.c_int, ip, C.c_int, ip, ip]
        L.orc_update_phi.argtypes = [C.POINTER(Data), C.c_int, ip, dp, dp, ip, C.c_int, dp, C.c_long, lp, C.POINTER(Opts),
                                     C.c_long, ip, ip, ip, dp, lp, C.c_char_p, C.c_int]
        L.orc_prob_centers.argtypes = [C.POINTER(Data), C.c_int, ip, dp, C.c_int, dp, dp]
        L.orc_split_merge.argtypes = [C.POINTER(Data), C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_long, lp,
                                      C.POINTER(Opts), ip, ip, ip, dp, dp, ip, dp, dp, ip, dp, dp, dp, C.c_long, ip, ip, ip,
                                      ip, dp, lp, llp, C.c_char_p, C.c_int]
        L.orc_priors.restype = C.c_double
        L.orc_priors.argtypes = [C.POINTER(Data), dp, C.c_int, ip]
        L.orc_draw_pool.argtypes = [C.POINTER(Data), C.c_long, C.c_ulonglong, C.POINTER(Opts), dp, dp]
        L.orc_run_chain.argtypes = [C.POINTER(Data), C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(Opts), C.c_long, ip, ip, dp,
                                    ip, ip, C.c_int, dp, dp, dp, llp, C.c_char_p, C.c_int]
        L.orc_time_sweep.argtypes = [C.POINTER(Data), C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp, C.c_long, dp, dp,
                                     C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(Opts), dp, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def P(a, t=dp):
    return None if a is None else a.ctypes.data_as(t)


class OracleData:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, X, attrisize, gamma, v, w):
        self.X = np.asfortranarray(np.asarray(X, dtype=np.float64))
        self.n, self.p = self.X.shape
        self.attr = np.ascontiguousarray(attrisize, dtype=np.int32)
        self.v = np.ascontiguousarray(v, dtype=np.float64)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        self.gamma = float(gamma)
        self.c = Data(self.n, self.p, P(self.X), P(self.attr, ip), self.gamma, P(self.v), P(self.w))

    @property
    def ref(self):
        return C.byref(self.c)


class OracleError(RuntimeError):
    pass


def _err():
    return C.create_string_buffer(512)


def ll_block(d, center, sigma):
    K = center.shape[0]
    LL = np.empty((d.n, K))
    mm = np.empty((d.n, K), dtype=np.int32)
    ce, sg = np.ascontiguousarray(center, dtype=np.float64), np.ascontiguousarray(sigma, dtype=np.float64)
    lib().orc_ll_block(d.ref, K, P(ce), P(sg), P(LL), P(mm, ip))
    return LL, mm


def loglik(d, c, center, sigma):
    ce, sg = np.ascontiguousarray(center, dtype=np.float64), np.ascontiguousarray(sigma, dtype=np.float64)
    cc = np.ascontiguousarray(c, dtype=np.int32)
    return lib().orc_loglik(d.ref, ce.shape[0], P(cc, ip), P(ce), P(sg))


def neal8_scan(d, m_aux, c, center, sigma, pool_center, pool_sigma, tape, o=None, kcap=512):
    K = C.c_int(center.shape[0])
    cc = np.ascontiguousarray(c, dtype=np.int32).copy()
    ce = np.zeros((kcap, d.p))
    sg = np.zeros((kcap, d.p))
    ce[: K.value] = center
    sg[: K.value] = sigma
    pc, ps = np.ascontiguousarray(pool_center, dtype=np.float64), np.ascontiguousarray(pool_sigma, dtype=np.float64)
    tp = np.ascontiguousarray(tape, dtype=np.float64)
    diag = np.zeros(4, dtype=np.int64)
    e = _err()
    o = o or opts()
    rc = lib().orc_neal8_scan(d.ref, m_aux, kcap, C.byref(K), P(cc, ip), P(ce), P(sg), pc.shape[0], P(pc), P(ps), P(tp),
                              tp.size, C.byref(o), P(diag, llp), e, 512)
    if rc:
        raise OracleError(e.value.decode())
    k = K.value
    return {"K": k, "c": cc, "center": ce[:k].copy(), "sigma": sg[:k].copy(),
            "exact_pos_ties": int(diag[0]), "near_ties": int(diag[1])}


def histogram(d, K, c, mmax):
    H = np.zeros((K, d.p, mmax), dtype=np.int32)
    cnt = np.zeros(K, dtype=np.int32)
    cc = np.ascontiguousarray(c, dtype=np.int32)
    lib().orc_histogram(d.ref, K, P(cc, ip), mmax, P(H, ip), P(cnt, ip))
    return H, cnt


def update_phi(d, c, center, sigma, tape, clusters=(), o=None, log_cap=1 << 20):
    K = center.shape[0]
    cc = np.ascontiguousarray(c, dtype=np.int32)
    ce, sg = np.ascontiguousarray(center, dtype=np.float64).copy(), np.ascontiguousarray(sigma, dtype=np.float64).copy()
    cl = np.ascontiguousarray(clusters, dtype=np.int32)
    tp = np.ascontiguousarray(tape, dtype=np.float64)
    consumed = C.c_long()
    ls, la, lb_ = (np.zeros(log_cap, dtype=np.int32) for _ in range(3))
    lu = np.zeros(log_cap)
    ln = C.c_long()
    e = _err()
    o = o or opts()
    rc = lib().orc_update_phi(d.ref, K, P(cc, ip), P(ce), P(sg), P(cl, ip), cl.size, P(tp), tp.size, C.byref(consumed),
                              C.byref(o), log_cap, P(ls, ip), P(la, ip), P(lb_, ip), P(lu), C.byref(ln), e, 512)
    if rc:
        raise OracleError(e.value.decode())
    nl = ln.value
    return {"center": ce, "sigma": sg, "consumed": consumed.value,
            "log": {"site": ls[:nl], "a": la[:nl], "b": lb_[:nl], "u": lu[:nl]}}


def split_merge(d, t, r, c, center, sigma, tape, o=None, kcap=None, log_cap=1 << 22):
    K0 = center.shape[0]
    kcap = kcap or (K0 + 2)
    K = C.c_int(K0)
    cc = np.ascontiguousarray(c, dtype=np.int32).copy()

    def buf():
        return np.zeros((kcap, d.p)), np.zeros((kcap, d.p))

    ce, sg = buf()
    ce[:K0] = center
    sg[:K0] = sigma
    tp = np.ascontiguousarray(tape, dtype=np.float64)
    consumed = C.c_long()
    info = np.zeros(8, dtype=np.int32)
    S = np.zeros(d.n, dtype=np.int32)
    c_SL, c_ML, c_st = (np.zeros(d.n, dtype=np.int32) for _ in range(3))
    ce_SL, sg_SL = buf()
    ce_ML, sg_ML = buf()
    ce_st, sg_st = buf()
    terms = np.zeros(24)
    lph, ls, la, lb_ = (np.zeros(log_cap, dtype=np.int32) for _ in range(4))
    lu = np.zeros(log_cap)
    ln = C.c_long()
    diag = np.zeros(4, dtype=np.int64)
    e = _err()
    o = o or opts()
    rc = lib().orc_split_merge(d.ref, t, r, kcap, C.byref(K), P(cc, ip), P(ce), P(sg), P(tp), tp.size, C.byref(consumed),
                               C.byref(o), P(info, ip), P(S, ip), P(c_SL, ip), P(ce_SL), P(sg_SL), P(c_ML, ip), P(ce_ML),
                               P(sg_ML), P(c_st, ip), P(ce_st), P(sg_st), P(terms), log_cap, P(lph, ip), P(ls, ip),
                               P(la, ip), P(lb_, ip), P(lu), C.byref(ln), P(diag, llp), e, 512)
    if rc:
        raise OracleError(e.value.decode())
    nl = ln.value
    nS = int(info[2])
    return {"K": K.value, "c": cc, "center": ce[: K.value].copy(), "sigma": sg[: K.value].copy(),
            "i1": int(info[0]), "i2": int(info[1]), "S": S[:nS].copy(), "is_split": int(info[3]), "accepted": int(info[4]),
            "SL": {"K": int(info[5]), "c": c_SL, "center": ce_SL, "sigma": sg_SL},
            "ML": {"K": int(info[6]), "c": c_ML, "center": ce_ML, "sigma": sg_ML},
            "star": {"K": int(info[7]), "c": c_st, "center": ce_st, "sigma": sg_st},
            "terms": terms, "consumed": consumed.value,
            "log": {"phase": lph[:nl], "site": ls[:nl], "a": la[:nl], "b": lb_[:nl], "u": lu[:nl]}}


def initial_assignment(L, tape):
    """sample_initial_assignment (common_functions.cpp:174-183) from a uniform tape."""
    tp = np.ascontiguousarray(tape, dtype=np.float64)
    out = np.zeros(tp.size, dtype=np.int32)
    if lib().orc_initial_assignment(int(L), tp.size, P(tp), P(out, ip)):
        raise OracleError("initial_assignment failed")
    return out


def draw_pool(d, pool_size, seed, o=None):
    pc = np.zeros((pool_size, d.p))
    ps = np.zeros((pool_size, d.p))
    o = o or opts()
    rc = lib().orc_draw_pool(d.ref, pool_size, seed, C.byref(o), P(pc), P(ps))
    if rc:
        raise OracleError("draw_pool failed")
    return pc, ps


def run_chain(d, m_aux, iterations, L, c_init, burnin, t, r, neal8, split_merge_, seed, o=None, n8_step=1, sam_step=1,
              thinning=1, pool_size=0, keep_c=True, kcap=512):
    total = np.zeros(max(iterations, 1), dtype=np.int32)
    c_i = np.zeros((max(iterations, 1), d.n), dtype=np.int32) if keep_c else None
    ll = np.zeros(max(iterations, 1))
    acc = np.zeros(max(iterations, 1), dtype=np.int32)
    fin = np.zeros(d.n, dtype=np.int32)
    lc = np.zeros((kcap, d.p))
    ls = np.zeros((kcap, d.p))
    secs = C.c_double()
    diag = np.zeros(4, dtype=np.int64)
    e = _err()
    o = o or opts()
    ci = None if c_init is None else np.ascontiguousarray(c_init, dtype=np.int32)
    rc = lib().orc_run_chain(d.ref, m_aux, iterations, L, P(ci, ip), burnin, t, r, int(neal8), int(split_merge_), n8_step,
                             sam_step, thinning, seed, C.byref(o), pool_size, P(total, ip), P(c_i, ip), P(ll), P(acc, ip),
                             P(fin, ip), kcap, P(lc), P(ls), C.byref(secs), P(diag, llp), e, 512)
    if rc:
        raise OracleError(e.value.decode())
    return {"total_cls": total[:iterations], "c_i": None if c_i is None else c_i[:iterations], "loglikelihood": ll[:iterations],
            "accepted": acc[:iterations], "final_ass": fin, "seconds": secs.value, "diag": diag}


def time_sweep(d, m_aux, t, r, c, center, sigma, pool_center, pool_sigma, n_obs, do_sm, n_chains, seed, o=None):
    cc = np.ascontiguousarray(c, dtype=np.int32)
    ce, sg = np.ascontiguousarray(center, dtype=np.float64), np.ascontiguousarray(sigma, dtype=np.float64)
    pc, ps = np.ascontiguousarray(pool_center, dtype=np.float64), np.ascontiguousarray(pool_sigma, dtype=np.float64)
    out = np.zeros(5)
    e = _err()
    o = o or opts()
    rc = lib().orc_time_sweep(d.ref, m_aux, t, r, ce.shape[0], P(cc, ip), P(ce), P(sg), pc.shape[0], P(pc), P(ps), n_obs,
                              int(do_sm), n_chains, seed, C.byref(o), P(out), e, 512)
    if rc:
        raise OracleError(e.value.decode())
    return {"scan_s": out[0], "update_phi_s": out[1], "split_merge_s": out[2], "loglik_s": out[3], "n_obs": int(out[4])}
