"""Host-side mirror of the reference's only entry point, `run_markov_chain`
(code/launcher.cpp:6-14), plus a step-wise `Chain` handle used by tests and bench.

Same argument names, defaults and meaning as the Rcpp export; the returned dict has
the eight fields of the reference's R list (code/launcher.cpp:57-63):
total_cls, c_i, centers, sigmas, loglikelihood, final_ass, time, accepted.
Errors of the reference (Rcpp::stop from validate_state, bad probabilities) surface
as SmgError with the same message text.
"""
import ctypes as C

import numpy as np

from . import _lib as lb


def _colmajor(data):
    """R NumericMatrix layout: column-major float64 n x p."""
    a = np.asarray(data, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("data must be an n x p matrix")
    return np.asfortranarray(a)


def run_markov_chain(data, attrisize, gamma, v, w, verbose=0, m=5, iterations=1000, L=1, c_i=None, burnin=5000, t=10,
                     r=10, neal8=False, split_merge=True, n8_step_size=1, sam_step_size=1, thinning=1, seed=None,
                     device=0):
    lib = lb.load()
    X = _colmajor(data)
    n, p = X.shape
    attr = lb.as_i32(attrisize)
    vv, ww = lb.as_f64(v), lb.as_f64(w)
    if attr.shape != (p,) or vv.shape != (p,) or ww.shape != (p,):
        raise ValueError("attrisize, v and w must have length ncol(data)")
    ci = None if c_i is None else lb.as_i32(c_i)
    if ci is not None and ci.shape != (n,):
        raise ValueError("c_i must have length nrow(data)")
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array(
            [1, 1 << 32], dtype=np.uint64))
    res = lb.SmgResults()
    rc = lib.smg_run_markov_chain(X.ctypes.data_as(lb.c_dbl_p), n, p, lb.iptr(attr), float(gamma), lb.dptr(vv), lb.dptr(ww),
                                  int(verbose), int(m), int(iterations), int(L), lb.iptr(ci), int(burnin), int(t),
                                  int(r), int(bool(neal8)), int(bool(split_merge)), int(n8_step_size), int(sam_step_size),
                                  int(thinning), C.c_ulonglong(int(seed) & (2**64 - 1)), int(device), C.byref(res))
    partial_err = None
    if rc == 4 and res.iterations > 0 and res.total_cls:  # SMG_ERR_CAPACITY with the iterations kept so far
        partial_err = lb.SmgError(rc, lib.smg_last_error().decode("utf-8", "replace"))
    else:
        lb.check(rc)
    # The big result arrays (allocation trace, centres, sigmas) are NOT copied: the numpy arrays below are views of the
    # library's buffers, which are released when the last of them is garbage collected.
    owner = _ResultsOwner(lib, res)
    it = res.iterations
    total = _view(res.total_cls, max(it, 1), C.c_int, owner)[:it]
    c_all = _view(res.c_i, max(it, 1) * n, C.c_int, owner).reshape(max(it, 1), n)[:it]
    off = _view(res.phi_offset, it + 1, C.c_longlong, owner)
    nphi = int(off[-1])
    cen = _view(res.centers, max(nphi, 1) * p, C.c_double, owner).reshape(max(nphi, 1), p)[:nphi]
    sig = _view(res.sigmas, max(nphi, 1) * p, C.c_double, owner).reshape(max(nphi, 1), p)[:nphi]
    offl = off.tolist()
    out = {
        "total_cls": total.tolist(),
        "c_i": list(c_all),
        "centers": [list(cen[offl[i]:offl[i + 1]]) for i in range(it)],
        "sigmas": [list(sig[offl[i]:offl[i + 1]]) for i in range(it)],
        "loglikelihood": _view(res.loglikelihood, max(it, 1), C.c_double, owner)[:it].copy(),
        "final_ass": _view(res.final_ass, n, C.c_int, owner).copy(),
        "time": int(res.time),
        "accepted": _view(res.accepted, max(it, 1), C.c_int, owner)[:it].copy(),
        "seconds": float(res.seconds),
    }
    if partial_err is not None:  # the chain outgrew the cluster capacity: the iterations completed so far travel with the error
        partial_err.partial = out
        raise partial_err
    return out


class _ResultsOwner:
    """Keeps an smg_results alive for as long as a numpy view of one of its buffers exists."""

    def __init__(self, lib, res):
        self.lib, self.res = lib, res

    def __del__(self):
        try:
            self.lib.smg_free_results(C.byref(self.res))
        except Exception:
            pass


def _view(ptr, count, ctype, owner):
    if not ptr:  # nothing was allocated (failed or empty run)
        return np.zeros(count, dtype=np.ctypeslib.as_array((ctype * 1)()).dtype)
    buf = (ctype * count).from_address(C.addressof(ptr.contents))
    buf._owner = owner  # numpy keeps `buf` as the base of the array, `buf` keeps the owner
    return np.ctypeslib.as_array(buf)


class Chain:
    """One chain resident on one GPU (smg_create / smg_step / smg_snapshot / smg_destroy)."""

    def __init__(self, data, attrisize, gamma, v, w, m=5, L=1, c_i=None, t=10, r=10, neal8=True, split_merge=True,
                 n8_step_size=1, sam_step_size=1, thinning=1, seed=1, max_clusters=0, pool_size=0, device=0,
                 compact_init=False, data_u8=False, exact_sigma_inverse=False, pair_selection="random", aux_mode="pool"):
        self.lib = lb.load()
        self._attr = lb.as_i32(attrisize)
        self._v, self._w = lb.as_f64(v), lb.as_f64(w)
        if data_u8:
            X = np.ascontiguousarray(data, dtype=np.uint8)
        else:
            X = _colmajor(data)
        self.n, self.p = X.shape
        self.m = int(m)
        self.t, self.r = int(t), int(r)
        cfg = lb.SmgConfig(self.n, self.p, lb.iptr(self._attr), float(gamma), lb.dptr(self._v), lb.dptr(self._w), int(m),
                          int(L), int(t), int(r), int(bool(neal8)), int(bool(split_merge)), int(n8_step_size),
                          int(sam_step_size), int(thinning), int(seed) & (2**64 - 1), int(max_clusters), int(pool_size),
                          int(device), int(bool(compact_init)), int(bool(exact_sigma_inverse)),
                          {"random": 0, "deterministic": 1}[pair_selection], {"pool": 0, "philox": 1}[aux_mode])
        ci = None if c_i is None else lb.as_i32(c_i)
        h = C.c_void_p()
        if data_u8:
            rc = self.lib.smg_create_u8(C.byref(cfg), X.ctypes.data_as(C.POINTER(C.c_ubyte)), lb.iptr(ci), C.byref(h))
        else:
            rc = self.lib.smg_create(C.byref(cfg), X.ctypes.data_as(lb.c_dbl_p), lb.iptr(ci), C.byref(h))
        lb.check(rc)
        self.h = h
        self.kcap = 256

    def close(self):
        if getattr(self, "h", None):
            self.lib.smg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n_iters=1):
        lb.check(self.lib.smg_step(self.h, int(n_iters)))

    def validate_state(self):
        """The reference's validate_state invariant on the device state; raises SmgError on failure."""
        lb.check(self.lib.smg_validate_state(self.h))

    def checkpoint(self):
        """State needed to continue this chain later: iteration counter + snapshot (see `resume`)."""
        it = C.c_longlong()
        lb.check(self.lib.smg_get_iteration(self.h, C.byref(it)))
        s = self.snapshot()
        return {"iteration": it.value, "K": s["K"], "c_i": s["c_i"], "centers": s["centers"], "sigmas": s["sigmas"]}

    def resume(self, ckpt):
        """Continue from `checkpoint()` of a chain created with the same data, configuration and seed."""
        self.set_state(ckpt["K"], ckpt["c_i"], ckpt["centers"], ckpt["sigmas"])
        lb.check(self.lib.smg_resume_at(self.h, int(ckpt["iteration"])))

    def last_step_ms(self):
        ms = C.c_double()
        lb.check(self.lib.smg_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def snapshot(self, with_phi=True, with_c=True):
        K = C.c_int()
        ll = C.c_double()
        acc = C.c_int()
        c = np.empty(self.n, dtype=np.int32) if with_c else None
        cen = np.empty((self.kcap, self.p)) if with_phi else None
        sig = np.empty((self.kcap, self.p)) if with_phi else None
        lb.check(self.lib.smg_snapshot(self.h, C.byref(K), lb.iptr(c), lb.dptr(cen), lb.dptr(sig), self.kcap, C.byref(ll),
                                      C.byref(acc)))
        k = K.value
        return {"K": k, "c_i": c, "centers": None if cen is None else cen[:k].copy(),
                "sigmas": None if sig is None else sig[:k].copy(), "loglikelihood": ll.value, "accepted": acc.value}

    def stats(self):
        out = np.zeros(8, dtype=np.uint64)
        lb.check(self.lib.smg_get_stats(self.h, out.ctypes.data_as(lb.c_ull_p)))
        keys = ["scan_rounds", "scan_events", "births", "deaths", "sweeps", "launches", "sm_proposals", "sm_accepted"]
        return dict(zip(keys, (int(x) for x in out)))

    def scan_spec(self, mode=-1):
        """Set the scan's speculative-evaluation mode (0 off, 1 on, 2 on + self-check, -1 library default) and return its
        counters since creation."""
        out = np.zeros(4, dtype=np.uint64)
        lb.check(self.lib.smg_debug_scan_spec(self.h, int(mode), out.ctypes.data_as(lb.c_ull_p)))
        return {"mismatches": int(out[0]), "reevaluated": int(out[1]), "dropped": int(out[2])}

    def scan_profile(self):
        out = np.zeros(16, dtype=np.uint64)
        lb.check(self.lib.smg_debug_scan_profile(self.h, out.ctypes.data_as(lb.c_ull_p)))
        keys = ["prologue", "screen", "pick", "evaluate", "detect", "apply", "loop", "rows", "e8", "e9", "e10", "e11", "e12", "e13",
                "e14", "e15"]
        return dict(zip(keys, (int(x) for x in out)))

    def time_ll_block(self, reps=5):
        """Device time (ms) of the likelihood-block kernel alone on the current state, cold caches."""
        out = C.c_double()
        lb.check(self.lib.smg_debug_time_ll_block(self.h, int(reps), C.byref(out)))
        return out.value

    def sm_profile(self):
        out = np.zeros(64, dtype=np.uint64)
        lb.check(self.lib.smg_debug_sm_profile(self.h, out.ctypes.data_as(lb.c_ull_p)))
        names = {0: "select+prior", 1: "bar_a", 2: "S_write", 3: "bar_b", 4: "members+rows+hist", 5: "bar_c", 6: "table",
                 7: "dots|mg_draws", 8: "compact", 9: "bar1", 10: "walk", 11: "moves", 12: "bar2", 13: "logits|slices",
                 14: "draws", 15: "mg_draws", 16: "bar3", 17: "tail", 18: "bar_f", 19: "mh_load", 20: "mh_terms", 21: "bar_h", 22: "mh_sums",
                 23: "bar_i", 24: "accept", 25: "#scans_with_walk", 26: "#scans_with_change", 27: "#scans", 28: "#listed_cta0"}
        keys = {k: "M." + v for k, v in names.items()}
        keys.update({32 + k: "P." + v for k, v in names.items()})
        keys.update({50: "D.entry", 51: "D.philox", 52: "D.centre", 53: "D.sigma", 54: "D.tail", 55: "D.publish"})
        keys.update({58: "launches", 59: "nnr_total", 60: "nnr_scan0", 61: "nnr_scan1", 62: "nnr_scan2", 63: "nnr_scan3"})
        return {name: int(out[k]) for k, name in keys.items()}

    def timings(self):
        out = np.zeros(8)
        lb.check(self.lib.smg_get_timings(self.h, lb.dptr(out)))
        keys = ["ll_block_ms", "aux_ll_ms", "scan_ms", "update_phi_ms", "split_merge_ms", "pool_ms", "loglik_ms", "total_ms"]
        return dict(zip(keys, (float(x) for x in out)))

    # ---- parity hooks ---------------------------------------------------------------------
    def set_state(self, K, c_i, centers, sigmas):
        c = lb.as_i32(c_i)
        ce, sg = lb.as_f64(centers), lb.as_f64(sigmas)
        lb.check(self.lib.smg_debug_set_state(self.h, int(K), lb.iptr(c), lb.dptr(ce), lb.dptr(sg)))

    def set_pool(self, pool_center, pool_sigma):
        pc, ps = lb.as_f64(pool_center), lb.as_f64(pool_sigma)
        lb.check(self.lib.smg_debug_set_pool(self.h, pc.shape[0], lb.dptr(pc), lb.dptr(ps)))

    def get_pool(self, first, count):
        pc = np.empty((count, self.p))
        ps = np.empty((count, self.p))
        lb.check(self.lib.smg_debug_get_pool(self.h, int(first), int(count), lb.dptr(pc), lb.dptr(ps)))
        return pc, ps

    def ll_block(self, K):
        LL = np.empty((self.n, K))
        mm = np.empty((self.n, K), dtype=np.int32)
        lb.check(self.lib.smg_debug_ll_block(self.h, lb.dptr(LL), lb.iptr(mm)))
        return LL, mm

    def neal8_scan(self, tape=None):
        t = None if tape is None else lb.as_f64(tape)
        if t is not None and t.size != self.n * (self.m + 1):
            raise ValueError("tape must hold n*(m+1) uniforms")
        lb.check(self.lib.smg_debug_neal8_scan(self.h, lb.dptr(t)))

    def histogram(self, K):
        mm = C.c_int()
        lb.check(self.lib.smg_debug_histogram(self.h, None, None, C.byref(mm)))
        H = np.empty((K, self.p, mm.value), dtype=np.int32)
        cnt = np.empty(K, dtype=np.int32)
        lb.check(self.lib.smg_debug_histogram(self.h, lb.iptr(H), lb.iptr(cnt), C.byref(mm)))
        return H, cnt

    def update_phi(self, u_center=None, u_sigma=None):
        uc, us = lb.as_f64(u_center), lb.as_f64(u_sigma)
        lb.check(self.lib.smg_debug_update_phi(self.h, lb.dptr(uc), lb.dptr(us)))

    def loglik(self):
        out = C.c_double()
        lb.check(self.lib.smg_debug_loglik(self.h, C.byref(out)))
        return out.value

    def split_merge(self, tape=None):
        """tape: dict with any of u_pair,u_prior_c,u_prior_s,u_launch,u_rg,u_rg_c,u_rg_s,u_mg_c,u_mg_s,u_accept."""
        T = lb.SmgSmTape()
        keep = []
        if tape:
            for k, a in tape.items():
                arr = lb.as_f64(a)
                keep.append(arr)
                setattr(T, k, lb.dptr(arr))
        info = np.zeros(8, dtype=np.int32)
        S = np.zeros(self.n, dtype=np.int32)
        zl = np.zeros(self.n, dtype=np.int32)
        zs = np.zeros(self.n, dtype=np.int32)
        phi = np.zeros((6, 2, self.p))
        terms = np.zeros(24)
        lb.check(self.lib.smg_debug_split_merge(self.h, C.byref(T) if tape else None, lb.iptr(info), lb.iptr(S), lb.iptr(zl),
                                               lb.iptr(zs), lb.dptr(phi), lb.dptr(terms)))
        nS = int(info[2])
        return {"i1": int(info[0]), "i2": int(info[1]), "nS": nS, "is_split": int(info[3]), "accepted": int(info[4]),
                "nA": int(info[5]), "nB": int(info[6]), "K": int(info[7]), "S": S[:nS].copy(), "z_launch": zl[:nS].copy(),
                "z_star": zs[:nS].copy(), "phi": phi, "terms": terms}

    def aux_free(self, count):
        """aux_mode='philox': column values and parameters of the first `count` auxiliary components of the next pass."""
        ll = np.zeros(count)
        cen = np.zeros((count, self.p))
        sig = np.zeros((count, self.p))
        lb.check(self.lib.smg_debug_aux_free(self.h, int(count), lb.dptr(ll), lb.dptr(cen), lb.dptr(sig)))
        return ll, cen, sig

    def sm_terms(self, u_pair, z_launch, z_star, phi6, u_accept=0.5):
        """MH ratio alone (smg_debug_sm_terms) on an injected launch / proposal: phi6 [6][2][p] = (centre, sigma) of
        split-launch A, B, merge-launch M, proposal A*, B*, M*; sides by position in S.  Returns info + terms[24]."""
        up = lb.as_f64(u_pair)
        zl = lb.as_i32(z_launch if len(z_launch) else np.zeros(1, dtype=np.int32))
        zs = lb.as_i32(z_star if len(z_star) else np.zeros(1, dtype=np.int32))
        ph = lb.as_f64(phi6)
        assert ph.shape == (6, 2, self.p)
        info = np.zeros(8, dtype=np.int32)
        terms = np.zeros(24)
        lb.check(self.lib.smg_debug_sm_terms(self.h, lb.dptr(up), lb.iptr(zl), lb.iptr(zs), lb.dptr(ph), float(u_accept),
                                            lb.iptr(info), lb.dptr(terms)))
        return {"i1": int(info[0]), "i2": int(info[1]), "nS": int(info[2]), "is_split": int(info[3]),
                "accepted": int(info[4]), "terms": terms}


def synth_generate(n, p, attrisize, k_true, s=0.5, seed=1, device=0):
    """Synthetic Hamming-mixture data generated on the GPU (smg_synth_generate): X uint8 [n][p], labels, centres."""
    lib = lb.load()
    attr = lb.as_i32(np.full(p, attrisize) if np.isscalar(attrisize) else attrisize)
    X = np.empty((n, p), dtype=np.uint8)
    lab = np.empty(n, dtype=np.int32)
    cen = np.empty((k_true, p), dtype=np.uint8)
    lb.check(lib.smg_synth_generate(int(n), int(p), lb.iptr(attr), int(k_true), float(s), int(seed) & (2**64 - 1), int(device),
                                    X.ctypes.data_as(C.POINTER(C.c_ubyte)), lb.iptr(lab), cen.ctypes.data_as(C.POINTER(C.c_ubyte))))
    return X, lab, cen, attr


def initial_assignment(u, L, device=0):
    """sample_initial_assignment (common_functions.cpp:174-183) on the device under injected uniforms (parity hook)."""
    uu = lb.as_f64(u)
    out = np.zeros(uu.size, dtype=np.int32)
    lb.check(lb.load().smg_debug_initial_assignment(int(uu.size), int(L), lb.dptr(uu), int(device), lb.iptr(out)))
    return out


def step_many(chains, n_iters=1):
    """n_iters iterations on every chain of the list, overlapped on the GPU (smg_step_many)."""
    if not chains:
        return
    arr = (C.c_void_p * len(chains))(*[ch.h for ch in chains])
    lb.check(lb.load().smg_step_many(arr, len(chains), int(n_iters)))


class Psm:
    """Posterior similarity matrix accumulated on the tensor cores (smg_psm_*): int32 co-clustering counts.
    `external` may be a torch int32 CUDA tensor of shape (n, n) that the caller owns (e.g. to all-reduce it over
    GPUs with NCCL afterwards); otherwise the library owns the matrix."""

    def __init__(self, n, device=0, capacity_sweeps=64, external=None):
        self.lib = lb.load()
        self.n = int(n)
        self._keep = external
        ptr = None
        if external is not None:
            if tuple(external.shape) != (self.n, self.n) or external.element_size() != 4 or not external.is_contiguous():
                raise ValueError("external must be a contiguous int32 (n, n) CUDA tensor")
            ptr = C.c_void_p(external.data_ptr())
        h = C.c_void_p()
        lb.check(self.lib.smg_psm_create(self.n, int(device), int(capacity_sweeps), ptr, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.smg_psm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push_chain(self, chain):
        lb.check(self.lib.smg_psm_push_chain(self.h, chain.h))

    def push(self, c_i):
        c = lb.as_i32(c_i)
        if c.shape != (self.n,):
            raise ValueError("c_i must have length n")
        lb.check(self.lib.smg_psm_push_host(self.h, lb.iptr(c)))

    def flush(self, finalize=True):
        """Fold the buffered sweeps into the matrix; finalize=True also completes the lower triangle (the tensor-core
        kernel accumulates the tiles on or above the diagonal only), which readers of an external matrix need."""
        lb.check(self.lib.smg_psm_finalize(self.h) if finalize else self.lib.smg_psm_flush(self.h))

    def read(self, row0=0, nrows=None):
        nrows = self.n - row0 if nrows is None else int(nrows)
        out = np.empty((nrows, self.n), dtype=np.int32)
        lb.check(self.lib.smg_psm_read(self.h, int(row0), nrows, lb.iptr(out)))
        return out

    def info(self):
        sw = C.c_longlong()
        ms = C.c_double()
        nl = C.c_ulonglong()
        lb.check(self.lib.smg_psm_info(self.h, C.byref(sw), C.byref(ms), C.byref(nl)))
        return {"sweeps": sw.value, "last_flush_ms": ms.value, "launches": nl.value}

    def point_estimate(self, candidates, draws, row0=0, nrows=None):
        """Binder loss (x draws, exact integer) and n x the VI lower bound of each candidate allocation over the rows
        [row0, row0 + nrows) of the matrix (smg_psm_point_estimate; mcclust minbinder / minVI with method='draws')."""
        cand = np.ascontiguousarray(np.atleast_2d(candidates), dtype=np.int32)
        assert cand.shape[1] == self.n
        nrows = self.n - row0 if nrows is None else int(nrows)
        b = np.zeros(cand.shape[0], dtype=np.int64)
        v = np.zeros(cand.shape[0])
        lb.check(self.lib.smg_psm_point_estimate(self.h, int(row0), nrows, lb.iptr(cand), cand.shape[0], int(draws),
                                                b.ctypes.data_as(lb.c_ll_p), lb.dptr(v)))
        return b, v

    def reference(self):
        """CUDA-core evaluation of the currently buffered sweeps (n x n), for cross-checks."""
        out = np.empty((self.n, self.n), dtype=np.int32)
        lb.check(self.lib.smg_debug_psm_reference(self.h, lb.iptr(out)))
        return out


class Comm:
    """Communicator of the library's own NCCL reductions (include/smgibbs.h, smg_comm_* / smg_chains_*): PSM
    reduce-scatter / all-reduce, split-R-hat, K histogram.  `unique_id()` on rank 0, ship the 128 bytes to the other
    ranks with any transport (torch.distributed broadcast, a file, MPI), then Comm(rank, world, id, device)."""

    @staticmethod
    def unique_id():
        lib = lb.load()
        buf = C.create_string_buffer(128)
        lb.check(lib.smg_comm_unique_id(buf))
        return buf.raw

    def __init__(self, rank=0, world=1, unique_id=None, device=0):
        self.lib = lb.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        lb.check(self.lib.smg_comm_create(self.rank, self.world, unique_id, self.device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.smg_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def distribute_psm(self, psm):
        """Fused accumulation + reduce-scatter: from now on every flush of `psm` adds its tiles straight into the memory
        of the rank that owns the rows (peer-mapped over NVLink).  Call on a fresh matrix, on every rank."""
        lb.check(self.lib.smg_chains_psm_distribute(self.h, psm.h))

    def reduce_psm(self, psm, mode="reduce_scatter"):
        """Sums the PSM counts of all ranks in place.  Returns (row0, nrows, ms, bus_GBps): the rows of the reduced matrix
        this rank holds (all of them after an all-reduce)."""
        r0, nr, ms, bus = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        lb.check(self.lib.smg_chains_reduce_psm(self.h, psm.h, 1 if mode == "reduce_scatter" else 0, C.byref(r0), C.byref(nr),
                                                C.byref(ms), C.byref(bus)))
        return r0.value, nr.value, ms.value, bus.value

    def split_rhat(self, traces):
        """Split-R-hat over the chains of all ranks; traces: [local chains][draws]."""
        x = np.ascontiguousarray(traces, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("traces must be [chains][draws]")
        out, tot = C.c_double(), C.c_longlong()
        lb.check(self.lib.smg_chains_split_rhat(self.h, lb.dptr(x), x.shape[0], x.shape[1], C.byref(out), C.byref(tot)))
        return out.value, tot.value

    def point_estimate(self, psm, candidates, draws, row0=0, nrows=None):
        """Expected Binder loss and VI lower bound of each candidate over ALL ranks' row blocks, and the minimisers."""
        cand = np.ascontiguousarray(np.atleast_2d(candidates), dtype=np.int32)
        nrows = psm.n - row0 if nrows is None else int(nrows)
        b, v = np.zeros(cand.shape[0]), np.zeros(cand.shape[0])
        bb, bv = C.c_int(), C.c_int()
        lb.check(self.lib.smg_chains_point_estimate(self.h, psm.h, int(row0), nrows, lb.iptr(cand), cand.shape[0], int(draws),
                                                   lb.dptr(b), lb.dptr(v), C.byref(bb), C.byref(bv)))
        return {"binder": b, "vi_lower_bound": v, "best_binder": bb.value, "best_vi": bv.value}

    def k_histogram(self, K, kmax=256):
        k = np.ascontiguousarray(np.asarray(K).ravel(), dtype=np.int32)
        hist = np.zeros(kmax + 1, dtype=np.int64)
        over = C.c_longlong()
        lb.check(self.lib.smg_chains_k_histogram(self.h, lb.iptr(k), k.size, kmax, hist.ctypes.data_as(lb.c_ll_p), C.byref(over)))
        return hist, over.value


def adjusted_rand_index(a, b, device=0):
    """Hubert-Arabie adjusted Rand index on the device (smg_adjusted_rand_index; mcclust::arandi)."""
    aa, bb = lb.as_i32(a), lb.as_i32(b)
    out = C.c_double()
    lb.check(lb.load().smg_adjusted_rand_index(lb.iptr(aa), lb.iptr(bb), int(aa.size), int(device), C.byref(out)))
    return out.value


def trace_ess(traces, device=0):
    """(IAT, ESS) of scalar traces [ntraces][T] on the device (smg_trace_ess; LaplacesDemon::IAT / ESS)."""
    x = np.ascontiguousarray(np.atleast_2d(traces), dtype=np.float64)
    iat, ess = np.zeros(x.shape[0]), np.zeros(x.shape[0])
    lb.check(lb.load().smg_trace_ess(lb.dptr(x), x.shape[0], x.shape[1], int(device), lb.dptr(iat), lb.dptr(ess)))
    return iat, ess


def hig_inv_u(omega, v, w, m):
    lib = lb.load()
    om, vv, ww, mm = (np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.shape(omega)))
                      for x in (omega, v, w, m))
    out = np.empty(om.shape)
    lb.check(lib.smg_debug_hig_inv_u(om.size, lb.dptr(om), lb.dptr(vv), lb.dptr(ww), lb.dptr(mm), lb.dptr(out)))
    return out


def logdensity_hig(sigma, v, w, m):
    lib = lb.load()
    sg, vv, ww, mm = (np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.shape(sigma)))
                      for x in (sigma, v, w, m))
    out = np.empty(sg.shape)
    lb.check(lib.smg_debug_logdensity_hig(sg.size, lb.dptr(sg), lb.dptr(vv), lb.dptr(ww), lb.dptr(mm), lb.dptr(out)))
    return out


def rhig_u(count, v, w, m, seed=1):
    """`count` draws of u = exp(-1/sigma), sigma ~ HIG(v,w,m), from the device's production sampler."""
    lib = lb.load()
    out = np.empty(int(count))
    lb.check(lib.smg_debug_rhig_u(int(count), float(v), float(w), float(m), int(seed), lb.dptr(out)))
    return out
