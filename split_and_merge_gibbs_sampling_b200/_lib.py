"""ctypes binding of libsmgibbs.so (include/smgibbs.h).

The shared library is the product; this module only marshals numpy arrays to its
C ABI.  There is no Python/CPU implementation behind it: if the library is
missing or no CUDA device is usable every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMG_LIB_PATH") or os.path.join(_HERE, "libsmgibbs.so")  # (override: instrumented builds)

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_ull_p = C.POINTER(C.c_ulonglong)
c_ll_p = C.POINTER(C.c_longlong)


class SmgConfig(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("p", C.c_int), ("attrisize", c_int_p), ("gamma", C.c_double), ("v", c_dbl_p), ("w", c_dbl_p),
        ("m_aux", C.c_int), ("L", C.c_int), ("t", C.c_int), ("r", C.c_int), ("neal8", C.c_int), ("split_merge", C.c_int),
        ("n8_step_size", C.c_int), ("sam_step_size", C.c_int), ("thinning", C.c_int), ("seed", C.c_ulonglong),
        ("max_clusters", C.c_int), ("pool_size", C.c_longlong), ("device", C.c_int), ("compact_init", C.c_int), ("exact_sigma_inverse", C.c_int),
        ("pair_selection", C.c_int), ("aux_mode", C.c_int),
    ]


class SmgResults(C.Structure):
    _fields_ = [
        ("iterations", C.c_int), ("n", C.c_int), ("p", C.c_int), ("total_cls", c_int_p), ("c_i", c_int_p),
        ("phi_offset", c_ll_p), ("centers", c_dbl_p), ("sigmas", c_dbl_p), ("loglikelihood", c_dbl_p),
        ("final_ass", c_int_p), ("time", C.c_longlong), ("accepted", c_int_p), ("seconds", C.c_double),
    ]


class SmgSmTape(C.Structure):
    _fields_ = [(k, c_dbl_p) for k in ("u_pair", "u_prior_c", "u_prior_s", "u_launch", "u_rg", "u_rg_c", "u_rg_s",
                                       "u_mg_c", "u_mg_s", "u_accept")]


EXPORTS = [
    "smg_last_error", "smg_device_count", "smg_run_markov_chain", "smg_free_results", "smg_create", "smg_create_u8",
    "smg_step", "smg_step_many", "smg_get_iteration", "smg_resume_at", "smg_validate_state", "smg_synth_generate", "smg_snapshot", "smg_destroy", "smg_get_stats", "smg_get_timings", "smg_last_step_ms", "smg_debug_set_state",
    "smg_debug_set_pool", "smg_debug_get_pool", "smg_debug_ll_block", "smg_debug_neal8_scan", "smg_debug_histogram",
    "smg_debug_update_phi", "smg_debug_loglik", "smg_debug_hig_inv_u", "smg_debug_logdensity_hig", "smg_debug_rhig_u",
    "smg_debug_split_merge", "smg_debug_sm_terms", "smg_debug_aux_free", "smg_debug_initial_assignment", "smg_debug_time_ll_block", "smg_debug_scan_profile", "smg_debug_scan_spec", "smg_debug_sm_profile", "smg_psm_create", "smg_psm_push_chain", "smg_psm_push_host",
    "smg_psm_flush", "smg_psm_finalize", "smg_psm_read", "smg_psm_info", "smg_psm_destroy", "smg_debug_psm_reference",
    "smg_comm_unique_id", "smg_comm_create", "smg_comm_destroy", "smg_chains_reduce_psm", "smg_chains_split_rhat",
    "smg_chains_k_histogram", "smg_chains_psm_distribute", "smg_psm_point_estimate", "smg_chains_point_estimate", "smg_adjusted_rand_index", "smg_trace_ess",
]

_lib = None


class SmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[smgibbs status {code}] {msg}")
        self.code = code


def load():
    """Load libsmgibbs.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.smg_last_error.restype = C.c_char_p
    lib.smg_device_count.restype = C.c_int
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError when a declared symbol is not exported
    lib.smg_run_markov_chain.restype = C.c_int
    lib.smg_run_markov_chain.argtypes = [c_dbl_p, C.c_int, C.c_int, c_int_p, C.c_double, c_dbl_p, c_dbl_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.POINTER(SmgResults)]
    lib.smg_free_results.argtypes = [C.POINTER(SmgResults)]
    lib.smg_free_results.restype = None
    lib.smg_create.argtypes = [C.POINTER(SmgConfig), c_dbl_p, c_int_p, C.POINTER(C.c_void_p)]
    lib.smg_create_u8.argtypes = [C.POINTER(SmgConfig), C.POINTER(C.c_ubyte), c_int_p, C.POINTER(C.c_void_p)]
    lib.smg_step.argtypes = [C.c_void_p, C.c_int]
    lib.smg_step_many.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
    lib.smg_get_iteration.argtypes = [C.c_void_p, c_ll_p]
    lib.smg_validate_state.argtypes = [C.c_void_p]
    lib.smg_synth_generate.argtypes = [C.c_int, C.c_int, c_int_p, C.c_int, C.c_double, C.c_ulonglong, C.c_int,
                                       C.POINTER(C.c_ubyte), c_int_p, C.POINTER(C.c_ubyte)]
    lib.smg_resume_at.argtypes = [C.c_void_p, C.c_longlong]
    lib.smg_snapshot.argtypes = [C.c_void_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, C.c_int, c_dbl_p, c_int_p]
    lib.smg_destroy.argtypes = [C.c_void_p]
    lib.smg_destroy.restype = None
    lib.smg_get_stats.argtypes = [C.c_void_p, c_ull_p]
    lib.smg_get_timings.argtypes = [C.c_void_p, c_dbl_p]
    lib.smg_last_step_ms.argtypes = [C.c_void_p, c_dbl_p]
    lib.smg_debug_scan_profile.argtypes = [C.c_void_p, c_ull_p]
    lib.smg_debug_scan_spec.argtypes = [C.c_void_p, C.c_int, c_ull_p]
    lib.smg_debug_time_ll_block.argtypes = [C.c_void_p, C.c_int, c_dbl_p]
    lib.smg_debug_sm_profile.argtypes = [C.c_void_p, c_ull_p]
    lib.smg_psm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.smg_psm_push_chain.argtypes = [C.c_void_p, C.c_void_p]
    lib.smg_psm_push_host.argtypes = [C.c_void_p, c_int_p]
    lib.smg_psm_flush.argtypes = [C.c_void_p]
    lib.smg_psm_finalize.argtypes = [C.c_void_p]
    lib.smg_psm_read.argtypes = [C.c_void_p, C.c_int, C.c_int, c_int_p]
    lib.smg_psm_info.argtypes = [C.c_void_p, c_ll_p, c_dbl_p, c_ull_p]
    lib.smg_psm_destroy.argtypes = [C.c_void_p]
    lib.smg_psm_destroy.restype = None
    lib.smg_debug_psm_reference.argtypes = [C.c_void_p, c_int_p]
    lib.smg_comm_unique_id.argtypes = [C.c_char_p]
    lib.smg_comm_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.smg_comm_destroy.argtypes = [C.c_void_p]
    lib.smg_comm_destroy.restype = None
    lib.smg_chains_reduce_psm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p]
    lib.smg_chains_psm_distribute.argtypes = [C.c_void_p, C.c_void_p]
    lib.smg_chains_split_rhat.argtypes = [C.c_void_p, c_dbl_p, C.c_int, C.c_int, c_dbl_p, c_ll_p]
    lib.smg_chains_k_histogram.argtypes = [C.c_void_p, c_int_p, C.c_longlong, C.c_int, c_ll_p, c_ll_p]
    lib.smg_psm_point_estimate.argtypes = [C.c_void_p, C.c_int, C.c_int, c_int_p, C.c_int, C.c_longlong, c_ll_p, c_dbl_p]
    lib.smg_chains_point_estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_int_p, C.c_int, C.c_longlong, c_dbl_p,
                                              c_dbl_p, c_int_p, c_int_p]
    lib.smg_adjusted_rand_index.argtypes = [c_int_p, c_int_p, C.c_int, C.c_int, c_dbl_p]
    lib.smg_trace_ess.argtypes = [c_dbl_p, C.c_int, C.c_int, C.c_int, c_dbl_p, c_dbl_p]
    lib.smg_debug_set_state.argtypes = [C.c_void_p, C.c_int, c_int_p, c_dbl_p, c_dbl_p]
    lib.smg_debug_set_pool.argtypes = [C.c_void_p, C.c_longlong, c_dbl_p, c_dbl_p]
    lib.smg_debug_get_pool.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, c_dbl_p, c_dbl_p]
    lib.smg_debug_ll_block.argtypes = [C.c_void_p, c_dbl_p, c_int_p]
    lib.smg_debug_neal8_scan.argtypes = [C.c_void_p, c_dbl_p]
    lib.smg_debug_histogram.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p]
    lib.smg_debug_update_phi.argtypes = [C.c_void_p, c_dbl_p, c_dbl_p]
    lib.smg_debug_loglik.argtypes = [C.c_void_p, c_dbl_p]
    lib.smg_debug_hig_inv_u.argtypes = [C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]
    lib.smg_debug_logdensity_hig.argtypes = [C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]
    lib.smg_debug_rhig_u.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_ulonglong, c_dbl_p]
    lib.smg_debug_initial_assignment.argtypes = [C.c_int, C.c_int, c_dbl_p, C.c_int, c_int_p]
    lib.smg_debug_aux_free.argtypes = [C.c_void_p, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p]
    lib.smg_debug_sm_terms.argtypes = [C.c_void_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p, C.c_double, c_int_p, c_dbl_p]
    lib.smg_debug_split_merge.argtypes = [C.c_void_p, C.POINTER(SmgSmTape), c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p,
                                          c_dbl_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SmgError(rc, load().smg_last_error().decode("utf-8", "replace"))


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def as_f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
