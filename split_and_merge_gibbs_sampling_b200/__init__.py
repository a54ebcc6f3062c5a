"""B200-native Gibbs + split-merge sampler for Dirichlet-process Hamming mixtures.

Drop-in for the hot path of Filippo-Galli/Split_and_merge_Gibbs_sampling behind its
own entry point `run_markov_chain` (code/launcher.cpp:6-14).  All computation runs
in hand-written sm_100a CUDA kernels inside libsmgibbs.so (C ABI: include/smgibbs.h).
"""
from .api import Chain, Comm, Psm, adjusted_rand_index, trace_ess, step_many, synth_generate, hig_inv_u, logdensity_hig, rhig_u, run_markov_chain  # noqa: F401
from .synth import ham_mix_gen, zoo_dataset  # noqa: F401
from ._lib import SmgError, LIB_PATH  # noqa: F401

__all__ = ["run_markov_chain", "Chain", "Comm", "Psm", "adjusted_rand_index", "trace_ess", "step_many", "ham_mix_gen", "zoo_dataset", "SmgError", "hig_inv_u", "logdensity_hig"]
