"""The C-ABI library loads and exports every function include/smgibbs.h declares; without a GPU it
fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "smgibbs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("smg_run_markov_chain", "smg_create", "smg_step", "smg_snapshot", "smg_destroy", "smg_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from split_and_merge_gibbs_sampling_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/smgibbs.h but not exported"
    assert set(_lib.EXPORTS) == set(_declared())


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from split_and_merge_gibbs_sampling_b200 import Chain, SmgError
    X = np.ones((4, 2))
    with pytest.raises(SmgError, match="no CUDA device|CUDA"):
        Chain(X, [2, 2], 1.0, [6.0, 6.0], [0.25, 0.25])


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, link or load it
    pkg = os.path.join(ROOT, "split_and_merge_gibbs_sampling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "smg_oracle" not in txt and "oracle_lib" not in txt, f
