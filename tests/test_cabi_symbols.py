"""CPU-only: the shared library loads and exports exactly the entry points the public header
declares, and the ctypes table mirrors the header (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "cosmogp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from cosmogp_b200 import build, _lib
    lib = build.build()                      # no-op when up to date; nvcc cross-compiles without a GPU
    handle = ctypes.CDLL(lib)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), "libcosmogp_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert handle.cgp_version() >= 100


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the compute entry points must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosmogp_b200 import _lib
    with pytest.raises(_lib.CosmogpB200Error):
        _lib.require_device()
    import numpy as np
    import cosmogp_b200 as cg
    gp = cg.gaussian_process(np.sin(np.linspace(0, 5, 8)), np.linspace(0, 5, 8))      # host-only constructor works
    with pytest.raises(_lib.CosmogpB200Error):
        gp.compute_log_likelihood([1.0, 1.0])


def test_product_does_not_import_oracle():
    """Only tests/, bench.py and __graft_entry__.smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "cosmogp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, "%s mentions oracle" % f
