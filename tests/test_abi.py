"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/hankb200.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hankb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hank_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hankb200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.hank_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from hankb200.household import HouseholdBlock
    from hankb200._lib import HankError
    g = np.linspace(0, 1, 10)
    with pytest.raises(HankError) as ei:
        HouseholdBlock(g, np.ones(3), np.full((3, 3), 1 / 3), 0.98, 2.0, 0.0, 5)
    assert "no CPU fallback" in str(ei.value) or ei.value.code == 4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "julia-newtonraphsonhank_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
