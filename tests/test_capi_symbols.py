"""-m "not gpu": the C-ABI library loads and exports every symbol include/cannoles_b200.h
declares; without a device the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cannoles_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "cannoles_b200.h")).read()
    return sorted(set(re.findall(r"\b(b2b?_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    for name in _declared():
        assert hasattr(lib, name), f"{name} missing from libcannoles_b200.so"
    assert lib.b2_version() >= 100


def test_no_cpu_fallback_without_a_device():
    lib = _capi.load()
    if lib.b2_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from cannoles_b200.linsolve import B200Error, B200Struct
    one = np.array([1], dtype=np.int64)
    with pytest.raises(B200Error, match="no CUDA device"):
        B200Struct(1, one, one, np.array([1.0]), nvar=1, nequ=0, ncon=0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cannoles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "ldl_oracle" not in src, f
