"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/sqrn.h declares; compute entry points refuse to run without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "sqrn.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(sqrn_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    from squarna_b200 import _lib
    L = _lib.load()
    names = _declared()
    assert len(names) >= 11
    for n in names:
        assert getattr(L, n) is not None, n
    assert set(_lib.EXPORTS) <= set(names)
    assert L.sqrn_abi_version() == 3


def test_struct_layouts_match_header():
    """ctypes mirrors: field counts/sizes that would silently corrupt calls if they drifted"""
    from squarna_b200 import _abi
    assert C.sizeof(_abi.ParamSet) == 4 + 4 + 64 + 32 * 8 + 11 * 8       # n_bp + pad, keys, vals, 11 doubles
    assert C.sizeof(_abi.Batch) % 8 == 0 and C.sizeof(_abi.Result) == 16 * 8 and C.sizeof(_abi.Stems) == 5 * 8


def test_no_cpu_fallback():
    """without a CUDA device the library must fail loudly, never compute on the host"""
    from squarna_b200 import _lib
    L = _lib.load()
    if L.sqrn_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SqrnError, match="no usable CUDA device"):
        _lib.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under squarna_b200/ may reference it"""
    pkg = os.path.join(ROOT, "squarna_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, fn), encoding="utf-8").read()
                assert "import oracle" not in text and "from oracle" not in text and "sqrn_oracle" not in text, fn
