"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; the ctypes
binding declares exactly those; without a GPU the product fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from stripenn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "stripenn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snn_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES.keys()) == header_symbols()


def test_version_and_struct_layout():
    lib = _lib.load()
    assert lib.snn_version() == 100
    assert _lib.CANDIDATE_DTYPE.itemsize == ctypes.sizeof(_lib.Candidate) == 88


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stripenn_b200.engine import Device
    with pytest.raises(_lib.StripennB200Error) as e:
        Device(0)
    assert e.value.code == _lib.ENODEV
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stripenn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "snn_oracle" not in src and "refharness" not in src, f
