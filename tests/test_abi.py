"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/nk_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import neuronika_b200._lib as L
    syms = header_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(L.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # the Python binding declares a prototype for every header symbol (and nothing else)
    assert sorted(L.exported_symbols()) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import neuronika_b200 as nk
    with pytest.raises(nk.NkError, match="no CPU fallback"):
        nk.Device(0)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under neuronika_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "neuronika_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
