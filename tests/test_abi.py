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


def _header_arities(path, prefix):
    """{symbol: number of parameters} parsed from the C declarations of a header (comments stripped)."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(%s[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;" % prefix, text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        if "(*" in m.group(0).split(name)[0][-8:]:
            continue
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        out[name] = n
    return out


def test_binding_arity_matches_the_headers():
    """a ctypes prototype with the wrong number of arguments corrupts the call silently: every prototype of the two
    bindings (kernel ABI and graph handle API) has exactly as many parameters as the C declaration"""
    import neuronika_b200._lib as L
    from neuronika_b200 import variable as V
    want = _header_arities(os.path.join(ROOT, "include", "nk_b200.h"), "nk_")
    assert len(want) >= 50
    bad = [(s, len(L._PROTOS[s][1]), want[s]) for s in want if s in L._PROTOS and len(L._PROTOS[s][1]) != want[s]]
    assert not bad, bad
    gwant = _header_arities(os.path.join(ROOT, "include", "nk_graph.h"), "nkg_")
    gwant = {k: v for k, v in gwant.items() if k in V._G}      # typedef'd callback types are not functions
    assert len(gwant) >= 30
    bad = [(s, len(V._G[s][1]), gwant[s]) for s in gwant if len(V._G[s][1]) != gwant[s]]
    assert not bad, bad
