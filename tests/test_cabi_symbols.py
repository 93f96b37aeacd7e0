"""CPU: the C-ABI library loads and exports every symbol include/mdacp_b200.h declares; the
product path has no CPU fallback (compute entries fail loudly without a device)."""
import ctypes
import os
import re

import pytest

from mdacp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdacp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert L.mdb200_abi_version() == 1


def test_no_cpu_fallback():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    P = capi.Params()
    for d in range(3):
        P.L[d] = 20.0
        P.grid_size[d] = 1
    P.time_step, P.cutoff, P.buffer_length = 0.001, 3.0, 0.3
    with pytest.raises(capi.MDB200Error) as ei:
        capi.Unit(P, 0)
    assert ei.value.code == capi.ECUDA


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or run oracle/"""
    pkg = os.path.join(ROOT, "mdacp_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|pyoracle|oracle/|mdacp_oracle)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not pat.search(src), (dp, f)
