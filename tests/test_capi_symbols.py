"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/dismember_gpu.h declares; without a GPU the product fails loudly (no fallback)."""
import ctypes
import os

import pytest

import dismember_b200 as dmg
from dismember_b200 import _capi


def test_library_is_built_in_tree():
    L = dmg.load_library()
    assert os.path.dirname(_capi.LIB_PATH).endswith("dismember_b200")
    assert b"sm_100a" in L.dmg_version()


def test_every_declared_symbol_is_exported():
    syms = dmg.declared_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(_capi.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/dismember_gpu.h but not exported: {missing}"


def test_no_extra_public_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("dmg_")}
    assert exported == set(dmg.declared_symbols())


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under dismember_b200/ may import, include or dlopen it."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "dismember_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#include\s*[<\"][^>\"]*(oracle|orc_)[^>\"]*[>\"])|liboracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not bad.search(text), f"{f} reaches into oracle/"


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dmg.DmgError) as ei:
        dmg.Engine(0)
    assert "no CPU fallback" in str(ei.value)
