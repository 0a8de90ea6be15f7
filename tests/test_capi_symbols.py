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


def test_jni_shim_type_checks_against_the_header():
    """jni/com_mass_gpu_DismemberGPU.c cannot be built here (no JDK): type-check it against include/dismember_gpu.h with a
    stand-in jni.h, so that a signature drifting between the C ABI and the shim is caught (the shim forwards 1:1)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-I", os.path.join(root, "tests", "jni_stub"), "-I", os.path.join(root, "include"),
                        os.path.join(root, "jni", "com_mass_gpu_DismemberGPU.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


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


def test_jtm_assign_level_matches_the_python_mirror_of_rebalance():
    """dmg_jtm_assign_level is host code inside the C-ABI library (TreeLearning.reBalance, jtm/.../optim/TreeLearning.scala:217-265):
    it runs without a device and must pick exactly what the readable Python mirror picks, ties and overflow included."""
    import numpy as np
    from dismember_b200._capi import jtm_assign_level
    from dismember_b200.jtm import re_balance, stable_desc_order
    rng = np.random.default_rng(5)
    for n_items, n_child, max_assign, n_par in [(40, 4, 10, 1), (300, 8, 13, 3), (64, 2, 32, 2), (50, 4, 5, 2)]:
        parents = np.sort(rng.integers(3, 3 + n_par, n_items)).astype(np.int32)
        rng.shuffle(parents)
        w = rng.normal(0, 1, (n_items, n_child)).astype(np.float32)
        w[rng.random((n_items, n_child)) < 0.2] = 0.5                  # plenty of exact ties
        w[rng.random(n_items) < 0.1] = -1e6                            # items without samples (TreeLearning.scala:160)
        first = (parents.astype(np.int64) + 1) * n_child - 1
        old = (first + rng.integers(0, n_child, n_items)).astype(np.int32)
        got = jtm_assign_level(parents, old, w, max_assign)
        want = parents.copy()
        for par in np.unique(parents):
            rows = np.flatnonzero(parents == par)
            f0 = (int(par) + 1) * n_child - 1
            order = np.stack([stable_desc_order(w[i]) for i in rows])
            cn = f0 + order
            cw = np.take_along_axis(w[rows], order, 1)
            its = [int(i) for i in rows]
            balanced = re_balance(its, cn, cw, {int(i): int(old[i]) for i in rows}, [f0 + c for c in range(n_child)], max_assign)
            for node, assigned in balanced.items():
                assert len(assigned) <= max_assign
                for it in assigned:
                    want[it] = node
        assert (got == want).all()


def test_jni_shim_and_python_binding_call_only_declared_entry_points():
    """jni/com_mass_gpu_DismemberGPU.c cannot be compiled in this image (no jni.h): check statically that every dmg_*
    function it calls is declared in include/dismember_gpu.h, and that the ctypes binding gives every declared
    int32-returning entry point an argument signature."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set(dmg.declared_symbols())
    jni = open(os.path.join(root, "jni", "com_mass_gpu_DismemberGPU.c")).read()
    called = set(re.findall(r"\b(dmg_[a-z0-9_]+)\s*\(", jni))
    assert called and not (called - declared), f"JNI shim calls undeclared functions: {sorted(called - declared)}"
    L = dmg.load_library()
    unbound = [s for s in declared if getattr(L, s).argtypes is None]
    assert not unbound, f"declared but without a ctypes signature in _capi.py: {unbound}"
