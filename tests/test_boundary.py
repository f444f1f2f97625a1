"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads
without a GPU, exports every symbol include/bsls_b200.h declares, fails loudly (no CPU
fallback) when no device is present, and the product package never touches oracle/."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "block-simplex-least-squares_b200")


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as g
    return g.build()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "bsls_b200.h")) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"^BSLS_API\s+[\w\s\*]+?\b(bsls_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for name in ("bsls_proj_simplex", "bsls_proj_multi_simplex", "bsls_proj_multi_ball"):
        assert name in syms


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = set(re.findall(r" T (bsls_\w+)", out))
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())


def test_library_is_sm100a_native(libpath):
    sass = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    code = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    assert "UBLKCP" in code, "bulk (TMA) copies missing from SASS"


def test_no_cpu_fallback_without_device(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bsls_b200
    with pytest.raises(bsls_b200._lib.BslsError):
        bsls_b200.proj_multi_simplex_c(np.zeros(4), np.array([0, 2]))


def test_asserts_match_reference_without_device():
    """Argument errors raise AssertionError before any device work (c_extensions.pyx:24,33-34)."""
    import bsls_b200
    y = np.random.rand(7)
    for start, end in [(2, 8), (-1, 7), (-1, 4)]:
        with pytest.raises(AssertionError):
            bsls_b200.proj_simplex_c(y, start, end)
    for b in [np.array([0, 4, 2])]:
        with pytest.raises(AssertionError):
            bsls_b200.proj_multi_simplex_c(y, b)
    bsls_b200.proj_simplex_c(y, 4, 4)  # empty range: silent no-op as in the reference


def test_product_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(PKG):
        if os.path.basename(root) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(root, f)) as fh:
                    src = fh.read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "liboracle" in src or "bsls_oracle" in src:
                    bad.append(f)
    assert not bad, bad
