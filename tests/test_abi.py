"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/gs_b200.h
declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("gs_ctx_create", "gs_spline_create", "gs_grid1d_step", "gs_grid2d_iteration", "gs_grid2d_xiter",
                 "gs_grid2d_yiter", "gs_grid2d_update", "gs_predef"):
        assert must in syms


def test_library_exports_every_declared_symbol(gs):
    lib = ctypes.CDLL(gs.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_every_declared_symbol(gs):
    from gridsplines_b200 import _lib

    bound = set(_lib.SIGNATURES) | set(_lib._SPECIAL)
    assert set(declared_symbols()) == bound


def test_version(gs):
    from gridsplines_b200 import _lib

    assert _lib.lib().gs_version() == 100


def test_no_cpu_fallback(gs):
    """Without a device the product path raises; it never computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gs.GsError) as e:
        gs.Context(0)
    assert e.value.status == -2  # GS_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gridsplines_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gs_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
