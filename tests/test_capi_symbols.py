"""CPU-side checks of the C-ABI boundary: the library loads and exports exactly what
include/craft_b200.h declares, and the ctypes mirror (craft_b200/_lib.py) agrees with it."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "craft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(craft_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _header_functions()
    assert "craft_shift_gemm" in names and "craft_corr_build" in names and "craft_attn_pv" in names
    assert len(names) >= 19


def test_library_exports_every_declared_symbol():
    from craft_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()
    for name in _header_functions():
        assert hasattr(lib, name), "libcraft_b200.so does not export %s" % name
    assert lib.craft_b200_abi_version() == _lib.ABI_VERSION


def test_ctypes_mirror_matches_header():
    from craft_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_missing_device_fails_loudly():
    """No silent CPU fallback: a CPU tensor must raise, not compute."""
    import torch
    from craft_b200 import _lib, ops
    grid = ops.TokenGrid(4, 4)
    with pytest.raises((_lib.CraftB200Error, AssertionError)):
        ops.pack_tokens(torch.zeros((128, 4, 4)), grid, out_f=torch.zeros((grid.Mp, 128)))
