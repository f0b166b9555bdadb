"""
CPU: the C-ABI shared library loads and exports every symbol include/tapqir_b200.h declares, the
ctypes table in tapqir_b200/_lib.py covers the same set, and struct layouts agree.  No compute
calls (no GPU needed).
"""

import ctypes
import re
from pathlib import Path

import pytest

from tapqir_b200 import _lib
from tapqir_b200.models import layout as L

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "tapqir_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_header_symbol():
    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = header_symbols()
    assert len(syms) >= 18
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert header_symbols() == set(_lib.SIGNATURES)


def test_struct_layouts_match():
    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built")
    lib = _lib.load()
    assert lib.tq_sizeof_model_const() == ctypes.sizeof(L.ModelConst)
    assert lib.tq_site_record_rows() == L.NSAMP * 6 + 4
    assert lib.tq_version() >= 100
    # tq_patch_view: 7 int32 (+pad) followed by 8 pointers
    assert ctypes.sizeof(_lib.PatchView) == 32 + 8 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: a missing .so raises NativeLibraryError instead of computing something else."""
    from tapqir_b200.exceptions import NativeLibraryError

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(NativeLibraryError):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch

    with pytest.raises(ValueError):
        _lib.ptr(torch.zeros(3))
