"""The C-ABI library loads here (no GPU) and exports every symbol include/vip_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("vb_gram_f32", "vb_eigh_f64", "vb_pcs_f32", "vb_project_subtract_f32", "vb_derotate_f32",
                 "vb_collapse_f32", "vb_last_error", "vb_version"):
        assert must in syms


def test_library_exports_all_declared_symbols():
    from vip_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for sym in _declared_symbols():
        assert hasattr(handle, sym), f"{sym} declared in include/vip_b200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(_cabi.SIGNATURES) == _declared_symbols()
    lib = _cabi.lib()
    assert lib.vb_version() >= 1000
    assert lib.vb_gram_workspace_bytes(500, 262144) > 500 * 500 * 8
    # packed real planes T1[(S+1) x N], T2[S x N] + 2N per-frame scalars, fp32 (csrc/derotate.cu)
    assert lib.vb_derotate_scratch_bytes(2, 512, 2048, 0) == 2 * ((513 + 512) * 2048 + 2 * 2048) * 4
    # generic (direct) path: complex planes
    assert lib.vb_derotate_scratch_bytes(2, 101, 402, 0) == 2 * (102 + 101) * 402 * 8


def test_missing_library_fails_loudly(monkeypatch):
    from vip_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libvipb200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _cabi.lib()
