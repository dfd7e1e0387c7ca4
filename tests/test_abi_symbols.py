"""The C-ABI library loads on a machine without a GPU and exports every symbol include/kvpress_b200.h
declares (no compute calls here); argument validation paths that need no device are exercised too."""
import ctypes
import re
from pathlib import Path

import pytest

from kvpress_b200 import native

HEADER = Path(__file__).resolve().parent.parent / "include" / "kvpress_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(kvp_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = native.load()
    names = declared_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in native.SIGNATURES, f"{name} has no ctypes signature in native.py"
    assert sorted(native.SIGNATURES) == names


def test_status_strings_and_validation_without_gpu():
    lib = native.load()
    assert lib.kvp_abi_version() == 1
    assert lib.kvp_status_string(0) == b"ok"
    assert b"workspace" in lib.kvp_status_string(-5)
    p = native.KvpProblem()
    p.B, p.Hkv, p.Hq, p.S, p.D, p.n_kept, p.dtype = 1, 8, 32, 131072, 128, 39321, 0
    out = ctypes.c_size_t(0)
    assert lib.kvp_workspace_bytes(ctypes.byref(p), native.SCORER_KNORM, ctypes.byref(out)) == 0
    knorm_ws = out.value
    assert 2 * 8 * 131072 <= knorm_ws < 64 << 20
    assert lib.kvp_workspace_bytes(ctypes.byref(p), native.SCORER_EXPECTED_ATTENTION, ctypes.byref(out)) == 0
    assert out.value > knorm_ws
    p.D = 12
    assert lib.kvp_workspace_bytes(ctypes.byref(p), native.SCORER_KNORM, ctypes.byref(out)) == -2
    p.D, p.dtype = 128, 7
    assert lib.kvp_workspace_bytes(ctypes.byref(p), native.SCORER_KNORM, ctypes.byref(out)) == -3
    assert lib.kvp_workspace_bytes(None, native.SCORER_KNORM, ctypes.byref(out)) == -1
    n = ctypes.c_int(0)
    p.dtype = 0
    assert lib.kvp_launches_per_compress(ctypes.byref(p), native.SCORER_STREAMING, ctypes.byref(n)) == 0 and n.value == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setenv("KVPRESS_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(native.NativeLibraryError, match="no CPU or PyTorch fallback"):
        native.load()
