"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/lt_b200.h
declares (no compute call is made without a GPU)."""

import ctypes
import os
import re

import pytest

from lattice_based_tagger_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'lt_b200.h'), encoding='utf-8').read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lt_[a-z_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _native.load().lt_abi_version() == _native.LT_ABI_VERSION


def test_edge_record_is_16_bytes():
    assert _native.EDGE_DTYPE.itemsize == 16
    assert ctypes.sizeof(_native.lt_func) == 32


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_native.NativeLibraryError):
        _native.load()
