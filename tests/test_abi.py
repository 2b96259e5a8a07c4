"""CPU checks of the C-ABI boundary: the library loads and exports every symbol the header
declares, with no compute call (there is no GPU here)."""
import ctypes
import os
import re

from uglad_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "uglad_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uglad_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 15
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_load_and_pure_host_queries():
    lib = _lib.load()
    assert lib.uglad_abi_version() == 1
    assert lib.uglad_param_count(3) == 42  # 1 + 28 (rho_l1) + 13 (lambda_f)
    d = _lib.UgladDims(4, 100, 15, 3, 0, 4, 0, 1.0)
    n = lib.uglad_workspace_floats(ctypes.byref(d))
    assert n > 3 * 15 * 4 * 100 * 100
    off = lib.uglad_workspace_offset(ctypes.byref(d), b"theta")
    assert 0 < off < n
    assert lib.uglad_workspace_offset(ctypes.byref(d), b"nope") == ctypes.c_size_t(-1).value
    bad = _lib.UgladDims(4, 100, 15, 99, 0, 4, 0, 1.0)
    assert lib.uglad_workspace_floats(ctypes.byref(bad)) == 0
    assert b"H=99" in lib.uglad_last_error()


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.UgladDims) == 32
