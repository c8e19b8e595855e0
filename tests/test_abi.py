"""The C-ABI shared library loads and exports every symbol include/wavtok_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from wavtokenizer_b200 import _native

HEADER = os.path.join(_native.ROOT, "include", "wavtok_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wt_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def library():
    _native.build()
    return ctypes.CDLL(_native.LIB_PATH)


def test_header_and_binding_agree():
    assert header_functions() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol(library):
    for name in header_functions():
        assert hasattr(library, name), name


def test_version_and_error_strings(library):
    library.wt_version.restype = ctypes.c_char_p
    assert b"sm_100a" in library.wt_version()
    lib = _native.lib()
    assert lib.wt_frames_for(None, 100) == -1
    assert lib.wt_launch_count(None) == -1


def test_create_rejects_bad_arguments_without_touching_a_gpu():
    lib = _native.lib()
    out = ctypes.c_void_p()
    status = lib.wt_create(None, None, 0, 0, ctypes.byref(out))
    assert status == _native.WT_ERR_VALUE and out.value is None
    assert b"null" in lib.wt_last_error()
    cfg = _native.WtConfig()
    arr = (_native.WtTensor * 1)()
    status = lib.wt_create(ctypes.byref(cfg), arr, 0, 0, ctypes.byref(out))
    assert status == _native.WT_ERR_VALUE  # n_filters / dimension check comes first
    with pytest.raises(ValueError):
        _native.check(status)


def test_config_struct_layout_matches_header():
    assert ctypes.sizeof(_native.WtConfig) == 15 * 4
    assert ctypes.sizeof(_native.WtTensor) == 24
