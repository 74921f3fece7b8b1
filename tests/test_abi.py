"""The C-ABI shared library loads and exports every symbol include/wfo_b200.h
declares (no compute: there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wfo_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wfo_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pywfo_b200 import build, _lib
    build.build()          # no-op when the .so is newer than its sources
    return _lib.load()


def test_header_and_binding_agree(lib):
    from pywfo_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 18
    assert sorted(_lib.SIGNATURES) == syms


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_housekeeping_without_gpu(lib):
    assert lib.wfo_version() >= 100
    assert isinstance(lib.wfo_last_error(), bytes)
    n = lib.wfo_device_count()
    assert n >= 0
    if n == 0:
        h = ctypes.c_void_p()
        rc = lib.wfo_create(0, ctypes.byref(h))
        assert rc != 0 and lib.wfo_last_error()   # fails loudly, no CPU fallback


def test_product_fails_loudly_without_gpu(lib):
    import numpy as np
    from pywfo_b200 import _lib, main
    if lib.wfo_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.WfoError):
        main.moovlp(np.eye(2), np.eye(2), np.eye(2))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pywfo_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pywfo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
