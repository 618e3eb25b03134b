"""The drop-in boundary: every function include/basecount_b200.h declares is exported by the built library
and bound in basecount_b200/_lib.py; without a CUDA device the product path fails loudly instead of falling
back to anything.  CPU only (loads the library, launches nothing)."""
import ctypes
import os
import re

import pytest

from basecount_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "basecount_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                  # comments mention function names too
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(re.findall(r"\b(bc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("bc_create", "bc_destroy", "bc_begin", "bc_push_batch", "bc_sync", "bc_counts", "bc_stats", "bc_summary",
                 "bc_amplicons", "bc_halo_export", "bc_halo_add", "bc_pack_reads", "bc_bam_open", "bc_bam_open_region",
                 "bc_format_tsv"):
        assert must in names
    assert len(names) >= 40


def test_library_exports_every_declared_symbol():
    dll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(dll, n)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    unbound = [n for n in _declared() if n not in _lib._SIGNATURES]
    assert not unbound, unbound
    stale = [n for n in _lib._SIGNATURES if n not in _declared()]
    assert not stale, stale


def test_no_device_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.lib()
    assert L.bc_device_count() == 0
    h = ctypes.c_void_p()
    rc = L.bc_create(0, ctypes.byref(h))
    assert rc != _lib.BC_OK and not h.value
    from basecount_b200.engine import Engine
    with pytest.raises(Exception):
        Engine(0)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "basecount_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
