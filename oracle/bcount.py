"""oracle/bcount.py -- TEST INFRASTRUCTURE ONLY (see oracle/bcount_oracle.c header).

ctypes front-end to the C restatement of the reference operator
(reference: basecount/count.cpp:7-99) and, when present, a loader for the compiled
unmodified reference `oracle/_ref/count*.so`.  Parity status: pinned, see
tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> None:
    """Compile the C restatement (and _ref when /root/reference is mounted)."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbcount_oracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.bcount_oracle.restype = ctypes.c_int
        lib.bcount_oracle.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64] + [ctypes.c_void_p] * 7
        lib.bcount_oracle_aligned_bases.restype = ctypes.c_uint64
        lib.bcount_oracle_aligned_bases.argtypes = [ctypes.c_uint64, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def bcount_flat(ref_len: int, min_base_quality: int, batch, counts: np.ndarray | None = None) -> np.ndarray:
    """Run the C oracle on a flat ReadBatch; returns (ref_len, 6) uint32.

    Raises IndexError where the reference raises it (alignment increments past ref_len).
    """
    if counts is None:
        counts = np.zeros((ref_len, 6), dtype=np.uint32)
    seq = np.ascontiguousarray(batch.seq, dtype=np.uint8)
    qual = np.ascontiguousarray(batch.qual, dtype=np.uint8)
    seq_off = np.ascontiguousarray(batch.seq_off, dtype=np.uint64)
    starts = np.ascontiguousarray(batch.starts, dtype=np.uint32)
    cigar = np.ascontiguousarray(batch.cigar, dtype=np.uint32)
    cigar_off = np.ascontiguousarray(batch.cigar_off, dtype=np.uint64)
    rc = _lib().bcount_oracle(ref_len, min_base_quality, starts.shape[0], _p(seq), _p(qual), _p(seq_off),
                              _p(starts), _p(cigar), _p(cigar_off), _p(counts))
    if rc == 1:
        raise IndexError("alignment extends past the end of the reference")
    if rc == 2:
        raise ValueError("CIGAR consumes more bases than the read holds")
    return counts


def aligned_bases(cigar: np.ndarray) -> int:
    cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
    return int(_lib().bcount_oracle_aligned_bases(cigar.shape[0], _p(cigar)))


def load_ref_bcount():
    """The compiled, unmodified reference operator (oracle/_ref), or None if not built."""
    hits = glob.glob(os.path.join(_HERE, "_ref", "count*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("count", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.bcount
