"""Drop-in for the reference's native module `count` (basecount/count.cpp:102-105).

    bcount(refLen, minBaseQuality, reads, qualities, starts, ctuples) -> list[list[int]]

Same positional signature, same refLen x 6 result (columns A,C,G,T,DS,N), same
exceptions (IndexError past the reference end, TypeError for arguments pybind11 would
refuse) -- computed by the sm_100a kernels behind the C ABI.  No CPU path.
"""
from __future__ import annotations

import numbers

from .engine import Engine
from .pack import pack_batches
from .records import ReadBatch

_ENGINE = None


def default_engine() -> Engine:
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(0)
    return _ENGINE


def _unsigned(x, name):
    if isinstance(x, bool) or not isinstance(x, numbers.Integral) or x < 0 or x >= 2 ** 32:
        raise TypeError(f"bcount(): incompatible argument {name}={x!r} (unsigned int expected)")
    return int(x)


def bcount(refLen, minBaseQuality, reads, qualities, starts, ctuples):
    ref_len = _unsigned(refLen, "refLen")
    mbq = _unsigned(minBaseQuality, "minBaseQuality")
    if any(not isinstance(r, str) for r in reads):
        raise TypeError("bcount(): reads must be a list of str")           # e.g. SEQ '*' -> None (pybind11 TypeError)
    if any(q is None for q in qualities) or any(c is None for c in ctuples):
        raise TypeError("bcount(): qualities / ctuples must not contain None")
    batch = ReadBatch.from_lists(reads, qualities, starts, ctuples)
    eng = default_engine()
    eng.begin([ref_len])
    eng.push(pack_batches(batch, mbq))
    eng.sync()
    return eng.counts(0).tolist()
