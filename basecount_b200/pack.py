"""Host packer: ReadBatch (ASCII bases, phred bytes) -> the device's SoA batch.

This is what replaces pybind11's by-value list -> std::vector conversion at the
`count.bcount` boundary (basecount/count.cpp:10-13): start positions, BAM-native CIGAR
words and a 2-bit, bit-planar sequence (see struct bc_batch in include/basecount_b200.h).
The packing itself runs in native code (bc_pack_reads in the C-ABI library).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .records import ReadBatch


@dataclass
class PackedBatch:
    n_reads: int
    n_refs: int
    ref_read_off: np.ndarray      # uint32[n_refs+1]
    starts: np.ndarray            # uint32[n]
    cigar_off: np.ndarray         # uint32[n+1]
    cigar: np.ndarray             # uint32[m]
    seq_woff: np.ndarray          # uint32[n+1]
    planes: np.ndarray            # uint64[W]
    okmask: np.ndarray | None     # uint32[W]
    exc_read: np.ndarray          # uint32[e]
    exc_pos: np.ndarray           # uint32[e]
    sorted_hint: bool = False
    mean_read_len: int = 0
    aligned_bases: int = 0        # BASELINE.md's unit, for throughput reporting
    query_bases: int = 0
    _struct: object = field(default=None, repr=False)

    @property
    def n_exc(self) -> int:
        return int(self.exc_read.shape[0])

    def h2d_bytes(self) -> int:
        arrs = [self.ref_read_off, self.starts, self.cigar_off, self.cigar, self.seq_woff, self.planes,
                self.exc_read, self.exc_pos]
        if self.okmask is not None:
            arrs.append(self.okmask)
        return int(sum(a.nbytes for a in arrs))

    def algorithmic_bytes(self, ref_lens) -> int:
        """BASELINE.md's compulsory-traffic figure for the counting kernel:
        n_reads*12 + n_cigar_ops*4 + ceil(query_aligned_bases/4) + sum(L)*6*4
        (+ 1 bit/base for the quality mask when min_base_quality > 0; BASELINE.md budgets 1 B/base)."""
        b = self.n_reads * 12 + int(self.cigar.shape[0]) * 4 + (self.query_bases + 3) // 4
        b += int(sum(ref_lens)) * 6 * 4
        if self.okmask is not None:
            b += (self.query_bases + 7) // 8
        return b

    def as_struct(self) -> _lib.BcBatch:
        s = _lib.BcBatch()
        s.n_reads = self.n_reads
        s.n_refs = self.n_refs
        for name in ("ref_read_off", "starts", "cigar_off", "cigar", "seq_woff", "planes", "exc_read", "exc_pos"):
            a = getattr(self, name)
            setattr(s, name, a.ctypes.data if a.size else None)
        s.ref_read_off = self.ref_read_off.ctypes.data
        s.cigar_off = self.cigar_off.ctypes.data
        s.seq_woff = self.seq_woff.ctypes.data
        s.okmask = self.okmask.ctypes.data if self.okmask is not None and self.okmask.size else None
        s.n_exc = self.n_exc
        s.on_device = 0
        s.sorted_hint = int(self.sorted_hint)
        s.mean_read_len = int(self.mean_read_len)
        s.reserved = 0
        self._struct = s
        return s


def _alloc(n, dtype, pinned):
    return _lib.pinned_empty(n, dtype) if pinned else np.empty(int(n), dtype=dtype)


def pack_batches(batches, min_base_quality: int = 0, pinned: bool = False, canonical: bool = True) -> PackedBatch:
    """Pack one ReadBatch per reference slot (in slot order) into a single device batch.

    canonical: hand the device the CIGARs' normal form (csrc/cigar_canon.h: =/X spelled M, N spelled D, ignored
    and empty operations dropped, equal neighbours merged -- the same counts by count.cpp:51,74,80,92).  False
    keeps the BAM-native words as they are (tests: the general kernel takes those reads)."""
    L = _lib.lib()
    if isinstance(batches, ReadBatch):
        batches = [batches]
    n_refs = len(batches)
    if min_base_quality < 0:
        raise TypeError("min_base_quality must be unsigned")       # count.cpp:9 takes unsigned int
    counts = [b.n for b in batches]
    n = int(sum(counts))
    if n >= 2 ** 32:
        raise TypeError("more than 2^32-1 reads in one batch")
    ref_read_off = _alloc(n_refs + 1, np.uint32, pinned)
    ref_read_off[0] = 0
    np.cumsum(counts, out=ref_read_off[1:])

    if n_refs == 1:
        b = batches[0]
        starts_src, cigar_src, seq, qual = b.starts, b.cigar, b.seq, b.qual
        cigar_off64 = b.cigar_off.astype(np.uint64, copy=False)
        seq_off = np.ascontiguousarray(b.seq_off, dtype=np.uint64)
    else:
        starts_src = np.concatenate([b.starts for b in batches]) if n else np.zeros(0, np.uint32)
        cigar_src = np.concatenate([b.cigar for b in batches]) if n else np.zeros(0, np.uint32)
        seq = np.concatenate([b.seq for b in batches]) if n else np.zeros(0, np.uint8)
        qual = np.concatenate([b.qual for b in batches]) if n else np.zeros(0, np.uint8)
        co, so, cb, sb = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], 0, 0
        for b in batches:
            co.append(b.cigar_off[1:].astype(np.uint64) + np.uint64(cb))
            so.append(b.seq_off[1:].astype(np.uint64) + np.uint64(sb))
            cb += int(b.cigar_off[-1])
            sb += int(b.seq_off[-1])
        cigar_off64 = np.concatenate(co)
        seq_off = np.concatenate(so)
    if int(cigar_off64[-1]) >= 2 ** 32:
        raise TypeError("more than 2^32-1 CIGAR operations in one batch")
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    no_qual = min_base_quality == 0 and qual.shape[0] == 0          # nothing reads qualities at threshold 0 (count.cpp:56)
    if qual.shape[0] != seq.shape[0] and not no_qual:
        raise TypeError("qualities and reads differ in length")

    starts = _alloc(n, np.uint32, pinned)
    starts[:] = starts_src
    cigar_off = _alloc(n + 1, np.uint32, pinned)
    if canonical:
        cigar_src = np.ascontiguousarray(cigar_src, dtype=np.uint32)
        cigar_off64 = np.ascontiguousarray(cigar_off64, dtype=np.uint64)
        m = int(L.bc_canonical_cigars(n, _lib.ptr(cigar_src), _lib.ptr(cigar_off64), None, _lib.ptr(cigar_off)))
        if m >= 2 ** 32:
            raise TypeError("bc_canonical_cigars failed")
        cigar = _alloc(m, np.uint32, pinned)
        L.bc_canonical_cigars(n, _lib.ptr(cigar_src), _lib.ptr(cigar_off64), _lib.ptr(cigar), _lib.ptr(cigar_off))
    else:
        cigar_off[:] = cigar_off64
        cigar = _alloc(cigar_src.shape[0], np.uint32, pinned)
        cigar[:] = cigar_src

    n_words = int(L.bc_pack_words(n, _lib.ptr(seq_off)))
    if n_words >= 2 ** 32:
        raise TypeError("batch sequence does not fit 2^32 words")
    seq_woff = _alloc(n + 1, np.uint32, pinned)
    planes = _alloc(n_words, np.uint64, pinned)
    okmask = _alloc(n_words, np.uint32, pinned) if min_base_quality > 0 else None
    exc_cap = max(1024, seq.shape[0] // 256)
    n_exc = ctypes.c_uint32(0)
    while True:
        exc_read = _alloc(exc_cap, np.uint32, pinned)
        exc_pos = _alloc(exc_cap, np.uint32, pinned)
        rc = L.bc_pack_reads(n, _lib.ptr(seq), None if no_qual else _lib.ptr(qual), _lib.ptr(seq_off), _lib.ptr(cigar),
                             _lib.ptr(cigar_off),
                             int(min_base_quality), _lib.ptr(seq_woff), _lib.ptr(planes), _lib.ptr(okmask),
                             _lib.ptr(exc_read), _lib.ptr(exc_pos), exc_cap, ctypes.byref(n_exc))
        if rc == _lib.BC_ERR_READ_OVERRUN:
            raise ValueError("a CIGAR consumes more bases than its read holds")
        if rc != _lib.BC_OK:
            raise TypeError(f"bc_pack_reads failed with status {rc}")
        if n_exc.value <= exc_cap:
            break
        exc_cap = n_exc.value
    ne = n_exc.value
    query_bases = int(seq.shape[0])
    sorted_hint = all(bool(np.all(b.starts[1:] >= b.starts[:-1])) for b in batches)
    aligned = int(sum(b.aligned_bases() for b in batches))
    return PackedBatch(n, n_refs, ref_read_off, starts, cigar_off, cigar, seq_woff, planes, okmask,
                       exc_read[:ne], exc_pos[:ne], sorted_hint, (query_bases // n) if n else 0, aligned, query_bases)
