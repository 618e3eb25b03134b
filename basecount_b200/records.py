"""Flat (structure-of-arrays) containers for alignment records and read batches.

`Records` holds what a BAM alignment record carries for this path (the fields the
reference reads through pysam at basecount/main.py:165-173).  `ReadBatch` is what
the reference hands to `count.bcount` for ONE reference sequence
(basecount/main.py:146-153) -- soft clips already trimmed, unmapped / low-MAPQ
reads already dropped -- but flat instead of four Python lists.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# BAM CIGAR op codes (SAM spec section 4.2): M I D N S H P = X B
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X, OP_B = range(10)
FLAG_UNMAPPED = 0x4


@dataclass
class Records:
    ref_names: list
    ref_lengths: list
    ref_id: np.ndarray      # int32[n], -1 when the record has no reference
    pos: np.ndarray         # int32[n], 0-based leftmost position
    mapq: np.ndarray        # uint8[n]
    flag: np.ndarray        # uint16[n]
    cigar: np.ndarray       # uint32[m]  BAM-native  len << 4 | op   (soft clips included)
    cigar_off: np.ndarray   # int64[n+1]
    seq: np.ndarray         # uint8[t]   ASCII, full query (soft-clipped bases included)
    qual: np.ndarray        # uint8[t]   phred values (not +33)
    seq_off: np.ndarray     # int64[n+1]
    names: list = field(default_factory=list)   # optional read names (BAM writer)

    @property
    def n(self) -> int:
        return int(self.ref_id.shape[0])


@dataclass
class ReadBatch:
    """Inputs of one `bcount` call, flat.  seq/qual are query_alignment_* (trimmed)."""
    starts: np.ndarray      # uint32[n]
    cigar: np.ndarray       # uint32[m]
    cigar_off: np.ndarray   # uint64[n+1]
    seq: np.ndarray         # uint8[t] ASCII
    qual: np.ndarray        # uint8[t]
    seq_off: np.ndarray     # uint64[n+1]

    @property
    def n(self) -> int:
        return int(self.starts.shape[0])

    def aligned_bases(self) -> int:
        """BASELINE.md's unit: sum of opLen over ops in {M,=,X,D,N}."""
        op = self.cigar & 0xF
        keep = (op == OP_M) | (op == OP_EQ) | (op == OP_X) | (op == OP_D) | (op == OP_N)
        return int((self.cigar[keep] >> 4).astype(np.int64).sum())

    def to_lists(self):
        """The four Python lists `count.bcount` takes (basecount/count.cpp:10-13)."""
        seq_b = self.seq.tobytes()
        so = self.seq_off.tolist()
        co = self.cigar_off.tolist()
        ops = (self.cigar & 0xF).tolist()
        lens = (self.cigar >> 4).tolist()
        ql = self.qual.tolist()
        reads = [seq_b[so[i]:so[i + 1]].decode("ascii") for i in range(self.n)]
        quals = [ql[so[i]:so[i + 1]] for i in range(self.n)]
        ctuples = [list(zip(ops[co[i]:co[i + 1]], lens[co[i]:co[i + 1]])) for i in range(self.n)]
        return reads, quals, self.starts.tolist(), ctuples

    @staticmethod
    def from_lists(reads, qualities, starts, ctuples) -> "ReadBatch":
        n = len(reads)
        if not (len(qualities) == len(starts) == len(ctuples) == n):
            raise TypeError("reads, qualities, starts and ctuples must have equal lengths")
        seq_len = np.fromiter((len(r) for r in reads), dtype=np.int64, count=n)
        seq_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(seq_len, out=seq_off[1:])
        seq = np.frombuffer("".join(reads).encode("latin-1"), dtype=np.uint8)
        total = int(seq_off[-1])
        qual = np.empty(total, dtype=np.uint8)
        p = 0
        for q in qualities:
            m = len(q)
            qual[p:p + m] = q
            p += m
        if p != total:
            raise TypeError("qualities and reads differ in total length")
        n_ops = np.fromiter((len(c) for c in ctuples), dtype=np.int64, count=n)
        cigar_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(n_ops, out=cigar_off[1:])
        flat = [t for c in ctuples for t in c]
        if flat:
            arr = np.asarray(flat, dtype=np.int64).reshape(-1, 2)
            if (arr < 0).any():
                raise TypeError("negative CIGAR op or length")
            cigar = ((arr[:, 1] << 4) | arr[:, 0]).astype(np.uint32)
        else:
            cigar = np.zeros(0, dtype=np.uint32)
        st = np.asarray(starts, dtype=np.int64)
        if st.size and (st < 0).any():
            raise TypeError("negative start")
        return ReadBatch(st.astype(np.uint32), cigar, cigar_off, seq, qual, seq_off)


def _leading_trailing_clips(rec: Records):
    """Per record: (soft-clipped bases at the left, at the right) -- what pysam's
    query_alignment_start / query_alignment_end trim (hard clips hold no bases)."""
    n = rec.n
    lead = np.zeros(n, dtype=np.int64)
    trail = np.zeros(n, dtype=np.int64)
    co = rec.cigar_off
    nops = co[1:] - co[:-1]
    has = nops > 0
    if not has.any():
        return lead, trail
    op = (rec.cigar & 0xF).astype(np.int64)
    ln = (rec.cigar >> 4).astype(np.int64)
    first = co[:-1].copy()
    last = co[1:] - 1
    idx = np.flatnonzero(has)
    f = first[idx]
    l = last[idx]
    # optional hard clip outside the soft clip
    f_is_h = op[f] == OP_H
    f2 = np.where(f_is_h & (f + 1 <= l), f + 1, f)
    lead[idx] = np.where(op[f2] == OP_S, ln[f2], 0)
    l_is_h = op[l] == OP_H
    l2 = np.where(l_is_h & (l - 1 >= f), l - 1, l)
    tr = np.where(op[l2] == OP_S, ln[l2], 0)
    # a single soft-clip op must not be counted on both sides
    tr = np.where((l2 == f2) & (op[f2] == OP_S), 0, tr)
    trail[idx] = tr
    return lead, trail


def select_reads(rec: Records, ref_id: int, min_mapping_quality: int = 0) -> ReadBatch:
    """Apply the reference's read filter and soft-clip trimming for one reference.

    Filter: `not is_unmapped and mapping_quality >= min_mapping_quality`
    (basecount/main.py:165) and `reference_name == ref` (main.py:166).
    """
    keep = ((rec.flag & FLAG_UNMAPPED) == 0) & (rec.mapq >= min_mapping_quality) & (rec.ref_id == ref_id)
    idx = np.flatnonzero(keep)
    # QUAL '*' is stored as 0xFF bytes: pysam returns None for query_alignment_qualities and the reference's bcount
    # raises TypeError (pybind11 cannot cast None at count.cpp:11), whatever min_base_quality is
    first = rec.seq_off[:-1][idx]
    has = rec.seq_off[1:][idx] > first
    if has.any() and bool((rec.qual[first[has]] == 0xFF).any()):
        raise TypeError("a read has no base qualities (QUAL '*'): the reference's bcount raises TypeError on None")
    lead, trail = _leading_trailing_clips(rec)
    s0 = rec.seq_off[:-1][idx] + lead[idx]
    s1 = rec.seq_off[1:][idx] - trail[idx]
    s1 = np.maximum(s1, s0)
    lens = s1 - s0
    seq_off = np.zeros(idx.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=seq_off[1:])
    total = int(seq_off[-1])
    # gather the trimmed bases
    if total:
        base = np.repeat(s0 - seq_off[:-1].astype(np.int64), lens)
        src = base + np.arange(total, dtype=np.int64)
        seq = rec.seq[src]
        qual = rec.qual[src]
    else:
        seq = np.zeros(0, dtype=np.uint8)
        qual = np.zeros(0, dtype=np.uint8)
    c0 = rec.cigar_off[:-1][idx]
    c1 = rec.cigar_off[1:][idx]
    clen = c1 - c0
    cigar_off = np.zeros(idx.size + 1, dtype=np.uint64)
    np.cumsum(clen, out=cigar_off[1:])
    ctot = int(cigar_off[-1])
    if ctot:
        cbase = np.repeat(c0 - cigar_off[:-1].astype(np.int64), clen)
        cigar = rec.cigar[cbase + np.arange(ctot, dtype=np.int64)]
    else:
        cigar = np.zeros(0, dtype=np.uint32)
    return ReadBatch(rec.pos[idx].astype(np.uint32), cigar.astype(np.uint32), cigar_off,
                     np.ascontiguousarray(seq), np.ascontiguousarray(qual), seq_off)
