"""Deterministic synthetic alignments for the five BASELINE.json configs.

Shapes follow SURVEY.md section 8(d): position-sorted reads, MAPQ 60 except 2 % at
0-29, 1 % unmapped records interleaved, base qualities uniform 2-40, bases copied
from a fixed random reference with 0.5 % substitutions and 0.1 % N, CIGAR mix
88 % pure M / 5 % one 1-3 bp D / 3 % one 1-3 bp I / 3 % soft clips 5-30 bp /
1 % =,X spelling.  No network, no real BAMs: everything comes from
numpy.random.default_rng(seed).
"""
from __future__ import annotations

import numpy as np

from .records import (FLAG_UNMAPPED, OP_D, OP_EQ, OP_I, OP_M, OP_S, OP_X, ReadBatch, Records)

SARS2_NAME = "MN908947.3"
SARS2_LEN = 29903
CHR20_NAME = "chr20"
CHR20_LEN = 64444167
N_AMPLICONS = 98
AMPLICON_LEN = 400

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_reference(length: int, seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def amplicon_starts(ref_len: int = SARS2_LEN, n: int = N_AMPLICONS, amp_len: int = AMPLICON_LEN) -> np.ndarray:
    """Left ends of n tiled amplicons (~amp_len bp, ~25 % overlap), ARTIC-v3-shaped."""
    lo, hi = 30, ref_len - amp_len - 30
    return np.round(np.linspace(lo, hi, n)).astype(np.int64)


def artic_like_bed(path: str, ref_name: str = SARS2_NAME, ref_len: int = SARS2_LEN, seed: int = 11,
                   n: int = N_AMPLICONS, amp_len: int = AMPLICON_LEN) -> None:
    """Write a primer BED of the shape load_scheme expects (basecount/scheme.py:6-9):
    whitespace-separated, column 4 = SCHEME_TILE_SIDE[_altN]."""
    rng = np.random.default_rng(seed)
    st = amplicon_starts(ref_len, n, amp_len)
    lines = []
    for i, a in enumerate(st):
        pl = int(rng.integers(22, 31))
        pr = int(rng.integers(22, 31))
        pool = 1 + (i % 2)
        lines.append(f"{ref_name}\t{a}\t{a + pl}\tnCoV-2019_{i + 1}_LEFT\t{pool}\t+")
        if i % 17 == 3:   # a few alternate primers, as ARTIC v3 has
            lines.append(f"{ref_name}\t{a - 4}\t{a + pl - 2}\tnCoV-2019_{i + 1}_LEFT_alt{i % 5}\t{pool}\t+")
        lines.append(f"{ref_name}\t{a + amp_len - pr}\t{a + amp_len}\tnCoV-2019_{i + 1}_RIGHT\t{pool}\t-")
        if i % 23 == 7:
            lines.append(f"{ref_name}\t{a + amp_len - pr + 3}\t{a + amp_len + 5}\tnCoV-2019_{i + 1}_RIGHT_alt{i % 3}\t{pool}\t-")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def _build(ref: np.ndarray, ref_name: str, starts: np.ndarray, spans: np.ndarray,
           rng: np.random.Generator, unmapped_frac: float = 0.01) -> Records:
    """Turn (start, reference span) pairs into full records with the CIGAR mix above."""
    n = starts.size
    L = ref.size
    order = np.argsort(starts, kind="stable")
    starts = starts[order].astype(np.int64)
    spans = spans[order].astype(np.int64)

    u = rng.random(n)
    kind = np.zeros(n, dtype=np.int8)            # 0 M | 1 D | 2 I | 3 soft clips | 4 =/X
    kind[u >= 0.88] = 1
    kind[u >= 0.93] = 2
    kind[u >= 0.96] = 3
    kind[u >= 0.99] = 4
    kind[spans < 12] = 0                        # too short to carry an indel cleanly

    indel = rng.integers(1, 4, size=n)          # 1..3 bp
    cut = (spans * rng.uniform(0.15, 0.85, size=n)).astype(np.int64)
    cut = np.clip(cut, 4, np.maximum(spans - 8, 4))
    clipL = rng.integers(5, 31, size=n)
    clipR = rng.integers(5, 31, size=n)
    side = rng.integers(0, 3, size=n)           # 0 left, 1 right, 2 both

    # up to 5 pieces per read: [S] [M1] [I|D|X] [M2] [S]; q = query bases, r = ref bases
    P = 5
    op = np.zeros((n, P), dtype=np.int64)
    ln = np.zeros((n, P), dtype=np.int64)
    isM, isD, isI, isS, isE = (kind == k for k in range(5))

    op[:, 1] = OP_M
    ln[:, 1] = spans
    # deletion: M1=cut, D=indel, M2=span-cut-indel
    op[isD, 2] = OP_D
    ln[isD, 1] = cut[isD]
    ln[isD, 2] = indel[isD]
    op[isD, 3] = OP_M
    ln[isD, 3] = spans[isD] - cut[isD] - indel[isD]
    # insertion: M1=cut, I=indel, M2=span-cut
    op[isI, 2] = OP_I
    ln[isI, 1] = cut[isI]
    ln[isI, 2] = indel[isI]
    op[isI, 3] = OP_M
    ln[isI, 3] = spans[isI] - cut[isI]
    # soft clips around a full match
    sl = isS & (side != 1)
    sr = isS & (side != 0)
    op[sl, 0] = OP_S
    ln[sl, 0] = clipL[sl]
    op[sr, 4] = OP_S
    ln[sr, 4] = clipR[sr]
    # '=' / 'X' spelling: cut '=' , 1 'X', rest '='
    op[isE, 1] = OP_EQ
    ln[isE, 1] = cut[isE]
    op[isE, 2] = OP_X
    ln[isE, 2] = 1
    op[isE, 3] = OP_EQ
    ln[isE, 3] = spans[isE] - cut[isE] - 1

    present = ln > 0
    consumes_q = present & ((op == OP_M) | (op == OP_I) | (op == OP_S) | (op == OP_EQ) | (op == OP_X))
    consumes_r = present & ((op == OP_M) | (op == OP_D) | (op == OP_EQ) | (op == OP_X))
    qlen_piece = np.where(consumes_q, ln, 0)
    rlen_piece = np.where(consumes_r, ln, 0)
    r_begin = starts[:, None] + np.cumsum(rlen_piece, axis=1) - rlen_piece
    from_ref = consumes_q & consumes_r           # bases copied from the reference

    nops = present.sum(axis=1)
    cigar_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nops, out=cigar_off[1:])
    cigar = ((ln[present] << 4) | op[present]).astype(np.uint32)

    qlen = qlen_piece.sum(axis=1)
    seq_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(qlen, out=seq_off[1:])
    total = int(seq_off[-1])

    fq = qlen_piece.reshape(-1)
    piece_of_base = np.repeat(np.arange(n * P, dtype=np.int64), fq)
    piece_q_off = np.cumsum(fq) - fq
    within = np.arange(total, dtype=np.int64) - piece_q_off[piece_of_base]
    rb = r_begin.reshape(-1)[piece_of_base] + within
    fr = from_ref.reshape(-1)[piece_of_base]
    seq = _ACGT[rng.integers(0, 4, size=total, dtype=np.uint8)]
    ok = fr & (rb < L)
    seq[ok] = ref[rb[ok]]
    v = rng.random(total)
    sub = v < 0.005
    seq[sub] = _ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    seq[v >= 0.999] = ord("N")
    qual = rng.integers(2, 41, size=total, dtype=np.uint8)

    mapq = np.full(n, 60, dtype=np.uint8)
    low = rng.random(n) < 0.02
    mapq[low] = rng.integers(0, 30, size=int(low.sum()), dtype=np.uint8)
    flag = np.zeros(n, dtype=np.uint16)
    ref_id = np.zeros(n, dtype=np.int32)
    pos = starts.astype(np.int32)

    rec = Records([ref_name], [int(L)], ref_id, pos, mapq, flag, cigar, cigar_off, seq, qual, seq_off)
    n_un = int(round(n * unmapped_frac))
    if n_un:
        rec = _interleave_unmapped(rec, n_un, rng)
    return rec


def _interleave_unmapped(rec: Records, n_un: int, rng: np.random.Generator) -> Records:
    """Insert n_un unmapped records (flag 4, no reference, no CIGAR, 100 random bases)."""
    n = rec.n
    where = np.sort(rng.integers(0, n + 1, size=n_un))
    is_un = np.zeros(n + n_un, dtype=bool)
    is_un[where + np.arange(n_un)] = True
    m = n + n_un
    ref_id = np.full(m, -1, dtype=np.int32)
    pos = np.full(m, -1, dtype=np.int32)
    mapq = np.zeros(m, dtype=np.uint8)
    flag = np.full(m, FLAG_UNMAPPED, dtype=np.uint16)
    ref_id[~is_un] = rec.ref_id
    pos[~is_un] = rec.pos
    mapq[~is_un] = rec.mapq
    flag[~is_un] = rec.flag
    nops = np.zeros(m, dtype=np.int64)
    nops[~is_un] = rec.cigar_off[1:] - rec.cigar_off[:-1]
    cigar_off = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(nops, out=cigar_off[1:])
    qlen = np.full(m, 100, dtype=np.int64)
    qlen[~is_un] = rec.seq_off[1:] - rec.seq_off[:-1]
    seq_off = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(qlen, out=seq_off[1:])
    total = int(seq_off[-1])
    base_is_un = np.repeat(is_un, qlen)
    seq = np.empty(total, dtype=np.uint8)
    qual = np.empty(total, dtype=np.uint8)
    seq[~base_is_un] = rec.seq
    qual[~base_is_un] = rec.qual
    k = int(base_is_un.sum())
    seq[base_is_un] = _ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    qual[base_is_un] = rng.integers(2, 41, size=k, dtype=np.uint8)
    return Records(rec.ref_names, rec.ref_lengths, ref_id, pos, mapq, flag, rec.cigar, cigar_off, seq, qual, seq_off)


def amplicon_sample(seed: int = 1, n_reads: int = 124_000, ref_len: int = SARS2_LEN,
                    ref_name: str = SARS2_NAME, ref_seed: int = 7) -> Records:
    """Config 1 / 2 (and one sample of config 4): every read spans one of 98 amplicons,
    start jittered +-5, uniform over amplicons (~1,650x depth at 124k reads)."""
    rng = np.random.default_rng(seed)
    ref = random_reference(ref_len, ref_seed)
    n_amp = N_AMPLICONS if ref_len >= 4 * AMPLICON_LEN else 1
    amp_len = min(AMPLICON_LEN, max(ref_len - 60, 8))
    st = amplicon_starts(ref_len, n_amp, amp_len) if n_amp > 1 else np.array([min(30, ref_len // 4)])
    amp = rng.integers(0, st.size, size=n_reads)
    starts = np.clip(st[amp] + rng.integers(-5, 6, size=n_reads), 0, None)
    spans = np.minimum(np.full(n_reads, amp_len, dtype=np.int64), ref_len - starts)
    return _build(ref, ref_name, starts, spans, rng)


def deep_short_read_sample(seed: int = 3, n_reads: int = 2_000_000, read_len: int = 150,
                           ref_len: int = SARS2_LEN, ref_name: str = SARS2_NAME, ref_seed: int = 7) -> Records:
    """Config 3: 150 bp reads placed inside the amplicons (~10,000x at 2 M reads)."""
    rng = np.random.default_rng(seed)
    ref = random_reference(ref_len, ref_seed)
    st = amplicon_starts(ref_len)
    amp = rng.integers(0, st.size, size=n_reads)
    starts = st[amp] + rng.integers(0, AMPLICON_LEN - read_len + 1, size=n_reads)
    spans = np.minimum(np.full(n_reads, read_len, dtype=np.int64), ref_len - starts)
    return _build(ref, ref_name, starts, spans, rng)


def uniform_short_read_sample(seed: int = 5, ref_len: int = CHR20_LEN, n_reads: int = 12_888_833,
                              read_len: int = 150, ref_name: str = CHR20_NAME, ref_seed: int = 9,
                              start_lo: int = 0, start_hi: int | None = None) -> Records:
    """Config 5 (or a region of it): 150 bp reads at uniform starts (30x at full size)."""
    rng = np.random.default_rng(seed)
    ref = random_reference(ref_len, ref_seed)
    hi = (ref_len - read_len) if start_hi is None else start_hi
    starts = rng.integers(start_lo, max(hi, start_lo + 1), size=n_reads)
    spans = np.minimum(np.full(n_reads, read_len, dtype=np.int64), ref_len - starts)
    return _build(ref, ref_name, starts, spans, rng)


def fuzz_batch(seed: int, n_reads: int = 200, ref_len: int = 500, max_ops: int = 9,
               sorted_by_pos: bool = False, allow_overflow: bool = False, long_op_frac: float = 0.1) -> ReadBatch:
    """Adversarial small `bcount` inputs (already in the trimmed domain the operator sees):
    arbitrary CIGARs over all ten op codes, zero-length ops, IUPAC / lower-case letters,
    empty reads and CIGARs, surplus trailing bases, optionally alignments past ref_len."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMSWBDHV*=", dtype=np.uint8)
    starts, cig, cig_off, seqs, quals, seq_off = [], [], [0], [], [], [0]
    for _ in range(n_reads):
        k = int(rng.integers(0, max_ops + 1))
        ops = rng.choice(10, size=k, p=[.42, .1, .1, .04, .08, .04, .03, .08, .08, .03])
        lens = rng.integers(0, 40, size=k)
        if k and rng.random() < long_op_frac:
            lens[int(rng.integers(0, k))] = int(rng.integers(100, 300))
        span = int(sum(int(l) for o, l in zip(ops, lens) if o in (0, 2, 3, 7, 8)))
        if not allow_overflow and span > ref_len:
            ops, lens, k, span = ops[:0], lens[:0], 0, 0
        p = int(rng.integers(0, ref_len)) if allow_overflow else int(rng.integers(0, ref_len - span + 1))
        qn = int(sum(int(l) for o, l in zip(ops, lens) if o in (0, 1, 7, 8)))
        qn += int(rng.integers(0, 3)) if rng.random() < 0.2 else 0      # surplus bases are legal
        starts.append(p)
        cig.extend(((int(l) << 4) | int(o)) for o, l in zip(ops, lens))
        cig_off.append(len(cig))
        seqs.append(alphabet[rng.integers(0, alphabet.size, size=qn)] if rng.random() < 0.3
                    else _ACGT[rng.integers(0, 4, size=qn)])
        quals.append(rng.integers(0, 61, size=qn).astype(np.uint8))
        seq_off.append(seq_off[-1] + qn)
    b = ReadBatch(np.asarray(starts, dtype=np.uint32), np.asarray(cig, dtype=np.uint32),
                  np.asarray(cig_off, dtype=np.uint64),
                  np.concatenate(seqs) if seqs else np.zeros(0, np.uint8),
                  np.concatenate(quals) if quals else np.zeros(0, np.uint8),
                  np.asarray(seq_off, dtype=np.uint64))
    if sorted_by_pos:
        b = take_batch(b, np.argsort(b.starts, kind="stable"))
    return b


def take_batch(b: ReadBatch, idx: np.ndarray) -> ReadBatch:
    """Reorder / subset the reads of a batch."""
    idx = np.asarray(idx, dtype=np.int64)
    so = b.seq_off.astype(np.int64)
    co = b.cigar_off.astype(np.int64)
    qlen = (so[1:] - so[:-1])[idx]
    seq_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(qlen, out=seq_off[1:])
    src = np.repeat(so[:-1][idx] - seq_off[:-1], qlen) + np.arange(int(seq_off[-1]), dtype=np.int64)
    nops = (co[1:] - co[:-1])[idx]
    cigar_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(nops, out=cigar_off[1:])
    csrc = np.repeat(co[:-1][idx] - cigar_off[:-1], nops) + np.arange(int(cigar_off[-1]), dtype=np.int64)
    return ReadBatch(b.starts[idx], b.cigar[csrc], cigar_off.astype(np.uint64), b.seq[src], b.qual[src],
                     seq_off.astype(np.uint64))


def concat_batches(batches) -> ReadBatch:
    so, co, s_base, c_base = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], 0, 0
    for b in batches:
        so.append(b.seq_off[1:] + np.uint64(s_base))
        co.append(b.cigar_off[1:] + np.uint64(c_base))
        s_base += int(b.seq_off[-1])
        c_base += int(b.cigar_off[-1])
    return ReadBatch(np.concatenate([b.starts for b in batches]), np.concatenate([b.cigar for b in batches]),
                     np.concatenate(co), np.concatenate([b.seq for b in batches]),
                     np.concatenate([b.qual for b in batches]), np.concatenate(so))


def take_records(rec: Records, idx: np.ndarray) -> Records:
    """Reorder / subset records (used to unsort inputs in tests)."""
    idx = np.asarray(idx, dtype=np.int64)
    qlen = (rec.seq_off[1:] - rec.seq_off[:-1])[idx]
    seq_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(qlen, out=seq_off[1:])
    tot = int(seq_off[-1])
    src = np.repeat(rec.seq_off[:-1][idx] - seq_off[:-1], qlen) + np.arange(tot, dtype=np.int64)
    nops = (rec.cigar_off[1:] - rec.cigar_off[:-1])[idx]
    cigar_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(nops, out=cigar_off[1:])
    ctot = int(cigar_off[-1])
    csrc = np.repeat(rec.cigar_off[:-1][idx] - cigar_off[:-1], nops) + np.arange(ctot, dtype=np.int64)
    return Records(rec.ref_names, rec.ref_lengths, rec.ref_id[idx], rec.pos[idx], rec.mapq[idx], rec.flag[idx],
                   rec.cigar[csrc], cigar_off, rec.seq[src], rec.qual[src], seq_off)
