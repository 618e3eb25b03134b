"""Host packer (native, runs without a GPU): planes / okmask / exception list."""
import numpy as np

from basecount_b200 import _lib
import pytest

from basecount_b200 import synth
from basecount_b200.pack import pack_batches
from basecount_b200.records import ReadBatch, select_reads


def unpack(p, i):
    """Decode read i of a PackedBatch back to (codes, ok bits)."""
    w0, w1 = int(p.seq_woff[i]), int(p.seq_woff[i + 1])
    words = p.planes[w0:w1]
    lo = (words & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (words >> np.uint64(32)).astype(np.uint32)
    bits = np.arange(32, dtype=np.uint32)
    lo_b = ((lo[:, None] >> bits) & 1).reshape(-1)
    hi_b = ((hi[:, None] >> bits) & 1).reshape(-1)
    ok = None
    if p.okmask is not None:
        ok = ((p.okmask[w0:w1][:, None] >> bits) & 1).reshape(-1)
    return lo_b | (hi_b << 1), ok


@pytest.mark.parametrize("mbq", [0, 20])
def test_pack_roundtrip(mbq):
    b = synth.fuzz_batch(11, n_reads=120, ref_len=400)
    p = pack_batches(b, mbq)
    assert p.n_reads == b.n and p.seq_woff[-1] == p.planes.shape[0]
    code_of = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    exc = {(int(r), int(q) >> 2): int(q) & 3 for r, q in zip(p.exc_read, p.exc_pos)}
    assert list(zip(p.exc_read.tolist(), (p.exc_pos >> 2).tolist())) == sorted(exc)   # sorted by (read, pos)
    for i in range(b.n):
        s0, s1 = int(b.seq_off[i]), int(b.seq_off[i + 1])
        codes, ok = unpack(p, i)
        assert codes.shape[0] >= s1 - s0
        for j in range(s1 - s0):
            ch, q = int(b.seq[s0 + j]), int(b.qual[s0 + j])
            if ch in code_of:
                assert codes[j] == code_of[ch]
                assert (i, j) not in exc
                if mbq:
                    assert ok[j] == (q >= mbq)
            else:
                assert codes[j] == 0
                if mbq == 0:
                    assert exc[(i, j)] == (3 if ch == ord("N") else 2)
                else:
                    assert ok[j] == 0
                    assert exc.get((i, j), 0) == (1 if (ch == ord("N") and q >= mbq) else 0)
        assert not codes[s1 - s0:].any()          # padding bits are zero


def test_pack_multi_ref_and_empty():
    a = synth.fuzz_batch(1, n_reads=30)
    e = ReadBatch.from_lists([], [], [], [])
    c = synth.fuzz_batch(2, n_reads=50)
    p = pack_batches([a, e, c], 0, canonical=False)
    assert p.ref_read_off.tolist() == [0, 30, 30, 80]
    assert p.n_reads == 80 and p.cigar_off[-1] == a.cigar.shape[0] + c.cigar.shape[0]
    assert p.aligned_bases == a.aligned_bases() + c.aligned_bases()
    pc = pack_batches([a, e, c], 0)                 # the packers' normal form: never longer, same reads
    assert pc.n_reads == 80 and pc.cigar_off[-1] <= p.cigar_off[-1] and pc.cigar_off[-1] == pc.cigar.shape[0]
    assert np.array_equal(pc.planes, p.planes) and np.array_equal(pc.starts, p.starts)
    p0 = pack_batches(e, 0)
    assert p0.n_reads == 0 and p0.planes.shape[0] == 0


def _walk(ops, start=0):
    """What count.cpp:40-96 does with a CIGAR: (reference column, read position) of every counted base, and the
    deletion / skip columns."""
    r, q, bases, dels = start, 0, [], []
    for op, ln in ops:
        if op in (0, 7, 8):
            bases += [(r + k, q + k) for k in range(ln)]
            r += ln
            q += ln
        elif op == 1:
            q += ln
        elif op in (2, 3):
            dels += list(range(r, r + ln))
            r += ln
    return bases, dels


def test_canonical_cigars_mean_the_same():
    """The packers' normal form (csrc/cigar_canon.h) against a plain restatement of the reference's walk, on random
    CIGARs over all ten op codes with empty operations and long runs of equal neighbours."""
    rng = np.random.default_rng(5)
    L = _lib.lib()
    cig, off = [], [0]
    reads = []
    for _ in range(400):
        n = int(rng.integers(0, 9))
        ops = [(int(rng.integers(0, 10)), int(rng.integers(0, 6))) for _ in range(n)]
        reads.append(ops)
        cig += [(ln << 4) | op for op, ln in ops]
        off.append(len(cig))
    reads.append([(0, (1 << 28) - 1), (7, 5)])                  # a merge that would not fit 28 bits stays two ops
    cig += [(((1 << 28) - 1) << 4) | 0, (5 << 4) | 7]
    off.append(len(cig))
    cig = np.array(cig, dtype=np.uint32)
    off = np.array(off, dtype=np.uint64)
    n = len(reads)
    out_off = np.zeros(n + 1, dtype=np.uint32)
    m = int(L.bc_canonical_cigars(n, _lib.ptr(cig), _lib.ptr(off), None, _lib.ptr(out_off)))
    out = np.zeros(m, dtype=np.uint32)
    assert int(L.bc_canonical_cigars(n, _lib.ptr(cig), _lib.ptr(off), _lib.ptr(out), _lib.ptr(out_off))) == m
    assert m <= cig.shape[0] and out_off[-1] == m
    for i, ops in enumerate(reads):
        canon = [(int(w & 15), int(w >> 4)) for w in out[out_off[i]:out_off[i + 1]]]
        assert all(op in (0, 1, 2) and ln > 0 for op, ln in canon)
        if i < n - 1:
            assert all(a[0] != b[0] for a, b in zip(canon, canon[1:]))       # equal neighbours are merged
            assert _walk(canon, 7) == _walk(ops, 7)
        else:
            assert canon == [(0, (1 << 28) - 1), (0, 5)]


def test_pack_rejects_read_overrun():
    b = ReadBatch.from_lists(["ACG"], [[30, 30, 30]], [0], [[(0, 4)]])
    with pytest.raises(ValueError):
        pack_batches(b, 0)
    ok = ReadBatch.from_lists(["ACG"], [[30, 30, 30]], [0], [[(0, 3), (1, 5)]])     # trailing insertion is never read
    pack_batches(ok, 0)


def test_select_reads_trims_soft_clips():
    rec = synth.amplicon_sample(seed=3, n_reads=500, ref_len=3000)
    b = select_reads(rec, 0, 0)
    # query bases consumed by M/I/=/X equal the trimmed sequence length
    op, ln = b.cigar & 0xF, (b.cigar >> 4).astype(np.int64)
    per_op = np.where(np.isin(op, [0, 1, 7, 8]), ln, 0)
    consumed = np.add.reduceat(per_op, b.cigar_off[:-1].astype(np.int64))
    assert np.array_equal(consumed, (b.seq_off[1:] - b.seq_off[:-1]).astype(np.int64))
    assert b.n == int((((rec.flag & 4) == 0)).sum())
    assert select_reads(rec, 0, 30).n < b.n
