"""BAI index build and index-aware region fetch.  CPU only: host code of the C-ABI library.

Nothing runnable pins this boundary to htslib (no pysam / samtools in the image), so the native builder is
checked byte for byte against an independent restatement of SAM spec section 5.2 written here with plain
Python loops, and the region reader against the whole-file decoder filtered by start position.
"""
import struct

import numpy as np
import pytest

from basecount_b200 import bamio, synth
from basecount_b200.records import Records, select_reads


# ----------------------------------------------------------------------------- spec restatement (test-side)
def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _spec_bai(bam_path):
    data = open(bam_path, "rb").read()
    blocks, u = [], 0                                    # (file offset, uncompressed start, isize)
    off = 0
    for c0, c1, isize in bamio._bgzf_blocks(data):
        xlen = data[off + 10] | (data[off + 11] << 8)
        assert c0 == off + 12 + xlen
        blocks.append((off, u, isize))
        u += isize
        off = c1 + 8
    raw = bamio.bgzf_decompress(data)
    names, lengths, p = bamio._parse_header(raw)

    def voff(x):
        for f, u0, isize in blocks:
            if u0 <= x < u0 + isize:
                return (f << 16) | (x - u0)
        return len(data) << 16

    refs = [dict(bins={}, lin={}, beg=None, end=0, nm=0, nu=0) for _ in names]
    n_no_coor, last = 0, (None, None)
    while p < len(raw):
        sz, rid, pos = struct.unpack_from("<iii", raw, p)
        l_name = raw[p + 12]
        n_cig, flag = struct.unpack_from("<HH", raw, p + 16)
        v0, v1 = voff(p), voff(p + 4 + sz)
        cig = struct.unpack_from(f"<{n_cig}I", raw, p + 36 + l_name)
        p += 4 + sz
        if rid < 0 or pos < 0:
            n_no_coor += rid < 0
            last = (None, None)
            continue
        span = 0 if flag & 4 else sum(w >> 4 for w in cig if (w & 15) in (0, 2, 3, 7, 8))
        span = span or 1
        b = _reg2bin(pos, pos + span)
        r = refs[rid]
        chunks = r["bins"].setdefault(b, [])
        if last == (rid, b) and chunks:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        for w in range(pos >> 14, ((pos + span - 1) >> 14) + 1):
            r["lin"].setdefault(w, v0)
        r["beg"] = v0 if r["beg"] is None else min(r["beg"], v0)
        r["end"] = max(r["end"], v1)
        r["nu" if flag & 4 else "nm"] += 1
        last = (rid, b)
    out = [b"BAI\x01", struct.pack("<i", len(refs))]
    for r in refs:
        any_rec = r["beg"] is not None
        out.append(struct.pack("<i", len(r["bins"]) + any_rec))
        for b in sorted(r["bins"]):
            out.append(struct.pack("<Ii", b, len(r["bins"][b])))
            out += [struct.pack("<QQ", *c) for c in r["bins"][b]]
        if any_rec:
            out.append(struct.pack("<IiQQQQ", 37450, 2, r["beg"], r["end"], r["nm"], r["nu"]))
        n_intv = max(r["lin"]) + 1 if r["lin"] else 0
        lin, nxt = [0] * n_intv, 0
        for w in range(n_intv - 1, -1, -1):
            nxt = r["lin"].get(w, nxt)
            lin[w] = nxt
        out.append(struct.pack("<i", n_intv))
        out += [struct.pack("<Q", v) for v in lin]
    out.append(struct.pack("<Q", n_no_coor))
    return b"".join(out)


def _concat(recs, names, lengths):
    """Records of several single-reference samples as one multi-reference file (ids in order)."""
    rid = [np.where(r.ref_id >= 0, i, -1).astype(np.int32) for i, r in enumerate(recs)]
    coff, soff, c0, s0 = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], 0, 0
    for r in recs:
        coff.append(r.cigar_off[1:].astype(np.int64) + c0)
        soff.append(r.seq_off[1:].astype(np.int64) + s0)
        c0 += int(r.cigar_off[-1])
        s0 += int(r.seq_off[-1])
    cat = lambda f: np.concatenate([getattr(r, f) for r in recs])
    return Records(names, lengths, np.concatenate(rid), cat("pos"), cat("mapq"), cat("flag"), cat("cigar"),
                   np.concatenate(coff), cat("seq"), cat("qual"), np.concatenate(soff))


def _same_batch(a, b):
    for f in ("starts", "cigar", "cigar_off", "seq", "qual", "seq_off"):
        assert np.array_equal(np.asarray(getattr(a, f), dtype=np.int64), np.asarray(getattr(b, f), dtype=np.int64)), f


def _expect(rec, rid, beg, end, mapq):
    keep = np.flatnonzero((rec.ref_id == rid) & (rec.pos >= beg) & (rec.pos < end))
    return select_reads(synth.take_records(rec, keep), rid, mapq)


@pytest.fixture(scope="module")
def two_ref_bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("bai")
    a = synth.uniform_short_read_sample(seed=11, ref_len=90_000, n_reads=6000, read_len=150, ref_name="chrA")
    b = synth.amplicon_sample(seed=12, n_reads=3000, ref_len=40_000, ref_name="chrB")
    rec = _concat([a, b], ["chrA", "chrB"], [90_000, 40_000])
    p = str(d / "two.bam")
    bamio.write_bam(p, rec)
    return p, rec


def test_index_matches_spec_restatement(two_ref_bam):
    p, _ = two_ref_bam
    bai = bamio.write_bai(p)
    assert open(bai, "rb").read() == _spec_bai(p)


@pytest.mark.parametrize("threads", [1, 0])
def test_region_fetch_equals_filtered_whole_file(two_ref_bam, threads):
    p, rec = two_ref_bam
    bamio.write_bai(p)
    regions = [(0, 0, 90_000), (0, 0, 1), (0, 16_383, 16_385), (0, 16_384, 32_768), (0, 30_000, 61_234), (0, 89_000, 90_000),
               (0, 89_990, 200_000), (1, 0, 40_000), (1, 5_000, 5_001), (1, 12_345, 33_333), (1, 39_999, 40_000)]
    for rid, beg, end in regions:
        for mapq in (0, 30):
            nb = bamio.NativeBam(p, threads=threads, region=(rid, beg, end))
            assert nb.ref_names == ["chrA", "chrB"] and nb.ref_lengths == [90_000, 40_000]
            got = nb.select(rid, mapq)
            _same_batch(got, _expect(rec, rid, beg, end, mapq))
            rid_all, pos_all, _, _ = nb.core()                       # nothing outside the region was kept
            assert ((rid_all == rid) & (pos_all >= beg) & (pos_all < end)).all()
            nb.close()


def test_region_shards_tile_the_reference(two_ref_bam):
    """The np.linspace split of tests/test_basecount.py:146-150: the shards' reads are a partition."""
    p, rec = two_ref_bam
    bamio.write_bai(p)
    for world in (2, 3, 8):
        bounds = np.linspace(0, 90_000, num=world + 1, dtype=np.int64)
        n = 0
        for r in range(world):
            nb = bamio.NativeBam(p, region=(0, int(bounds[r]), int(bounds[r + 1])))
            n += nb.select(0, 0).n
            nb.close()
        assert n == select_reads(rec, 0, 0).n


def test_long_skips_force_the_walk_to_extend(tmp_path):
    """A read that starts early but spans many 16 kbp windows (an N skip) drags the linear index of far
    windows back to the file start, so the first guess of where the region ends is far too short."""
    rng = np.random.default_rng(3)
    n = 4000
    pos = np.sort(rng.integers(0, 60_000, n)).astype(np.int32)
    pos[0] = min(100, int(pos[1]))
    cigar, coff = [], [0]
    for i in range(n):
        cigar += [(10 << 4) | 0, (70_000 << 4) | 3, (10 << 4) | 0] if i == 0 else [(20 << 4) | 0]
        coff.append(len(cigar))
    soff = np.arange(n + 1, dtype=np.int64) * 20
    rec = Records(["r"], [200_000], np.zeros(n, np.int32), pos, np.full(n, 60, np.uint8), np.zeros(n, np.uint16),
                  np.asarray(cigar, np.uint32), np.asarray(coff, np.int64),
                  np.frombuffer(b"ACGT" * (5 * n), dtype=np.uint8).copy(), np.full(20 * n, 30, np.uint8), soff)
    p = str(tmp_path / "skip.bam")
    bamio.write_bam(p, rec)
    bai = bamio.write_bai(p)
    assert open(bai, "rb").read() == _spec_bai(p)
    for beg, end in [(0, 20_000), (0, 60_000), (15_000, 45_000), (59_000, 60_000), (100, 101)]:
        nb = bamio.NativeBam(p, region=(0, beg, end))
        _same_batch(nb.select(0, 0), _expect(rec, 0, beg, end, 0))
        nb.close()


def test_empty_reference_and_errors(tmp_path):
    a = synth.uniform_short_read_sample(seed=4, ref_len=5000, n_reads=300, read_len=50, ref_name="a")
    rec = _concat([a], ["a", "empty"], [5000, 7000])
    p = str(tmp_path / "e.bam")
    bamio.write_bam(p, rec)
    bamio.write_bai(p)
    nb = bamio.NativeBam(p, region=(1, 0, 7000))
    assert nb.n == 0 and nb.select(1, 0).n == 0
    nb.close()
    nb = bamio.NativeBam(p, region=(0, 3000, 3000))                  # empty interval
    assert nb.n == 0
    nb.close()
    with pytest.raises(ValueError):
        bamio.NativeBam(p, region=(2, 0, 10))                        # no such reference
    with pytest.raises(ValueError):
        bamio.NativeBam(p, region=(0, 0, 10), index=str(tmp_path / "missing.bai"))
    # a BAM that is not coordinate-sorted cannot be indexed
    idx = np.arange(a.n)[::-1].copy()
    q = str(tmp_path / "unsorted.bam")
    bamio.write_bam(q, synth.take_records(a, idx))
    with pytest.raises(ValueError):
        bamio.write_bai(q)
    # an index of another file is refused
    with pytest.raises(ValueError):
        bamio.NativeBam(q, region=(0, 0, 10), index=p + ".bai")
