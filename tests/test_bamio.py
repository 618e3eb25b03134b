"""BGZF / BAM writer <-> reader round trip (SAM spec section 4).  CPU only."""
import struct
import zlib

import numpy as np

from basecount_b200 import bamio, synth
from basecount_b200.records import select_reads


def _same(a, b):
    assert a.ref_names == b.ref_names and a.ref_lengths == b.ref_lengths
    for f in ("ref_id", "pos", "mapq", "flag", "cigar", "cigar_off", "seq", "qual", "seq_off"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_bam_roundtrip(tmp_path):
    rec = synth.amplicon_sample(seed=5, n_reads=2500, ref_len=4000, ref_name="chrT")
    p = str(tmp_path / "t.bam")
    bamio.write_bam(p, rec)
    back = bamio.read_bam(p)
    _same(rec, back)
    # an unmapped record keeps refID -1 / pos -1 and has no CIGAR
    un = np.flatnonzero(back.flag & 4)
    assert un.size and (back.ref_id[un] == -1).all() and ((back.cigar_off[1:] - back.cigar_off[:-1])[un] == 0).all()


def test_bgzf_container_is_standard_gzip(tmp_path):
    """Every BGZF block is an RFC1952 member: Python's gzip/zlib must read the file as-is."""
    rec = synth.deep_short_read_sample(seed=2, n_reads=3000, ref_len=3000, ref_name="x")
    p = str(tmp_path / "t.bam")
    bamio.write_bam(p, rec)
    data = open(p, "rb").read()
    assert data.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # EOF marker
    raw, d, rest = b"", zlib.decompressobj(31), data
    while rest:
        raw += d.decompress(rest)
        rest = d.unused_data
        d = zlib.decompressobj(31)
    assert raw == bamio.encode_bam_bytes(rec) == bamio.bgzf_decompress(data)
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    assert b"@SQ\tSN:x\tLN:3000" in raw[8:8 + l_text]


def test_odd_lengths_empty_and_multi_reference(tmp_path):
    a = synth.uniform_short_read_sample(seed=1, ref_len=900, n_reads=40, read_len=33, ref_name="a")
    b = synth.uniform_short_read_sample(seed=2, ref_len=500, n_reads=25, read_len=1, ref_name="b")
    rec = synth.take_records(a, np.arange(a.n))
    # append b's records under reference id 1
    from basecount_b200.records import Records
    rid = np.where(b.ref_id >= 0, 1, -1).astype(np.int32)
    rec = Records(["a", "b"], [900, 500], np.concatenate([a.ref_id, rid]), np.concatenate([a.pos, b.pos]),
                  np.concatenate([a.mapq, b.mapq]), np.concatenate([a.flag, b.flag]),
                  np.concatenate([a.cigar, b.cigar]), np.concatenate([a.cigar_off, b.cigar_off[1:] + a.cigar_off[-1]]),
                  np.concatenate([a.seq, b.seq]), np.concatenate([a.qual, b.qual]),
                  np.concatenate([a.seq_off, b.seq_off[1:] + a.seq_off[-1]]))
    p = str(tmp_path / "m.bam")
    bamio.write_bam(p, rec)
    _same(rec, bamio.read_bam(p))
    empty = synth.take_records(rec, np.zeros(0, dtype=np.int64))
    bamio.write_bam(p, empty)
    back = bamio.read_bam(p)
    assert back.n == 0 and back.ref_names == ["a", "b"]


def test_pysam_shaped_surface(tmp_path):
    rec = synth.amplicon_sample(seed=6, n_reads=300, ref_len=3000, ref_name="toy")
    p = str(tmp_path / "t.bam")
    bamio.write_bam(p, rec)
    f = bamio.AlignmentFile(p, mode="rb")
    assert f.references == ("toy",) and f.lengths == (3000,)
    reads = list(f.fetch(until_eof=True))
    assert len(reads) == rec.n
    kept = [r for r in reads if not r.is_unmapped and r.mapping_quality >= 30]
    b = select_reads(rec, 0, 30)
    assert len(kept) == b.n
    seqs, quals, starts, ctuples = b.to_lists()
    for i in (0, 1, len(kept) // 2, len(kept) - 1):
        assert kept[i].query_alignment_sequence == seqs[i]
        assert list(kept[i].query_alignment_qualities) == quals[i]
        assert kept[i].reference_start == starts[i] and kept[i].cigartuples == ctuples[i]
        assert kept[i].reference_name == "toy"
    f.close()
