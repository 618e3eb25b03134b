"""BGZF / BAM writer <-> reader round trip (SAM spec section 4).  CPU only."""
import struct
import zlib

import numpy as np
import pytest

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


# ----------------------------------------------------------------------------- native decoder (C-ABI library)
def _same_batch(a, b):
    for f in ("starts", "cigar", "cigar_off", "seq", "qual", "seq_off"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def _clip_torture_records():
    """Records that exercise the soft-clip trimming rule (pysam query_alignment_*): H outside S on
    either side, a CIGAR that is a single S, no CIGAR at all, missing SEQ, odd lengths, IUPAC letters,
    unmapped records with a reference id, two references.  (Missing QUAL makes the reference raise:
    test_missing_qualities_raise_type_error.)"""
    from basecount_b200.records import Records
    rng = np.random.default_rng(17)
    cigs = [
        [(5, 3), (4, 5), (0, 20), (4, 2), (5, 7)], [(4, 7), (0, 13)], [(0, 9), (4, 4)], [(4, 6)], [(5, 2), (4, 6)],
        [], [(0, 11), (1, 2), (0, 3), (2, 4), (0, 5)], [(5, 4), (0, 7), (5, 1)], [(4, 1), (4, 2), (0, 5)],
        [(0, 1)], [(3, 50), (0, 6)], [(4, 3), (7, 4), (8, 1), (7, 2), (4, 3)],
    ]
    ref_id, pos, mapq, flag, cigar, coff, seq, qual, soff = [], [], [], [], [], [0], [], [], [0]
    letters = np.frombuffer(b"ACGTNRYKMSWBDHV=", dtype=np.uint8)
    for rep in range(40):
        for k, ct in enumerate(cigs):
            qn = sum(l for o, l in ct if o in (0, 1, 4, 7, 8))
            if k == 5:
                qn = 6                                   # bases but no CIGAR
            if rep % 7 == 3 and k == 1:
                qn = 0                                   # SEQ '*'
            ref_id.append((rep + k) % 2 if (rep * 13 + k) % 11 else -1)
            pos.append(int(rng.integers(0, 400)))
            mapq.append(int(rng.integers(0, 61)))
            flag.append(4 if (rep + k) % 9 == 0 else (16 if k % 2 else 0))
            cigar += [(l << 4) | o for o, l in ct]
            coff.append(len(cigar))
            seq.append(letters[rng.integers(0, letters.size, size=qn)])
            qual.append(rng.integers(0, 61, size=qn).astype(np.uint8))
            soff.append(soff[-1] + qn)
    return Records(["r0", "r1"], [1000, 700], np.asarray(ref_id, np.int32), np.asarray(pos, np.int32),
                   np.asarray(mapq, np.uint8), np.asarray(flag, np.uint16), np.asarray(cigar, np.uint32),
                   np.asarray(coff, np.int64), np.concatenate(seq), np.concatenate(qual), np.asarray(soff, np.int64))


def test_native_decoder_matches_python_decoder(tmp_path):
    """csrc/bam_decode.h against the numpy decoder (itself pinned to the SAM spec by the round trips
    above): header, per-record core fields, and selections (filter + soft-clip trimming) over
    whole files and record ranges, for every thread count."""
    cases = [synth.amplicon_sample(seed=5, n_reads=2500, ref_len=4000, ref_name="chrT"), _clip_torture_records(),
             synth.take_records(synth.amplicon_sample(seed=5, n_reads=50, ref_len=4000, ref_name="chrT"),
                                np.zeros(0, dtype=np.int64))]
    for ci, rec in enumerate(cases):
        p = str(tmp_path / f"n{ci}.bam")
        bamio.write_bam(p, rec)
        py = bamio.read_bam(p)
        for threads in (1, 3, 0):
            nb = bamio.NativeBam(p, threads)
            assert nb.n == py.n and nb.ref_names == py.ref_names and nb.ref_lengths == py.ref_lengths
            ref_id, pos, mapq, flag = nb.core()
            assert np.array_equal(ref_id, py.ref_id) and np.array_equal(pos, py.pos)
            assert np.array_equal(mapq, py.mapq) and np.array_equal(flag, py.flag)
            for rid in range(len(py.ref_names)):
                for mmq in (0, 30, 60):
                    _same_batch(nb.select(rid, mmq), select_reads(py, rid, mmq))
            if py.n > 10:
                a, b = py.n // 3, 2 * py.n // 3 + 1
                from basecount_b200.main import _slice_records
                _same_batch(nb.select(0, 20, a, b), select_reads(_slice_records(py, a, b), 0, 20))
            nb.close()


def test_native_decoder_rejects_damaged_files(tmp_path):
    import pytest
    rec = synth.amplicon_sample(seed=8, n_reads=400, ref_len=3000, ref_name="x")
    p = str(tmp_path / "ok.bam")
    bamio.write_bam(p, rec)
    data = bytearray(open(p, "rb").read())
    bad = str(tmp_path / "bad.bam")
    for damage in ("flip", "truncate", "notbgzf"):
        d = bytearray(data)
        if damage == "flip":
            d[len(d) // 2] ^= 0x55                       # payload corruption -> inflate / CRC failure
        elif damage == "truncate":
            d = d[:len(d) // 2]
        else:
            d = bytearray(b"plain text, not a BAM file")
        open(bad, "wb").write(bytes(d))
        with pytest.raises(ValueError):
            bamio.NativeBam(bad)
    with pytest.raises(ValueError):
        bamio.NativeBam(str(tmp_path / "missing.bam"))


def _same_packed(a, b):
    assert a.n_reads == b.n_reads and a.n_refs == b.n_refs == 1
    for f in ("ref_read_off", "starts", "cigar_off", "cigar", "seq_woff", "planes", "exc_read", "exc_pos"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (a.okmask is None) == (b.okmask is None)
    if a.okmask is not None:
        assert np.array_equal(a.okmask, b.okmask)
    assert (a.sorted_hint, a.mean_read_len, a.aligned_bases, a.query_bases) == \
           (b.sorted_hint, b.mean_read_len, b.aligned_bases, b.query_bases)


def test_one_pass_native_pack_matches_select_then_pack(tmp_path):
    """NativeBam.pack (4-bit bases straight to the 2-bit planes, one native pass) against
    pack_batches(NativeBam.select(...)): identical arrays, including the sparse exception list, the
    quality mask, odd-length soft clips (the packed bytes are then read at an odd nibble), IUPAC letters,
    and the error for a CIGAR that consumes more bases than the read holds."""
    from basecount_b200.pack import pack_batches
    from basecount_b200.records import Records
    rec = synth.amplicon_sample(seed=9, n_reads=3000, ref_len=4000, ref_name="chrT")
    # a copy of the torture set without the records whose CIGAR overruns the read (those are checked below)
    tort = _clip_torture_records()
    cases = [rec, synth.uniform_short_read_sample(seed=3, ref_len=3000, n_reads=800, read_len=33, ref_name="u")]
    for ci, r in enumerate(cases):
        p = str(tmp_path / f"p{ci}.bam")
        bamio.write_bam(p, r)
        for threads in (1, 0):
            nb = bamio.NativeBam(p, threads)
            for mbq in (0, 20, 41):
                for mmq in (0, 30):
                    _same_packed(nb.pack(0, mmq, mbq), pack_batches([nb.select(0, mmq)], mbq))
            a, b = r.n // 4, 3 * r.n // 4
            _same_packed(nb.pack(0, 0, 0, a, b), pack_batches([nb.select(0, 0, a, b, want_qual=False)], 0))
            nb.close()
    p = str(tmp_path / "tort.bam")
    bamio.write_bam(p, tort)
    nb = bamio.NativeBam(p)
    for rid in (0, 1):
        for mbq in (0, 15):
            try:
                want = pack_batches([nb.select(rid, 0)], mbq)
            except ValueError:
                with pytest.raises(ValueError):
                    nb.pack(rid, 0, mbq)
                continue
            _same_packed(nb.pack(rid, 0, mbq), want)
    # record ranges of the torture file that hold no overrunning CIGAR must agree exactly
    checked = 0
    for a in range(0, nb.n - 12, 12):
        try:
            want = pack_batches([nb.select(0, 0, a, a + 12)], 7)
        except ValueError:
            continue
        _same_packed(nb.pack(0, 0, 7, a, a + 12), want)
        checked += 1
    assert checked > 5
    nb.close()


# ----------------------------------------------------------------------------- random records (hypothesis)
def _random_records(draw):
    from hypothesis import strategies as st
    from basecount_b200.records import Records
    letters = np.frombuffer(b"ACGTNRYKMSWBDHV=", dtype=np.uint8)
    n = draw(st.integers(0, 12))
    ref_id, pos, mapq, flag, cigar, coff, seq, qual, soff = [], [], [], [], [], [0], [], [], [0]
    for _ in range(n):
        lead = draw(st.lists(st.tuples(st.sampled_from([4, 5]), st.integers(0, 9)), max_size=2))
        body = draw(st.lists(st.tuples(st.sampled_from([0, 1, 2, 3, 6, 7, 8]), st.integers(0, 35)), max_size=5))
        tail = draw(st.lists(st.tuples(st.sampled_from([4, 5]), st.integers(0, 9)), max_size=2))
        ct = lead + body + tail
        qn = sum(l for o, l in ct if o in (0, 1, 4, 7, 8))
        if draw(st.integers(0, 9)) == 0:
            qn = 0                                        # SEQ '*'
        mapped = draw(st.integers(0, 7)) != 0
        ref_id.append(draw(st.integers(0, 1)) if mapped or draw(st.booleans()) else -1)
        pos.append(draw(st.integers(0, 600)))
        mapq.append(draw(st.integers(0, 60)))
        flag.append(0 if mapped else 4)
        cigar += [(l << 4) | o for o, l in ct]
        coff.append(len(cigar))
        idx = draw(st.lists(st.integers(0, letters.size - 1), min_size=qn, max_size=qn))
        seq.append(letters[np.asarray(idx, dtype=np.int64)] if qn else np.zeros(0, np.uint8))
        if qn and draw(st.integers(0, 5)) == 0:
            qual.append(np.full(qn, 0xFF, np.uint8))      # QUAL '*'
        else:
            qual.append(np.asarray(draw(st.lists(st.integers(0, 60), min_size=qn, max_size=qn)), dtype=np.uint8))
        soff.append(soff[-1] + qn)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
    return Records(["r0", "r1"], [1000, 700], np.asarray(ref_id, np.int32), np.asarray(pos, np.int32),
                   np.asarray(mapq, np.uint8), np.asarray(flag, np.uint16), np.asarray(cigar, np.uint32),
                   np.asarray(coff, np.int64), cat(seq), cat(qual), np.asarray(soff, np.int64))


def test_random_records_roundtrip_and_native_selection(tmp_path):
    """Arbitrary clip layouts, op mixes, missing SEQ / QUAL, unmapped records: writer -> numpy reader gives the
    records back, and the native decoder's filter + soft-clip trimming agrees with the numpy path."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    p = str(tmp_path / "h.bam")

    @settings(max_examples=80, deadline=None, derandomize=True,
              suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large, HealthCheck.function_scoped_fixture])
    @given(st.composite(_random_records)())
    def run(rec):
        bamio.write_bam(p, rec)
        back = bamio.read_bam(p)
        _same(rec, back)
        nb = bamio.NativeBam(p, 2)
        try:
            assert nb.n == rec.n
            for rid in (0, 1):
                for mmq in (0, 30):
                    try:
                        want = select_reads(back, rid, mmq)
                    except TypeError:                    # a kept read without QUAL: both paths refuse, as the reference
                        with pytest.raises(TypeError):
                            nb.select(rid, mmq)
                        with pytest.raises(TypeError):
                            nb.pack(rid, mmq, 0)
                        continue
                    _same_batch(nb.select(rid, mmq), want)
            # the same file as a stream of one-block spans: the records' core fields in order, nothing lost or doubled
            with bamio.NativeBamStream(p, 2, span_bytes=1) as stream:
                spans = list(stream)
            try:
                assert sum(s.n for s in spans) == rec.n
                for k, want_k in enumerate(nb.core()):
                    assert np.array_equal(np.concatenate([s.core()[k] for s in spans]), want_k)
            finally:
                for s in spans:
                    s.close()
        finally:
            nb.close()

    run()


def test_block_crc_matches_zlib():
    """bc_bgzf_crc32 (carry-less-multiply folding where the CPU has PCLMULQDQ) against zlib.crc32: every length
    around the 16- and 64-byte steps of the fold, unaligned starts, one BGZF-sized block, a few MB."""
    from basecount_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    assert L.bc_bgzf_crc32(None, 0) == zlib.crc32(b"")
    big = rng.integers(0, 256, size=(3 << 20) + 77, dtype=np.uint8)
    for n in list(range(1, 260)) + [1023, 4095, 4096, 65279, 65280, 65281, big.size - 3]:
        for off in (0, 1, 3):
            d = np.ascontiguousarray(big[off:off + n])
            assert L.bc_bgzf_crc32(_lib.ptr(d), d.size) == zlib.crc32(d.tobytes()), (n, off)
    z = np.zeros(100_000, dtype=np.uint8)
    assert L.bc_bgzf_crc32(_lib.ptr(z), z.size) == zlib.crc32(z.tobytes())


def _raw_deflate(data: bytes, level: int, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def test_block_inflater_matches_zlib():
    """csrc/inflate_fast.h (the decoder that runs before zlib on every BGZF block) against zlib on raw DEFLATE
    streams of every kind: stored, fixed-Huffman and dynamic blocks, levels 1..9, literal-heavy and repetitive
    payloads, long matches at distance 1, empty input, multi-block streams; a wrong output size is declined."""
    from basecount_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)

    def run(raw: bytes, out_len: int):
        src = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
        out = np.full(out_len, 0xEE, dtype=np.uint8)
        ok = L.bc_inflate_raw(_lib.ptr(src) if src.size else None, src.size, _lib.ptr(out) if out_len else None, out_len)
        return ok, out.tobytes()

    payloads = [b"", b"A", b"ACGT" * 5000, bytes(70000), rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(),
                rng.integers(2, 41, 60000, dtype=np.uint8).tobytes(),                    # qualities
                (rng.integers(0, 4, 30000, dtype=np.uint8) * 17).astype(np.uint8).tobytes(),
                b"".join(bytes([i % 251]) * (i % 300) for i in range(400)),
                open(__file__, "rb").read()]
    rec = synth.amplicon_sample(seed=6, n_reads=300, ref_len=3000, ref_name="x")
    payloads.append(bamio.encode_bam_bytes(rec)[:65000])
    n = 0
    for data in payloads:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                raw = _raw_deflate(data, level, strategy)
                ok, out = run(raw, len(data))
                assert ok == 1 and out == data, (len(data), level, strategy)
                n += 1
                if len(data) > 1:
                    assert run(raw, len(data) - 1)[0] == 0                 # too small an output: declined
                    assert run(raw, len(data) + 1)[0] == 0                 # too large: declined
    assert n == len(payloads) * 16


def test_block_inflater_survives_damaged_streams():
    """Truncated and bit-flipped streams: the decoder must decline or return bytes (the readers' CRC check
    catches those) without reading or writing out of bounds -- the output array carries guard bytes."""
    from basecount_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(9)
    data = rng.integers(2, 41, 20000, dtype=np.uint8).tobytes() + b"ACGTTGCA" * 800
    raw = bytearray(_raw_deflate(data, 6))
    declined = 0
    for trial in range(400):
        d = bytearray(raw)
        if trial % 2:
            d = d[:int(rng.integers(0, len(d)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        src = np.frombuffer(bytes(d), dtype=np.uint8) if d else np.zeros(0, np.uint8)
        out = np.full(len(data) + 64, 0xEE, dtype=np.uint8)
        ok = L.bc_inflate_raw(_lib.ptr(src) if src.size else None, src.size, _lib.ptr(out), len(data))
        assert (out[len(data):] == 0xEE).all()                             # nothing past the block's ISIZE bytes
        if ok:
            try:
                want = zlib.decompress(bytes(d), -15)
            except zlib.error:
                want = None
            # an accepted stream either is what zlib makes of it, or differs and is left to the CRC check
            assert want is None or want != out[:len(data)].tobytes() or len(want) == len(data)
        else:
            declined += 1
    assert declined > 100


def test_native_decoder_on_randomly_damaged_files(tmp_path):
    """Single-byte damage anywhere in the file: the reader either raises (header, inflate or CRC failure -- the block
    decoder declines, zlib refuses, or the CRC differs) or, when the byte did not matter (gzip MTIME / XFL / OS
    fields), returns exactly the undamaged records.  It never returns different data and never crashes."""
    rec = synth.amplicon_sample(seed=12, n_reads=600, ref_len=3000, ref_name="x")
    p = str(tmp_path / "ok.bam")
    bamio.write_bam(p, rec)
    data = open(p, "rb").read()
    good = bamio.NativeBam(p, 2)
    want = good.select(0, 0)
    good.close()
    rng = np.random.default_rng(21)
    bad = str(tmp_path / "bad.bam")
    raised = 0
    for _ in range(60):
        d = bytearray(data)
        d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        open(bad, "wb").write(bytes(d))
        try:
            nb = bamio.NativeBam(bad, 2)
        except ValueError:
            raised += 1
            continue
        try:
            _same_batch(nb.select(0, 0), want)
        finally:
            nb.close()
    assert raised >= 50


def _raw_record(ref_id, pos, mapq, flag, cigar_words, seq_ascii, qual, aux=b"", name=b"q1\0"):
    """One BAM alignment record (SAM spec 4.2) with optional fields, block_size included."""
    code = {c: i for i, c in enumerate(b"=ACMGRSVTWYHKDBN")}
    l_seq = len(seq_ascii)
    packed = bytearray((l_seq + 1) // 2)
    for j, ch in enumerate(seq_ascii):
        packed[j >> 1] |= code[ch] << (4 if j % 2 == 0 else 0)
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name), mapq, 4680, len(cigar_words), flag, l_seq, -1, -1, 0)
    body += name + b"".join(struct.pack("<I", w) for w in cigar_words) + bytes(packed) + bytes(qual) + aux
    return struct.pack("<i", len(body)) + body


def _bam_with(records_bytes, ref_name="chrT", ref_len=5000):
    rec0 = synth.take_records(synth.amplicon_sample(seed=5, n_reads=5, ref_len=ref_len, ref_name=ref_name),
                              np.zeros(0, dtype=np.int64))
    return bamio.bgzf_compress(bamio.encode_bam_bytes(rec0) + b"".join(records_bytes), 1)


def test_long_cigar_is_taken_from_the_cg_tag(tmp_path):
    """More than 65535 CIGAR operations: the record holds the placeholder <l_seq>S<span>N and the real CIGAR sits
    in CG:B,I (SAM spec 4.2.2); htslib / pysam hand out the real one, and so must both decoders -- with other
    optional fields of every type in front of the tag."""
    real = [(3 << 4) | 0, (1 << 4) | 2, (5 << 4) | 0]                      # 3M1D5M over "ACGTACGT"
    seq, qual = b"ACGTACGT", [30] * 8
    aux = (b"NMC\x01" + b"XAA!" + b"XSs" + struct.pack("<h", -2) + b"XIi" + struct.pack("<i", 7) + b"XFf" + struct.pack("<f", 1.5) +
           b"RGZgroup\0" + b"XHH1AE3\0" + b"XBBc" + struct.pack("<I", 3) + b"\x01\x02\x03" +
           b"CGBI" + struct.pack("<I", len(real)) + b"".join(struct.pack("<I", w) for w in real) + b"ZZC\x05")
    placeholder = [(len(seq) << 4) | 4, (9 << 4) | 3]
    plain = _raw_record(0, 100, 60, 0, real, seq, qual, name=b"q0\0")
    tagged = _raw_record(0, 200, 60, 0, placeholder, seq, qual, aux)
    untagged = _raw_record(0, 300, 60, 0, placeholder, seq, qual, b"NMC\x01")     # no CG tag: stays S + N, as in htslib
    p = str(tmp_path / "cg.bam")
    open(p, "wb").write(_bam_with([plain, tagged, untagged]))
    py = bamio.read_bam(p)
    assert py.n == 3
    assert py.cigar[py.cigar_off[0]:py.cigar_off[1]].tolist() == real
    assert py.cigar[py.cigar_off[1]:py.cigar_off[2]].tolist() == real
    assert py.cigar[py.cigar_off[2]:py.cigar_off[3]].tolist() == placeholder
    nb = bamio.NativeBam(p, 2)
    _same_batch(nb.select(0, 0), select_reads(py, 0, 0))
    b = nb.select(0, 0)
    assert b.cigar[b.cigar_off[1]:b.cigar_off[2]].tolist() == real and b.seq_off.tolist() == [0, 8, 16, 16]
    packed = nb.pack(0, 0, 0)
    assert packed.n_reads == 3 and packed.aligned_bases == 2 * 9 + 9
    nb.close()


def test_missing_qualities_raise_type_error(tmp_path):
    """QUAL '*' (0xFF bytes): pysam gives the reference None and bcount raises TypeError (count.cpp:11), at any
    min_base_quality -- but only for reads the filter keeps (main.py:165)."""
    good = _raw_record(0, 100, 60, 0, [(8 << 4) | 0], b"ACGTACGT", [30] * 8, name=b"q0\0")
    noqual = _raw_record(0, 200, 20, 0, [(8 << 4) | 0], b"ACGTACGT", [0xFF] * 8)
    p = str(tmp_path / "nq.bam")
    open(p, "wb").write(_bam_with([good, noqual]))
    py = bamio.read_bam(p)
    nb = bamio.NativeBam(p, 2)
    for fn in (lambda mmq: nb.select(0, mmq), lambda mmq: nb.pack(0, mmq, 0), lambda mmq: select_reads(py, 0, mmq)):
        with pytest.raises(TypeError):
            fn(0)
        fn(30)                                      # the read without qualities is filtered out: no error
    nb.close()


def test_negative_mapping_quality_keeps_every_mapped_read(tmp_path):
    rec = synth.amplicon_sample(seed=6, n_reads=400, ref_len=3000, ref_name="chrT")
    p = str(tmp_path / "m.bam")
    bamio.write_bam(p, rec)
    py = bamio.read_bam(p)
    nb = bamio.NativeBam(p, 2)
    want = select_reads(py, 0, -1)
    assert want.n == select_reads(py, 0, 0).n > 0
    _same_batch(nb.select(0, -1), want)
    assert nb.pack(0, -1, 0).n_reads == want.n
    nb.close()


def test_stream_of_spans_equals_the_whole_file(tmp_path):
    """bamio.NativeBamStream (bc_bam_stream_*: the file span by span, bounded host memory) must hand out exactly the
    records of the whole-file decoder, in order and whole, for any span size -- spans of one BGZF block, spans that
    end inside a record, a record longer than the span (the reader doubles it), a file without records (one empty span
    that still carries the header), two references -- and reject a file cut inside a record."""
    big = _raw_record(0, 100, 60, 0, [(200_000 << 4) | 0], b"ACGT" * 50_000, [30] * 200_000)       # 300 kB record
    cases = {"amplicon": synth.amplicon_sample(seed=5, n_reads=6000, ref_len=4000, ref_name="chrT"),
             "torture": _clip_torture_records(),
             "empty": synth.take_records(synth.amplicon_sample(seed=5, n_reads=50, ref_len=4000, ref_name="chrT"),
                                         np.zeros(0, dtype=np.int64))}
    paths = {}
    for name, rec in cases.items():
        paths[name] = str(tmp_path / f"{name}.bam")
        bamio.write_bam(paths[name], rec)
    paths["long_record"] = str(tmp_path / "long.bam")
    with open(paths["long_record"], "wb") as fh:
        small = _raw_record(0, 7, 60, 0, [(4 << 4) | 0], b"ACGT", [30] * 4)
        fh.write(_bam_with([small, big, small, small], ref_len=300_000))
    for name, p in paths.items():
        whole = bamio.NativeBam(p, 2)
        w_core = whole.core()
        for span_bytes in (1, 70_000, 200_000, 1 << 30):               # (the library's floor is one BGZF block, 64 KiB)
            with bamio.NativeBamStream(p, 2, span_bytes=span_bytes) as st:
                spans = list(st)
            assert len(spans) >= 1 and all(s.ref_names == whole.ref_names and s.ref_lengths == whole.ref_lengths for s in spans)
            assert sum(s.n for s in spans) == whole.n, (name, span_bytes)
            if span_bytes < 100_000 and whole.n > 1000:
                assert len(spans) > 3                                   # it really was cut
            for k in range(4):
                got = np.concatenate([s.core()[k] for s in spans]) if spans else np.zeros(0)
                assert np.array_equal(got, w_core[k]), (name, span_bytes, k)
            for rid in range(len(whole.ref_names)):
                want = whole.select(rid, 0)
                parts = [s.select(rid, 0) for s in spans]
                assert sum(b.n for b in parts) == want.n
                assert np.array_equal(np.concatenate([b.starts for b in parts]), want.starts)
                assert np.array_equal(np.concatenate([b.seq for b in parts]), want.seq)
                assert np.array_equal(np.concatenate([b.qual for b in parts]), want.qual)
                assert np.array_equal(np.concatenate([b.cigar for b in parts]), want.cigar)
            for s in spans:
                s.close()
        whole.close()
    # cut inside the last record: the whole-file decoder and the stream both refuse
    raw = bamio.encode_bam_bytes(cases["amplicon"])
    cut = str(tmp_path / "cut.bam")
    with open(cut, "wb") as fh:
        fh.write(bamio.bgzf_compress(raw[:-11], 1))
    with pytest.raises(ValueError):
        bamio.NativeBam(cut, 2)
    with pytest.raises(ValueError):
        with bamio.NativeBamStream(cut, 2, span_bytes=70_000) as st:
            for s in st:
                s.close()
