"""BGZF / BAM reader and writer (SAM spec v1 section 4), vectorised with numpy.

The reference reads alignments through pysam (basecount/main.py:97-99,127,165-173).
pysam / htslib are not installable in this image (no network), so the product carries
its own decoder: `read_bam()` turns a BAM file straight into the flat `Records` arrays
the packer consumes -- no per-read Python objects -- and `AlignmentFile` offers the small
pysam surface the reference touches (references, lengths, fetch(until_eof=True), close,
and per-read is_unmapped / mapping_quality / reference_name / reference_start /
query_alignment_sequence / query_alignment_qualities / cigartuples) for callers and tests
that want read objects.  If a real `pysam` is importable, main.py prefers it.

Parity note: nothing the reference ships pins this boundary (its tests need pysam and
external BAMs); the decoder is validated against the SAM specification by round-trip
with the writer below (tests/test_bamio.py).
"""
from __future__ import annotations

import array
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .records import FLAG_UNMAPPED, Records, _leading_trailing_clips

_NIB2ASCII = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
_ASCII2NIB = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(b"=ACMGRSVTWYHKDBN"):
    _ASCII2NIB[_c] = _i
    _ASCII2NIB[ord(chr(_c).lower())] = _i
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_MAX_BLOCK = 0xFF00


# ----------------------------------------------------------------------------- BGZF
def _bgzf_blocks(data: bytes):
    """Yield (cdata_start, cdata_end, isize) for every BGZF block in `data`."""
    off, n = 0, len(data)
    while off < n:
        if n - off < 18 or data[off] != 31 or data[off + 1] != 139 or data[off + 2] != 8 or not (data[off + 3] & 4):
            raise ValueError("not a BGZF file (bad gzip member header)")
        xlen = data[off + 10] | (data[off + 11] << 8)
        p, end, bsize = off + 12, off + 12 + xlen, None
        while p + 4 <= end:
            slen = data[p + 2] | (data[p + 3] << 8)
            if data[p] == 66 and data[p + 1] == 67 and slen == 2:
                bsize = (data[p + 4] | (data[p + 5] << 8)) + 1
            p += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield")
        isize = int.from_bytes(data[off + bsize - 4:off + bsize], "little")
        yield off + 12 + xlen, off + bsize - 8, isize
        off += bsize


def bgzf_decompress(data: bytes, threads: int = 8) -> bytes:
    blocks = list(_bgzf_blocks(data))
    mv = memoryview(data)

    def inflate(b):
        return zlib.decompress(mv[b[0]:b[1]], -15) if b[2] else b""

    if threads > 1 and len(blocks) > 64:
        with ThreadPoolExecutor(max_workers=threads) as ex:       # zlib releases the GIL
            parts = list(ex.map(inflate, blocks, chunksize=64))
    else:
        parts = [inflate(b) for b in blocks]
    return b"".join(parts)


def bgzf_compress(raw: bytes, level: int = 1) -> bytes:
    out = []
    for a in range(0, len(raw), _MAX_BLOCK):
        chunk = raw[a:a + _MAX_BLOCK]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = co.compress(chunk) + co.flush()
        bsize = len(cdata) + 25
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + cdata +
                   struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    out.append(_BGZF_EOF)
    return b"".join(out)


# ----------------------------------------------------------------------------- BAM -> Records
def _parse_header(buf: bytes):
    if buf[:4] != b"BAM\x01":
        raise ValueError("not a BAM file (bad magic)")
    l_text = struct.unpack_from("<i", buf, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, p)[0]
    p += 4
    names, lengths = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", buf, p)[0]
        names.append(buf[p + 4:p + 4 + l_name - 1].decode("ascii"))
        lengths.append(struct.unpack_from("<i", buf, p + 4 + l_name)[0])
        p += 8 + l_name
    return names, lengths, p


def _record_offsets(buf: bytes, p: int) -> np.ndarray:
    offs = array.array("q")
    n = len(buf)
    unpack = struct.Struct("<i").unpack_from
    append = offs.append
    while p + 4 <= n:
        append(p)
        p += 4 + unpack(buf, p)[0]
    if p != n:
        raise ValueError("truncated BAM record")
    return np.frombuffer(offs, dtype=np.int64) if len(offs) else np.zeros(0, np.int64)


_AUX_SIZE = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}


def _find_cg_tag(buf: bytes, a: int, end: int):
    """(offset of the first word, number of words) of a CG:B,I tag among the optional fields buf[a:end], or None."""
    while a + 3 <= end:
        tag, ty = buf[a:a + 2], buf[a + 2:a + 3]
        a += 3
        if ty in _AUX_SIZE:
            a += _AUX_SIZE[ty]
        elif ty in (b"Z", b"H"):
            z = buf.find(b"\0", a, end)
            if z < 0:
                raise ValueError("truncated BAM record")
            a = z + 1
        elif ty == b"B":
            if a + 5 > end:
                raise ValueError("truncated BAM record")
            sub, cnt = buf[a:a + 1], struct.unpack_from("<I", buf, a + 1)[0]
            if sub not in _AUX_SIZE:
                raise ValueError("malformed BAM optional field")
            if tag == b"CG" and sub == b"I":
                if a + 5 + 4 * cnt > end:
                    raise ValueError("truncated BAM record")
                return a + 5, cnt
            a += 5 + _AUX_SIZE[sub] * cnt
        else:
            raise ValueError("malformed BAM optional field")
    return None


def _ragged_gather(u8: np.ndarray, starts: np.ndarray, lens: np.ndarray):
    """Concatenate u8[starts[i] : starts[i]+lens[i]] for all i; returns (bytes, offsets)."""
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    if total == 0:
        return np.zeros(0, dtype=np.uint8), off
    idx = np.repeat(starts - off[:-1], lens) + np.arange(total, dtype=np.int64)
    return u8[idx], off


def decode_bam_bytes(buf: bytes) -> Records:
    """Uncompressed BAM stream -> Records (all alignment records, in file order)."""
    names, lengths, p0 = _parse_header(buf)
    u8 = np.frombuffer(buf, dtype=np.uint8)
    offs = _record_offsets(buf, p0)
    n = offs.size
    if n == 0:
        z = np.zeros(0, np.int64)
        return Records(names, lengths, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                       np.zeros(0, np.uint16), np.zeros(0, np.uint32), np.zeros(1, np.int64), np.zeros(0, np.uint8),
                       np.zeros(0, np.uint8), np.zeros(1, np.int64))
    fixed = u8[offs[:, None] + np.arange(4, 36, dtype=np.int64)]          # (n, 32) core fields after block_size
    i32 = np.ascontiguousarray(fixed).view("<i4")                        # (n, 8)
    ref_id = i32[:, 0].copy()
    pos = i32[:, 1].copy()
    l_read_name = fixed[:, 8].astype(np.int64)
    mapq = fixed[:, 9].copy()
    u16 = np.ascontiguousarray(fixed[:, 12:16]).view("<u2")
    n_cigar = u16[:, 0].astype(np.int64)
    flag = u16[:, 1].copy()
    l_seq = i32[:, 4].astype(np.int64)
    c_start = offs + 36 + l_read_name
    s_start = c_start + 4 * n_cigar
    q_start = s_start + (l_seq + 1) // 2

    # A CIGAR of more than 65535 operations is stored as the placeholder "<l_seq>S<span>N" with the real CIGAR in
    # the CG:B,I tag (SAM spec 4.2.2); htslib / pysam's cigartuples (basecount/main.py:173) hand out the real one.
    cg_start, cg_n = c_start, n_cigar
    for i in np.flatnonzero(n_cigar == 2):
        w0, w1 = struct.unpack_from("<II", buf, int(c_start[i]))
        if (w0 & 15) == 4 and (w0 >> 4) == int(l_seq[i]) and (w1 & 15) == 3:
            end = int(offs[i + 1]) if i + 1 < n else len(buf)
            hit = _find_cg_tag(buf, int(q_start[i] + l_seq[i]), end)
            if hit is not None:
                if cg_start is c_start:
                    cg_start, cg_n = c_start.copy(), n_cigar.copy()
                cg_start[i], cg_n[i] = hit
    cig_bytes, cb_off = _ragged_gather(u8, cg_start, 4 * cg_n)
    cigar = np.ascontiguousarray(cig_bytes).view("<u4").astype(np.uint32)
    cigar_off = cb_off // 4

    seq_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(l_seq, out=seq_off[1:])
    total = int(seq_off[-1])
    if total:
        within = np.arange(total, dtype=np.int64) - np.repeat(seq_off[:-1], l_seq)
        packed = u8[np.repeat(s_start, l_seq) + (within >> 1)]
        nib = np.where((within & 1) == 0, packed >> 4, packed & 15)
        seq = _NIB2ASCII[nib]
    else:
        seq = np.zeros(0, dtype=np.uint8)
    qual, _ = _ragged_gather(u8, q_start, l_seq)
    return Records(names, lengths, ref_id, pos, mapq, flag, cigar, cigar_off, seq, qual, seq_off)


def read_bam(path: str, threads: int = 8) -> Records:
    with open(path, "rb") as fh:
        data = fh.read()
    return decode_bam_bytes(bgzf_decompress(data, threads))


# ----------------------------------------------------------------------------- native decoder
class NativeBam:
    """BAM file decoded by the C-ABI library (csrc/bam_decode.h): block-parallel inflate, one record
    index, and thread-parallel selections straight into ReadBatch arrays -- no per-read objects.
    Same results as `select_reads(read_bam(path), ...)` (tests/test_bamio.py)."""

    def __init__(self, path: str, threads: int = 0, region=None, index: str | None = None):
        """region = (reference id, beg, end): only the records of that reference STARTING in [beg, end)
        (0-based, half open) are read, through the BAI index (`index`, default `path + ".bai"`)."""
        import ctypes
        from . import _lib
        self._h = None
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        if region is None:
            rc = self._L.bc_bam_open(str(path).encode(), int(threads), ctypes.byref(h))
        else:
            rid, beg, end = region
            rc = self._L.bc_bam_open_region(str(path).encode(), str(index or (str(path) + ".bai")).encode(), int(rid),
                                            int(beg), int(end), int(threads), ctypes.byref(h))
        if rc != 0 or not h.value:
            msg = self._L.bc_bam_last_error()
            raise ValueError(msg.decode() if msg else "cannot read BAM file")
        self._adopt(h)

    @classmethod
    def _from_handle(cls, h):
        """A NativeBam over a bc_bam the library has already opened (a span of a NativeBamStream)."""
        from . import _lib
        self = cls.__new__(cls)
        self._h = None
        self._L = _lib.lib()
        self._adopt(h)
        return self

    def _adopt(self, h):
        self._h = h
        self.n = int(self._L.bc_bam_num_records(h))
        nref = int(self._L.bc_bam_num_refs(h))
        self.ref_names = [self._L.bc_bam_ref_name(h, i).decode("ascii") for i in range(nref)]
        self.ref_lengths = [int(self._L.bc_bam_ref_len(h, i)) for i in range(nref)]
        self._core = None

    def close(self):
        if self._h is not None:
            self._L.bc_bam_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def core(self):
        """(ref_id int32[n], pos int32[n], mapq uint8[n], flag uint16[n]) of every record."""
        if self._core is None:
            from . import _lib
            ref_id = np.empty(self.n, np.int32)
            pos = np.empty(self.n, np.int32)
            mapq = np.empty(self.n, np.uint8)
            flag = np.empty(self.n, np.uint16)
            self._L.bc_bam_core(self._h, _lib.ptr(ref_id), _lib.ptr(pos), _lib.ptr(mapq), _lib.ptr(flag))
            self._core = (ref_id, pos, mapq, flag)
        return self._core

    def select(self, ref_id: int, min_mapping_quality: int = 0, rec_a: int = 0, rec_b: int | None = None,
               want_qual: bool = True):
        """ReadBatch of the kept reads of one reference among records [rec_a, rec_b).  want_qual=False leaves
        the qualities out (an empty array): with min_base_quality 0 nothing reads them (count.cpp:56)."""
        import ctypes
        from . import _lib
        from .records import ReadBatch
        rec_b = self.n if rec_b is None else rec_b
        min_mapping_quality = max(int(min_mapping_quality), 0)      # MAPQ is unsigned: a negative threshold keeps everything
        nr, nc, nb = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        rc = self._L.bc_bam_select_sizes(self._h, rec_a, rec_b, int(ref_id), int(min_mapping_quality), ctypes.byref(nr),
                                         ctypes.byref(nc), ctypes.byref(nb))
        if rc == _lib.BC_ERR_MISSING_QUAL:
            raise TypeError("a read has no base qualities (QUAL '*'): the reference's bcount raises TypeError on None")
        if rc != _lib.BC_OK:
            raise TypeError(f"bc_bam_select_sizes failed with status {rc}")
        n = nr.value
        starts = np.empty(n, np.uint32)
        cigar = np.empty(nc.value, np.uint32)
        cigar_off = np.empty(n + 1, np.uint64)
        seq = np.empty(nb.value, np.uint8)
        qual = np.empty(nb.value if want_qual else 0, np.uint8)
        seq_off = np.empty(n + 1, np.uint64)
        self._L.bc_bam_select_fill(self._h, rec_a, rec_b, int(ref_id), int(min_mapping_quality), _lib.ptr(starts),
                                   _lib.ptr(cigar), _lib.ptr(cigar_off), _lib.ptr(seq), _lib.ptr(qual) if want_qual else None,
                                   _lib.ptr(seq_off))
        return ReadBatch(starts, cigar, cigar_off, seq, qual, seq_off)


def _native_pack(self, ref_id: int, min_mapping_quality: int = 0, min_base_quality: int = 0, rec_a: int = 0,
                 rec_b: int | None = None, pinned: bool = False):
    """PackedBatch (one reference slot) of the kept reads among records [rec_a, rec_b): selection, soft-clip
    trimming and 2-bit packing in one native pass -- the same arrays as pack_batches(self.select(...))."""
    import ctypes
    from . import _lib
    from .pack import PackedBatch, _alloc
    if min_base_quality < 0:
        raise TypeError("min_base_quality must be unsigned")
    rec_b = self.n if rec_b is None else rec_b
    min_mapping_quality = max(int(min_mapping_quality), 0)          # MAPQ is unsigned: a negative threshold keeps everything
    z = np.zeros(6, np.uint64)
    rc = self._L.bc_bam_pack_sizes(self._h, rec_a, rec_b, int(ref_id), int(min_mapping_quality), _lib.ptr(z))
    if rc == _lib.BC_ERR_MISSING_QUAL:
        raise TypeError("a read has no base qualities (QUAL '*'): the reference's bcount raises TypeError on None")
    if rc != 0:
        raise TypeError("bc_bam_pack_sizes failed")
    n, n_cigar, n_words, n_bases, aligned, is_sorted = (int(x) for x in z)
    if n >= 2 ** 32 or n_cigar >= 2 ** 32 or n_words >= 2 ** 32:
        raise TypeError("batch does not fit 32-bit offsets")
    ref_read_off = _alloc(2, np.uint32, pinned)
    ref_read_off[0], ref_read_off[1] = 0, n
    starts = _alloc(n, np.uint32, pinned)
    cigar = _alloc(n_cigar, np.uint32, pinned)
    cigar_off = _alloc(n + 1, np.uint32, pinned)
    seq_woff = _alloc(n + 1, np.uint32, pinned)
    planes = _alloc(n_words, np.uint64, pinned)
    okmask = _alloc(n_words, np.uint32, pinned) if min_base_quality > 0 else None
    exc_cap = max(1024, n_bases // 256)
    n_exc = ctypes.c_uint64(0)
    while True:
        exc_read = _alloc(exc_cap, np.uint32, pinned)
        exc_pos = _alloc(exc_cap, np.uint32, pinned)
        rc = self._L.bc_bam_pack_fill(self._h, rec_a, rec_b, int(ref_id), int(min_mapping_quality), int(min_base_quality),
                                      _lib.ptr(starts), _lib.ptr(cigar), _lib.ptr(cigar_off), _lib.ptr(seq_woff),
                                      _lib.ptr(planes), _lib.ptr(okmask), _lib.ptr(exc_read), _lib.ptr(exc_pos), exc_cap,
                                      ctypes.byref(n_exc))
        if rc == _lib.BC_ERR_READ_OVERRUN:
            raise ValueError("a CIGAR consumes more bases than its read holds")
        if rc != _lib.BC_OK:
            raise TypeError(f"bc_bam_pack_fill failed with status {rc}")
        if n_exc.value <= exc_cap:
            break
        exc_cap = int(n_exc.value)
    ne = int(n_exc.value)
    return PackedBatch(n, 1, ref_read_off, starts, cigar_off, cigar, seq_woff, planes, okmask, exc_read[:ne], exc_pos[:ne],
                       bool(is_sorted), (n_bases // n) if n else 0, aligned, n_bases)


NativeBam.pack = _native_pack


class NativeBamStream:
    """A BAM file read span by span (bc_bam_stream_*): iterating yields NativeBam objects that each hold the whole
    records starting in the next ~span_bytes of the inflated stream, so host memory is bounded by the span (two of them:
    the next span is inflated on a second thread while the caller works on the current one) whatever the size of the
    file -- the role of the reference's fetch(until_eof=True) loop with its chunk_size (main.py:127,142).  The first
    span exists even for a file without records, so `ref_names` / `ref_lengths` are always available from it.
    The caller closes every span it is handed."""

    def __init__(self, path: str, threads: int = 0, span_bytes: int | None = None):
        import ctypes
        import os
        from . import _lib
        self._L = _lib.lib()
        self._s = None
        s = ctypes.c_void_p()
        rc = self._L.bc_bam_stream_open(str(path).encode(), int(threads), ctypes.byref(s))
        if rc != 0 or not s.value:
            msg = self._L.bc_bam_last_error()
            raise ValueError(msg.decode() if msg else "cannot read BAM file")
        self._s = s
        if span_bytes is None:
            span_bytes = int(float(os.environ.get("BASECOUNT_B200_SPAN_MB", "512")) * (1 << 20))
        self.span_bytes = max(int(span_bytes), 1)

    def _next(self):
        import ctypes
        h = ctypes.c_void_p()
        rc = self._L.bc_bam_stream_next(self._s, self.span_bytes, ctypes.byref(h))
        if rc != 0:
            msg = self._L.bc_bam_last_error()
            raise ValueError(msg.decode() if msg else "cannot read BAM file")
        return NativeBam._from_handle(h) if h.value else None

    def __iter__(self):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as pool:
            fut = pool.submit(self._next)
            while True:
                try:
                    span = fut.result()
                except BaseException:
                    raise
                if span is None:
                    return
                fut = pool.submit(self._next)
                try:
                    yield span
                except GeneratorExit:
                    nxt = fut.result()
                    if nxt is not None:
                        nxt.close()
                    raise

    def close(self):
        if self._s is not None:
            self._L.bc_bam_stream_close(self._s)
            self._s = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def write_bai(bam_path: str, bai_path: str | None = None, threads: int = 0) -> str:
    """Index a coordinate-sorted BAM (what `pysam.index` does for the reference's tests,
    tests/test_basecount.py:343-344); returns the path of the .bai."""
    from . import _lib
    L = _lib.lib()
    bai_path = bai_path or (str(bam_path) + ".bai")
    if L.bc_bam_index_build(str(bam_path).encode(), str(bai_path).encode(), int(threads)) != 0:
        msg = L.bc_bam_last_error()
        raise ValueError(msg.decode() if msg else "cannot index BAM file")
    return bai_path


# ----------------------------------------------------------------------------- Records -> BAM
def _reg2bin(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    end = end - 1
    out = np.zeros(beg.shape, dtype=np.int64)
    done = np.zeros(beg.shape, dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = base + (beg[hit] >> shift)
        done |= hit
    return out


def encode_bam_bytes(rec: Records) -> bytes:
    """Records -> uncompressed BAM stream (no aux tags; read names r000000000...)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(
        f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(rec.ref_names, rec.ref_lengths))
    head = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(rec.ref_names))]
    for n, l in zip(rec.ref_names, rec.ref_lengths):
        nm = n.encode() + b"\x00"
        head += [struct.pack("<i", len(nm)), nm, struct.pack("<i", l)]
    header = b"".join(head)
    n = rec.n
    if n == 0:
        return header
    name_len = 11                                                     # "r%09d\0"
    n_cigar = (rec.cigar_off[1:] - rec.cigar_off[:-1]).astype(np.int64)
    l_seq = (rec.seq_off[1:] - rec.seq_off[:-1]).astype(np.int64)
    if (n_cigar > 65535).any():
        raise ValueError("more than 65535 CIGAR ops in a record is not supported by this writer")
    size = 32 + name_len + 4 * n_cigar + (l_seq + 1) // 2 + l_seq      # block_size value
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(size + 4, out=off[1:])
    out = np.zeros(int(off[-1]), dtype=np.uint8)
    o = off[:-1]

    op = (rec.cigar & 0xF).astype(np.int64)
    ln = (rec.cigar >> 4).astype(np.int64)
    ref_ln = np.where(np.isin(op, [0, 2, 3, 7, 8]), ln, 0)
    span = np.add.reduceat(np.concatenate([ref_ln, [0]]), np.minimum(rec.cigar_off[:-1], ref_ln.size)) if ref_ln.size else np.zeros(n, np.int64)
    span = np.where(n_cigar > 0, span, 0)
    p64 = rec.pos.astype(np.int64)
    bins = np.where(p64 >= 0, _reg2bin(np.maximum(p64, 0), np.maximum(p64, 0) + np.maximum(span, 1)), 4680)

    core = np.zeros((n, 36), dtype=np.uint8)
    core[:, 0:4] = size.astype("<i4").view(np.uint8).reshape(n, 4)
    core[:, 4:8] = rec.ref_id.astype("<i4").view(np.uint8).reshape(n, 4)
    core[:, 8:12] = rec.pos.astype("<i4").view(np.uint8).reshape(n, 4)
    core[:, 12] = name_len
    core[:, 13] = rec.mapq
    core[:, 14:16] = bins.astype("<u2").view(np.uint8).reshape(n, 2)
    core[:, 16:18] = n_cigar.astype("<u2").view(np.uint8).reshape(n, 2)
    core[:, 18:20] = rec.flag.astype("<u2").view(np.uint8).reshape(n, 2)
    core[:, 20:24] = l_seq.astype("<i4").view(np.uint8).reshape(n, 4)
    core[:, 24:28] = np.full(n, -1, dtype="<i4").view(np.uint8).reshape(n, 4)
    core[:, 28:32] = np.full(n, -1, dtype="<i4").view(np.uint8).reshape(n, 4)
    out[o[:, None] + np.arange(36)] = core
    digits = (np.arange(n, dtype=np.int64)[:, None] // 10 ** np.arange(8, -1, -1, dtype=np.int64)) % 10
    names = np.concatenate([np.full((n, 1), ord("r")), digits + ord("0"), np.zeros((n, 1), dtype=np.int64)], axis=1)
    out[o[:, None] + 36 + np.arange(name_len)] = names.astype(np.uint8)

    c_start = o + 36 + name_len
    total_c = int(n_cigar.sum())
    if total_c:
        cw = rec.cigar.astype("<u4").view(np.uint8).reshape(-1, 4)
        dst = np.repeat(c_start - 4 * rec.cigar_off[:-1], n_cigar) + 4 * np.arange(total_c, dtype=np.int64)
        out[dst[:, None] + np.arange(4)] = cw
    s_start = c_start + 4 * n_cigar
    total = int(l_seq.sum())
    if total:
        within = np.arange(total, dtype=np.int64) - np.repeat(rec.seq_off[:-1], l_seq)
        nib = _ASCII2NIB[rec.seq]
        dst = np.repeat(s_start, l_seq) + (within >> 1)
        hi = (within & 1) == 0
        np.add.at(out, dst[hi], nib[hi] << 4)        # each byte receives one high and at most one low nibble
        np.add.at(out, dst[~hi], nib[~hi])
        q_start = s_start + (l_seq + 1) // 2
        out[np.repeat(q_start, l_seq) + within] = rec.qual
    return header + out.tobytes()


def write_bam(path: str, rec: Records, level: int = 1) -> None:
    with open(path, "wb") as fh:
        fh.write(bgzf_compress(encode_bam_bytes(rec), level))


# ----------------------------------------------------------------------------- pysam-shaped surface
class AlignedSegment:
    __slots__ = ("is_unmapped", "mapping_quality", "reference_name", "reference_start", "query_alignment_sequence",
                 "query_alignment_qualities", "cigartuples", "flag")


class AlignmentFile:
    """The slice of pysam.AlignmentFile the reference uses (basecount/main.py:98,122,127,204)."""

    def __init__(self, path, mode="rb"):
        if mode != "rb":
            raise ValueError("only mode='rb' is supported")
        self._rec = read_bam(path)
        self.references = tuple(self._rec.ref_names)
        self.lengths = tuple(self._rec.ref_lengths)

    def fetch(self, until_eof=True):
        rec = self._rec
        lead, trail = _leading_trailing_clips(rec)
        seq_b = rec.seq.tobytes()
        for i in range(rec.n):
            r = AlignedSegment()
            r.flag = int(rec.flag[i])
            r.is_unmapped = bool(rec.flag[i] & FLAG_UNMAPPED)
            r.mapping_quality = int(rec.mapq[i])
            r.reference_name = rec.ref_names[rec.ref_id[i]] if rec.ref_id[i] >= 0 else None
            r.reference_start = int(rec.pos[i])
            a, b = int(rec.seq_off[i] + lead[i]), int(rec.seq_off[i + 1] - trail[i])
            has_seq = rec.seq_off[i + 1] > rec.seq_off[i]
            r.query_alignment_sequence = seq_b[a:b].decode("ascii") if has_seq else None
            q = rec.qual[a:b]
            r.query_alignment_qualities = (array.array("B", q.tobytes())
                                           if has_seq and not (q.size and q[0] == 0xFF) else None)
            c0, c1 = int(rec.cigar_off[i]), int(rec.cigar_off[i + 1])
            r.cigartuples = [(int(w & 0xF), int(w >> 4)) for w in rec.cigar[c0:c1]] if c1 > c0 else None
            yield r

    def close(self):
        self._rec = None
