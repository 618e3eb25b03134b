"""Host side of the pileup counting path: the reference's Python API and CLI over the GPU engine.

Mirrors the public surface of basecount/main.py -- get_basecounts (main.py:110-205),
class BaseCount (main.py:208-359), handle_arg / run (main.py:362-595) -- with the same
argument names, defaults, column names, row layout, error messages and printed text.
What differs is underneath: alignments are decoded in bulk into flat arrays, packed to
2-bit SoA batches in native code, counted by K1 on the GPU, and the statistics /
summary / amplicon numbers come from K2 / K3 instead of per-position Python loops.

Deliberate, documented divergences (DESIGN.md section "Reference quirks"):
  * references are reported in BAM-header order, not Python set order (main.py:92);
  * with --references, reads on other contigs are skipped (the reference raises KeyError
    at main.py:166; its own tests' copy of the pipeline guards this, tests/test_basecount.py:300).
"""
from __future__ import annotations

import argparse

import numpy as np

from .engine import Engine
from .pack import pack_batches
from .records import FLAG_UNMAPPED, ReadBatch, Records, select_reads
from .scheme import load_scheme
from .version import __version__

BASES = ("A", "C", "G", "T", "DS", "N")          # column order of the count matrix (count.cpp:16-17, main.py:16)


# ----------------------------------------------------------------------------- alignment input
def _records_via_pysam(bam):
    """If a real pysam is installed, read through it exactly as the reference does
    (main.py:95-100,127) and convert to flat Records.  Returns None when pysam is absent."""
    try:
        import pysam
    except ImportError:
        return None
    old = pysam.set_verbosity(0)
    f = pysam.AlignmentFile(bam, mode="rb")
    pysam.set_verbosity(old)
    names, lengths = list(f.references), list(f.lengths)
    idx = {n: i for i, n in enumerate(names)}
    ref_id, pos, mapq, flag, cig, coff, seqs, quals, soff = [], [], [], [], [], [0], [], [], [0]
    for r in f.fetch(until_eof=True):
        ref_id.append(idx.get(r.reference_name, -1))
        pos.append(r.reference_start)
        mapq.append(r.mapping_quality)
        flag.append(r.flag)
        s, q = r.query_sequence, r.query_qualities
        if s is None or q is None:
            if not (r.flag & FLAG_UNMAPPED):
                raise TypeError("read without SEQ or QUAL")        # the reference passes None to bcount -> TypeError
            s, q = "", []
        for op, ln in (r.cigartuples or []):
            cig.append((ln << 4) | op)
        coff.append(len(cig))
        seqs.append(s)
        quals.append(np.asarray(q, dtype=np.uint8))
        soff.append(soff[-1] + len(s))
    f.close()
    return Records(names, lengths, np.asarray(ref_id, np.int32), np.asarray(pos, np.int32), np.asarray(mapq, np.uint8),
                   np.asarray(flag, np.uint16), np.asarray(cig, np.uint32), np.asarray(coff, np.int64),
                   np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8),
                   np.concatenate(quals) if quals else np.zeros(0, np.uint8), np.asarray(soff, np.int64))


def load_records(bam) -> Records:
    if isinstance(bam, Records):
        return bam
    rec = _records_via_pysam(bam)
    if rec is None:
        from . import bamio
        rec = bamio.read_bam(bam)
    return rec


def _decoder_choice() -> str:
    """BASECOUNT_B200_DECODER = native (default) | pysam | python.  `native` is the C-ABI library's
    block-parallel decoder (csrc/bam_decode.h); `pysam` reads through a real pysam exactly as the
    reference does (main.py:95-100,127) when one is installed; `python` is the numpy decoder."""
    import os
    return os.environ.get("BASECOUNT_B200_DECODER", "native").lower()


class _RecordSource:
    """What count_alignments needs from an alignment file, over Records or the native decoder."""

    def __init__(self, bam):
        self.native = None
        self.rec = None
        from . import bamio as _bamio
        if isinstance(bam, _bamio.NativeBam) or (not isinstance(bam, Records) and _decoder_choice() == "native"):
            from . import bamio
            # (an already opened NativeBam lets a caller open the next file while this one is being counted)
            self.native = bam if isinstance(bam, _bamio.NativeBam) else bamio.NativeBam(bam)
            self.ref_names, self.ref_lengths, self.n = self.native.ref_names, self.native.ref_lengths, self.native.n
            self.ref_id, _, self.mapq, self.flag = self.native.core()
        else:
            if isinstance(bam, Records) or _decoder_choice() == "pysam":
                self.rec = load_records(bam)
            else:
                from . import bamio
                self.rec = bamio.read_bam(bam)
            r = self.rec
            self.ref_names, self.ref_lengths, self.n = r.ref_names, r.ref_lengths, r.n
            self.ref_id, self.mapq, self.flag = r.ref_id, r.mapq, r.flag

    def select(self, a, b, rid, min_mapping_quality, want_qual=True):
        if self.native is not None:
            return self.native.select(rid, min_mapping_quality, a, b, want_qual=want_qual)
        part = _slice_records(self.rec, a, b) if (a, b) != (0, self.rec.n) else self.rec
        return select_reads(part, rid, min_mapping_quality)

    def close(self):
        if self.native is not None:
            self.native.close()


def _slice_records(rec: Records, a: int, b: int) -> Records:
    c0, c1 = int(rec.cigar_off[a]), int(rec.cigar_off[b])
    s0, s1 = int(rec.seq_off[a]), int(rec.seq_off[b])
    return Records(rec.ref_names, rec.ref_lengths, rec.ref_id[a:b], rec.pos[a:b], rec.mapq[a:b], rec.flag[a:b],
                   rec.cigar[c0:c1], rec.cigar_off[a:b + 1] - c0, rec.seq[s0:s1], rec.qual[s0:s1],
                   rec.seq_off[a:b + 1] - s0)


def get_references(all_references, references=None):
    """Validate the requested references (main.py:82-92); BAM-header order is kept."""
    if references is None:
        return list(all_references)
    for reference in references:
        if not (reference in all_references):
            raise Exception(f"{reference} is not a valid reference")
    wanted = set(references)
    return [r for r in all_references if r in wanted]


# ----------------------------------------------------------------------------- counting
class Pileup:
    """Counts of one BAM held on the device, plus lazily fetched statistics.

    A Pileup made without an explicit engine owns its own (a handle of the C-ABI library with its accumulators):
    the reference computes a BaseCount's data eagerly (main.py:266-278), so two live objects never see each
    other; here the numbers stay on the device until they are asked for, and must not be overwritten by the next
    BAM that is counted."""

    def __init__(self, engine, references, lengths, num_reads, show_n_bases, owns_engine=False):
        self.engine = engine
        self.references = references
        self.lengths = lengths
        self.num_reads = num_reads
        self.show_n_bases = show_n_bases
        self.owns_engine = owns_engine
        self._stats = {}
        self._counts = {}

    def close(self):
        if self.owns_engine and self.engine is not None:
            self.engine.close()
        self.engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self, i):
        if i not in self._counts:
            self._counts[i] = self.engine.counts(i)
        return self._counts[i]

    def stats(self, i):
        if i not in self._stats:
            self._stats[i] = self.engine.stats(i, self.show_n_bases)
        return self._stats[i]


def count_alignments(bam, references=None, min_base_quality=0, min_mapping_quality=0, chunk_size=1000000,
                     show_n_bases=False, engine=None) -> Pileup:
    """BAM -> device count matrices (the numeric part of get_basecounts, main.py:119-189).  A file path read by the
    native decoder is taken span by span (bamio.NativeBamStream), so host memory is bounded by the span size whatever
    the size of the file, as the reference's fetch(until_eof=True) loop is by its chunk_size (main.py:127,142)."""
    from . import bamio as _bamio
    owns = engine is None
    eng = None
    stream = None
    spans = None
    if not isinstance(bam, (Records, _bamio.NativeBam)) and _decoder_choice() == "native":
        stream = _bamio.NativeBamStream(bam)
        spans = iter(stream)
        first = next(spans)                               # (exists even for a file without records)
    else:
        first = bam
    rec = None
    try:
        rec = _RecordSource(first)
        refs = get_references(rec.ref_names, references)
        ids = [rec.ref_names.index(r) for r in refs]
        lengths = [int(rec.ref_lengths[i]) for i in ids]
        eng = engine if engine is not None else Engine(0)
        num_reads = [0] * len(refs)
        if ids:
            eng.begin(lengths)
            while rec is not None:
                _push_chunks(eng, rec, ids, num_reads, min_base_quality, min_mapping_quality, chunk_size)
                rec.close()
                rec = None
                nxt = next(spans, None) if spans is not None else None
                if nxt is not None:
                    rec = _RecordSource(nxt)
            # one synchronisation at the end, where an alignment past the reference end surfaces as IndexError
            eng.sync()
    except BaseException:
        if owns and eng is not None:
            eng.close()
        raise
    finally:
        if rec is not None:
            rec.close()
        if spans is not None:
            spans.close()                                 # (closes the span the reader had prefetched)
        if stream is not None:
            stream.close()
    return Pileup(eng, refs, lengths, num_reads, show_n_bases, owns_engine=owns)


def _push_chunks(eng, rec, ids, num_reads, min_base_quality, min_mapping_quality, chunk_size):
    """Count the kept reads of `rec` (a whole file or one span of it) into the engine's accumulators."""
    # The reference flushes every `chunk_size` kept reads (main.py:142); the counts do not
    # depend on where the chunks fall, so chunking here only bounds the packed buffers.
    keep = ((rec.flag & FLAG_UNMAPPED) == 0) & (rec.mapq >= min_mapping_quality) & np.isin(rec.ref_id, ids)
    csum = np.cumsum(keep)
    total = int(csum[-1]) if csum.size else 0
    chunk_size = max(int(chunk_size), 1)
    cuts = [0]
    for k in range(chunk_size, total, chunk_size):
        cuts.append(int(np.searchsorted(csum, k, side="left")) + 1)
    cuts.append(rec.n)
    # Pushes are asynchronous (H2D on the copy stream, K1 on the compute stream, two staging sets in the
    # library): the next chunk is decoded and packed while the device counts this one.
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        if rec.native is not None:
            # one native pass per reference: selection (main.py:165-166), soft-clip trimming, CIGAR normal form
            # and 2-bit packing straight from the records; a batch fills one slot of the handle
            for j, rid in enumerate(ids):
                packed = rec.native.pack(rid, min_mapping_quality, min_base_quality, a, b)
                if packed.n_reads == 0:
                    continue
                num_reads[j] += packed.n_reads
                if len(ids) > 1:
                    off = np.zeros(len(ids) + 1, dtype=np.uint32)
                    off[j + 1:] = packed.n_reads
                    packed.ref_read_off, packed.n_refs = off, len(ids)
                eng.push(packed, keep=2)
        else:
            batches = [rec.select(a, b, rid, min_mapping_quality, want_qual=min_base_quality > 0) for rid in ids]
            for j, bt in enumerate(batches):
                num_reads[j] += bt.n
            eng.push(pack_batches(batches, min_base_quality), keep=2)


def build_rows(ref, counts, st, show_n_bases=False, long_format=False):
    """Row lists with the reference's cell types (main.py:55-78): ints for counts and for the
    zero-coverage sentinels (-1 / 1 / 1), Python floats elsewhere."""
    k = 6 if show_n_bases else 5
    L = counts.shape[0]
    cnt = counts[:, :k].tolist()
    cov = st["coverage"].tolist()
    pcs = st["pc"].T.tolist()
    ent = st["entropy"].tolist()
    sec = st["secondary"].tolist()
    flags = st["flags"].tolist()
    minus = [-1] * k
    rows = []
    names = BASES[:k]
    for i in range(L):
        f = flags[i]
        p = minus if f & 1 else pcs[i]
        e = 1 if f & 1 else ent[i]
        s = 1 if f & 2 else sec[i]
        if long_format:
            c = cnt[i]
            for j in range(k):
                rows.append([ref, i + 1, cov[i], names[j], c[j], p[j], e, s])
        else:
            rows.append([ref, i + 1, cov[i], *cnt[i], *p, e, s])
    return rows


def _offset_rows(rows, first):
    """build_rows numbers positions from 1: shift a window's rows to their place in the reference."""
    for r in rows:
        r[1] += first
        yield r


def format_rows_text(ref, counts, st, show_n_bases=False, long_format=False, decimal_places=3, first_pos=1,
                     threads=0):
    """The TSV lines of build_rows(...) as the reference prints them (main.py:456-466), produced by the
    native emitter (csrc/tsv_format.h).  Returns None when the request is outside its exactness
    envelope (decimal_places not in 0..4, non-finite or huge values): callers then format in Python."""
    import ctypes
    from . import _lib
    k = 6 if show_n_bases else 5
    L = int(counts.shape[0])
    cnt = np.ascontiguousarray(counts, dtype=np.int64)
    cov = np.ascontiguousarray(st["coverage"], dtype=np.int64)
    pc = np.ascontiguousarray(st["pc"], dtype=np.float64)
    ent = np.ascontiguousarray(st["entropy"], dtype=np.float64)
    sec = np.ascontiguousarray(st["secondary"], dtype=np.float64)
    flg = np.ascontiguousarray(st["flags"], dtype=np.uint8)
    text, n = ctypes.c_void_p(), ctypes.c_uint64()
    rc = _lib.lib().bc_format_tsv(str(ref).encode(), L, int(first_pos), k, int(bool(long_format)), int(decimal_places),
                                  _lib.ptr(cnt), _lib.ptr(cov), _lib.ptr(pc), int(pc.shape[1]) if pc.ndim == 2 else L,
                                  _lib.ptr(ent), _lib.ptr(sec), _lib.ptr(flg), int(threads), ctypes.byref(text),
                                  ctypes.byref(n))
    if rc != 0 or not text.value:
        return None
    try:
        return ctypes.string_at(text.value, n.value).decode("ascii")
    finally:
        _lib.lib().bc_free_text(text)


def get_basecounts(bam, references=None, min_base_quality=0, min_mapping_quality=0, chunk_size=1000000,
                   show_n_bases=False, long_format=False):
    """Same contract as the reference: {ref: {"rows": [...], "num_reads": int}} (main.py:192-205)."""
    pile = count_alignments(bam, references, min_base_quality, min_mapping_quality, chunk_size, show_n_bases)
    try:
        out = {}
        for i, ref in enumerate(pile.references):
            out[ref] = {"rows": build_rows(ref, pile.counts(i), pile.stats(i), show_n_bases, long_format),
                        "num_reads": pile.num_reads[i]}
        return out
    finally:
        pile.close()


def column_names(show_n_bases=False, long_format=False):
    if long_format:
        return ["reference", "position", "coverage", "base", "count", "percentage", "entropy", "secondary_entropy"]
    letters = ["a", "c", "g", "t", "ds"] + (["n"] if show_n_bases else [])
    return (["reference", "position", "coverage"] + ["num_" + x for x in letters] + ["pc_" + x for x in letters] +
            ["entropy", "secondary_entropy"])


class BaseCount:
    def __init__(self, bam, references=None, min_base_quality=0, min_mapping_quality=0, chunk_size=1000000,
                 show_n_bases=False, long_format=False):
        """
        Generate and store basecount data.

        * `bam`: path to a BAM file (no index needed).
        * `references`: names of the references to count; `None` = every reference.
        * `min_base_quality` / `min_mapping_quality`: inclusion thresholds. Default `0`.
        * `chunk_size`: max reads packed and sent to the GPU at a time. Default `1000000`.
        * `show_n_bases`: report `N` counts and include them in the statistics.
        * `long_format`: one row per (position, base) instead of one row per position.
        """
        self.columns = column_names(show_n_bases, long_format)
        self._show_n, self._long = show_n_bases, long_format
        self._pile = count_alignments(bam, references, min_base_quality, min_mapping_quality, chunk_size, show_n_bases)
        self.references = list(self._pile.references)
        # the reference's reference_lengths is len(rows) (main.py:276-278): the reference length in wide format,
        # K times it in long format, where every position is K rows
        rep = (6 if show_n_bases else 5) if long_format else 1
        self.reference_lengths = {r: n * rep for r, n in zip(self.references, self._pile.lengths)}
        self._data = None

    def close(self):
        """Release the device state behind this object (rows already materialised stay available)."""
        self._pile.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def data(self):
        """{ref: {"rows": [...], "num_reads": int}}; rows are materialised on first use."""
        if self._data is None:
            p = self._pile
            self._data = {ref: {"rows": build_rows(ref, p.counts(i), p.stats(i), self._show_n, self._long),
                                "num_reads": p.num_reads[i]} for i, ref in enumerate(self.references)}
        return self._data

    def rows_text(self, decimal_places=3):
        """Per reference, the TSV lines of rows() as `run` prints them, from the native emitter; None if
        the request is outside its exactness envelope (then format rows() in Python)."""
        p = self._pile
        out = []
        for i, ref in enumerate(self.references):
            t = format_rows_text(ref, p.counts(i), p.stats(i), self._show_n, self._long, decimal_places)
            if t is None:
                return None
            out.append(t)
        return out

    def write_tsv(self, out, decimal_places=3, window=1 << 19, header=True, threads=0):
        """Stream the per-position TSV (what `run` prints, main.py:454-466) to the text file object `out`, one window
        of `window` positions at a time: the device computes the next window's rows (K2) and copies them back
        while the native emitter formats the current one, so a 64 Mb reference never needs more host memory than
        two windows.  Returns the number of data rows written.  Falls back to the Python formatting loop, window by
        window, outside the emitter's exactness envelope (decimal_places not in 0..4)."""
        from concurrent.futures import ThreadPoolExecutor
        p = self._pile
        k = 6 if self._show_n else 5
        if header:
            out.write("\t".join(self.columns) + "\n")
        rows = 0
        with ThreadPoolExecutor(max_workers=1) as pool:      # the one thread that talks to the engine while we format
            for i, ref in enumerate(self.references):
                L = p.lengths[i]
                cuts = list(range(0, L, max(int(window), 1))) + [L]
                spans = list(zip(cuts[:-1], cuts[1:]))
                bufs = [None, None]
                fut = pool.submit(p.engine.rows_window, i, spans[0][0], spans[0][1] - spans[0][0], self._show_n, None) if spans else None
                for w, (a, b) in enumerate(spans):
                    cnt, st, bufs[w % 2] = fut.result()
                    if w + 1 < len(spans):
                        na, nb = spans[w + 1]
                        fut = pool.submit(p.engine.rows_window, i, na, nb - na, self._show_n, bufs[(w + 1) % 2])
                    text = format_rows_text(ref, cnt, st, self._show_n, self._long, decimal_places, first_pos=a + 1,
                                            threads=threads)
                    if text is None:
                        lines = []
                        for row in _offset_rows(build_rows(ref, cnt, st, self._show_n, self._long), a):
                            lines.append("\t".join([x if isinstance(x, str) else str(round(x, decimal_places)) for x in row]))
                        text = "\n".join(lines)
                    if text:
                        out.write(text)
                        out.write("\n")
                    rows += (b - a) * (k if self._long else 1)
        return rows

    def _index(self, reference):
        if reference not in self.reference_lengths:
            raise Exception(f"{reference} is not a valid reference")
        return self.references.index(reference)

    def rows(self, reference=None):
        """Iterator of row lists; restricted to `reference` if given."""
        if reference is not None:
            self._index(reference)
        for ref in ([reference] if reference is not None else self.references):
            yield from self.data[ref]["rows"]

    def records(self, reference=None):
        """Iterator of dicts keyed by column name; restricted to `reference` if given."""
        for row in self.rows(reference):
            yield dict(zip(self.columns, row))

    def num_reads(self, reference=None):
        """Total reads counted, over all references or for one."""
        if reference is None:
            return sum(self._pile.num_reads)
        return self._pile.num_reads[self._index(reference)]

    def _reduced(self, reference, min_coverage):
        """Device reductions over the chosen references: (positions with coverage >= min_coverage,
        coverage sum over all positions, entropy sum over the selected positions, positions)."""
        which = range(len(self.references)) if reference is None else [self._index(reference)]
        key = int(min_coverage)
        if not hasattr(self, "_red"):
            self._red = {}
        if key not in self._red:
            self._red[key] = self._pile.engine.summary_min_coverage(key, self._show_n)
        sel, cs, es = self._red[key]
        n_sel = sum(int(sel[i]) for i in which)
        cov_sum = sum(int(cs[i]) for i in which)
        ent_sum = np.float64(0.0)
        for i in which:
            ent_sum = ent_sum + np.float64(es[i])
        return n_sel, cov_sum, ent_sum, sum(self._pile.lengths[i] for i in which)

    def mean_coverage(self, reference=None):
        """Mean coverage over all positions (np.mean of the coverage column, main.py:325-340), from the
        device's integer coverage sum: sums of int64 below 2^53 are exact in np.mean's float64 accumulator,
        so this is the reference's value bit for bit."""
        _, cov_sum, _, n = self._reduced(reference, 0)
        if n == 0:
            return np.mean(np.zeros(0, np.int64))          # nan + RuntimeWarning, as the reference
        rep = len(BASES[:6 if self._show_n else 5]) if self._long else 1       # long format repeats each position
        return np.float64(cov_sum * rep) / (n * rep)

    def mean_entropy(self, reference=None, min_coverage=0):
        """Mean entropy over positions with coverage >= min_coverage (main.py:342-359), reduced on the
        device (fixed-order tree; agrees with np.mean's pairwise order to ~1e-15 relative)."""
        n_sel, _, ent_sum, _ = self._reduced(reference, max(int(np.ceil(min_coverage)), 0))
        if n_sel == 0:
            return np.mean(np.zeros(0, np.float64))        # nan + RuntimeWarning, as the reference
        return ent_sum / n_sel

    # -- device-side reductions used by the CLI's summarise modes
    def summary(self, reference):
        """(pc_reference_coverage, avg_depth, avg_entropy) from K3 (main.py:479-485)."""
        i = self._index(reference)
        if not hasattr(self, "_summary"):
            self._summary = self._pile.engine.summary(self._show_n)
        nz, cs, es = self._summary
        L = self._pile.lengths[i]
        return 100 * (int(nz[i]) / L), np.float64(int(cs[i])) / L, np.float64(es[i]) / L

    def amplicon_vectors(self, reference, scheme):
        """The six per-amplicon vectors in print order (main.py:506-551); ints -1 for empty windows."""
        i = self._index(reference)
        lo = [t[2]["inside_start"] for t in scheme]
        hi = [t[2]["inside_end"] for t in scheme]
        if not scheme:
            return [[] for _ in range(6)]
        out, empty = self._pile.engine.amplicons(i, lo, hi, self._show_n)
        return [[-1 if empty[t] else np.float64(out[k, t]) for t in range(len(scheme))] for k in range(6)]


def handle_arg(arg, name, default=None, provided_once=False):
    """`action="append"` values -> one value (must be given once) or the union of lists (main.py:362-375)."""
    if arg is None:
        return default
    if provided_once:
        if len(arg) > 1:
            raise Exception(f"Argument --{name} can only be provided once")
        return arg[0]
    return list({a for a_list in arg for a in a_list})


AMPLICON_VECTOR_NAMES = ("mean_coverage_amplicon_vector", "median_coverage_amplicon_vector",
                         "mean_entropy_amplicon_vector", "median_entropy_amplicon_vector",
                         "mean_secondary_entropy_amplicon_vector", "median_secondary_entropy_amplicon_vector")


def run(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("bam", help="Path to BAM file (an index file is not required)")
    parser.add_argument("-v", "--version", action="version", version=__version__)
    parser.add_argument("--references", default=None, nargs="+", action="append",
                        help="Choose specific reference(s) to run basecount on")
    parser.add_argument("--min-base-quality", default=None, action="append", help="Default value: 0")
    parser.add_argument("--min-mapping-quality", default=None, action="append", help="Default value: 0")
    parser.add_argument("--chunk-size", default=None, action="append",
                        help="Max number of reads loaded into memory and basecounted at a given time. Default value: 1000000")
    parser.add_argument("--show-n-bases", default=False, action="store_true",
                        help="Show counts of 'N' bases from reads, and include them in statistics")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--long-format", default=False, action="store_true",
                       help="Output per-position statistics in long format, instead of the default wide format")
    group.add_argument("--summarise", default=False, action="store_true", help="Output summary statistics")
    group.add_argument("--summarise-with-bed", default=None, action="append", metavar="BED_FILE",
                       help="Output summary statistics and amplicon vectors (calculated using the provided BED file)")
    parser.add_argument("--decimal-places", default=None, action="append", help="Default value: 3")
    args = parser.parse_args(argv)

    references = handle_arg(args.references, "references")
    min_base_quality = int(handle_arg(args.min_base_quality, "min-base-quality", default=0, provided_once=True))
    min_mapping_quality = int(handle_arg(args.min_mapping_quality, "min-mapping-quality", default=0, provided_once=True))
    chunk_size = int(handle_arg(args.chunk_size, "chunk-size", default=1000000, provided_once=True))
    bed = handle_arg(args.summarise_with_bed, "bed", provided_once=True)
    decimal_places = int(handle_arg(args.decimal_places, "decimal_places", default=3, provided_once=True))

    bc = BaseCount(args.bam, references=references, min_base_quality=min_base_quality,
                   min_mapping_quality=min_mapping_quality, chunk_size=chunk_size, show_n_bases=args.show_n_bases,
                   long_format=args.long_format)

    if (not args.summarise) and (bed is None):
        import sys
        # rows stream out window by window (native emitter, byte-identical to str(round(x, d)) per cell, main.py:461)
        bc.write_tsv(sys.stdout, decimal_places)
        return

    for ref in bc.references:
        pc_ref_coverage, avg_coverage, avg_entropy = bc.summary(ref)
        summary_stats = {
            "reference_name": ref,
            "reference_length": round(bc.reference_lengths[ref], decimal_places),
            "num_reads": round(bc.num_reads(ref), decimal_places),
            "pc_reference_coverage": round(pc_ref_coverage, decimal_places),
            "avg_depth": round(avg_coverage, decimal_places),
            "avg_entropy": round(avg_entropy, decimal_places),
        }
        for name, val in summary_stats.items():
            print(name, val, sep="\t")
        if bed is not None:
            scheme = load_scheme(bed)
            vectors = bc.amplicon_vectors(ref, scheme)
            for name, vec in zip(AMPLICON_VECTOR_NAMES, vectors):
                print(name, ", ".join([str(round(x, decimal_places)) for x in vec]) if vec else "-", sep="\t")
