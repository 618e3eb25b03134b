"""Property tests over random reads (hypothesis), CPU only.

  * the C oracle against the compiled, unmodified reference operator (oracle/_ref) on arbitrary CIGARs over all
    ten op codes, arbitrary letters and qualities, every quality threshold - including WHEN it raises IndexError;
  * the packed batch (the device's input contract: 2-bit planes, quality/ACGT mask, sparse exception list,
    BAM-native CIGAR words) carries everything the counts need: a plain-Python restatement of what K1 and its
    correction pass read from the packed arrays gives the oracle's matrix;
  * linearity: counting a batch in any split gives the same matrix (the chunk invariance of main.py:142-162);
  * region sharding (dist.region_bounds / select_region / halo_columns): shards plus halo columns add up to the
    single pass for every world size.
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from basecount_b200 import dist as bdist
from basecount_b200.pack import pack_batches
from basecount_b200.records import ReadBatch
from basecount_b200.synth import concat_batches, take_batch
from oracle import bcount as obc

REF_LEN = 96
LETTERS = "ACGTNRYKacgtn*="
SETTINGS = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@st.composite
def reads(draw, max_reads=6, ref_len=REF_LEN, may_overrun=True):
    """(reads, qualities, starts, ctuples) as count.bcount takes them; the CIGAR never consumes more bases than
    the read holds (undefined behaviour in the reference, count.cpp:56,58)."""
    n = draw(st.integers(0, max_reads))
    out = ([], [], [], [])
    for _ in range(n):
        ops = draw(st.lists(st.tuples(st.integers(0, 9), st.integers(0, 40)), min_size=0, max_size=7))
        if not may_overrun:                               # keep the alignment inside the reference
            while sum(l for o, l in ops if o in (0, 2, 3, 7, 8)) > ref_len:
                ops = ops[:-1]
        query = sum(l for o, l in ops if o in (0, 1, 7, 8))
        span = sum(l for o, l in ops if o in (0, 2, 3, 7, 8))
        slack = draw(st.integers(0, 3))
        seq = "".join(draw(st.lists(st.sampled_from(LETTERS), min_size=query + slack, max_size=query + slack)))
        qual = draw(st.lists(st.integers(0, 60), min_size=len(seq), max_size=len(seq)))
        hi = ref_len + 5 if may_overrun else max(ref_len - span, 0)
        out[0].append(seq)
        out[1].append(qual)
        out[2].append(draw(st.integers(0, hi)))
        out[3].append(ops)
    return out


_REF = []


def compiled_reference():
    if not _REF:
        _REF.append(obc.load_ref_bcount())
    return _REF[0]


def oracle_or_error(ref_len, mbq, batch):
    try:
        return obc.bcount_flat(ref_len, mbq, batch)
    except IndexError:
        return "IndexError"


@settings(max_examples=200, **SETTINGS)
@given(reads(), st.sampled_from([0, 1, 20, 41, 61]))
def test_oracle_matches_compiled_reference(r, mbq):
    ref = compiled_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    batch = ReadBatch.from_lists(*r)
    got = oracle_or_error(REF_LEN, mbq, batch)
    try:
        want = np.asarray(ref(REF_LEN, mbq, *batch.to_lists()), dtype=np.uint32).reshape(REF_LEN, 6)
    except IndexError:
        want = "IndexError"
    if isinstance(want, str) or isinstance(got, str):
        assert isinstance(want, str) and isinstance(got, str)
    else:
        assert np.array_equal(got, want)


def counts_from_packed(p, ref_len):
    """What the device reads: walk the BAM-native CIGAR words, take 2-bit codes from the bit planes, gate them by
    the mask when there is one, then apply the sparse corrections (flag 2: undo the provisional A, flag 1: count N)."""
    counts = np.zeros((ref_len, 6), dtype=np.int64)
    exc = {}
    for r, q in zip(p.exc_read.tolist(), p.exc_pos.tolist()):
        exc[(r, q >> 2)] = q & 3
    for i in range(p.n_reads):
        w0, w1 = int(p.seq_woff[i]), int(p.seq_woff[i + 1])
        words = p.planes[w0:w1]
        ok_words = p.okmask[w0:w1] if p.okmask is not None else None
        ref_pos, read_pos = int(p.starts[i]), 0
        for c in range(int(p.cigar_off[i]), int(p.cigar_off[i + 1])):
            op, ln = int(p.cigar[c]) & 0xF, int(p.cigar[c]) >> 4
            if op in (0, 7, 8):
                for j in range(ln):
                    rp, col = read_pos + j, ref_pos + j
                    w, b = rp >> 5, rp & 31
                    word = int(words[w])
                    code = ((word >> b) & 1) | (((word >> (32 + b)) & 1) << 1)
                    ok = True if ok_words is None else bool((int(ok_words[w]) >> b) & 1)
                    flags = exc.get((i, rp), 0)
                    if ok and col < ref_len:
                        counts[col, code] += 1
                    if col < ref_len:
                        if flags & 2:
                            counts[col, 0] -= 1
                        if flags & 1:
                            counts[col, 5] += 1
                ref_pos += ln
                read_pos += ln
            elif op == 1:
                read_pos += ln
            elif op in (2, 3):
                for j in range(ln):
                    if ref_pos + j < ref_len:
                        counts[ref_pos + j, 4] += 1
                ref_pos += ln
    return counts


@settings(max_examples=200, **SETTINGS)
@given(reads(may_overrun=False), st.sampled_from([0, 1, 20, 41]))
def test_packed_batch_carries_the_counts(r, mbq):
    batch = ReadBatch.from_lists(*r)
    want = obc.bcount_flat(REF_LEN, mbq, batch).astype(np.int64)
    p = pack_batches(batch, mbq)
    assert np.array_equal(counts_from_packed(p, REF_LEN), want)
    assert p.aligned_bases == batch.aligned_bases()


@settings(max_examples=150, **SETTINGS)
@given(reads(max_reads=10, may_overrun=False), st.sampled_from([0, 20]), st.data())
def test_counting_is_linear_in_the_reads(r, mbq, data):
    batch = ReadBatch.from_lists(*r)
    whole = obc.bcount_flat(REF_LEN, mbq, batch)
    perm = np.asarray(data.draw(st.permutations(list(range(batch.n)))), dtype=np.int64)
    cut = data.draw(st.integers(0, batch.n))
    a, b = take_batch(batch, perm[:cut]), take_batch(batch, perm[cut:])
    assert np.array_equal(obc.bcount_flat(REF_LEN, mbq, a) + obc.bcount_flat(REF_LEN, mbq, b), whole)
    assert np.array_equal(obc.bcount_flat(REF_LEN, mbq, concat_batches([b, a])), whole)


@settings(max_examples=150, **SETTINGS)
@given(reads(max_reads=12, may_overrun=False), st.integers(1, 8))
def test_region_shards_and_halos_add_up(r, world):
    batch = ReadBatch.from_lists(*r)
    whole = obc.bcount_flat(REF_LEN, 0, batch).astype(np.int64)
    bounds = bdist.region_bounds(REF_LEN, world)
    merged = np.zeros_like(whole)
    n = 0
    for rank in range(world):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        local = bdist.select_region(batch, lo, hi)
        n += local.n
        h = bdist.halo_columns(local, hi - lo, REF_LEN - hi)
        part = obc.bcount_flat(hi - lo + h, 0, local)          # raises if the halo were too small
        merged[lo:hi + h] += part
    assert n == batch.n and np.array_equal(merged, whole)


# ----------------------------------------------------------------------------- statistics, against the reference live
def reference_get_stats():
    """get_stats of the unmodified reference (main.py:14-79), imported from where it lies when this container has
    it; `import pysam` / `from count import bcount` (main.py:2,5) are satisfied by stand-ins for the import only."""
    import os
    import sys
    import types
    if not os.path.exists("/root/reference/basecount/main.py"):
        return None
    saved = {k: sys.modules.get(k) for k in ("pysam", "count")}
    stub = types.ModuleType("pysam")
    stub.set_verbosity = lambda v: 0
    cnt = types.ModuleType("count")
    cnt.bcount = compiled_reference()
    sys.modules["pysam"], sys.modules["count"] = stub, cnt
    sys.path.insert(0, "/root/reference")
    try:
        import basecount.main as refmain
        fn = refmain.get_stats
    finally:
        sys.path.remove("/root/reference")
        for k in [m for m in sys.modules if m == "basecount" or m.startswith("basecount.")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return fn


_GET_STATS = []

count_rows = st.lists(st.lists(st.one_of(st.integers(0, 6), st.integers(0, 3000), st.integers(0, 2 ** 40)),
                               min_size=6, max_size=6), min_size=1, max_size=8)


@settings(max_examples=300, **SETTINGS)
@given(count_rows, st.booleans(), st.booleans())
def test_stats_oracle_matches_reference_get_stats(rows6, show_n, long_format):
    """oracle/stats.py against the reference's own get_stats on arbitrary count rows: same values bit for bit
    and the same Python types (the int sentinels -1 / 1 / 1 of zero-coverage positions print differently)."""
    from conftest import typed_equal
    from oracle import stats as ost
    if not _GET_STATS:
        _GET_STATS.append(reference_get_stats())
    get_stats = _GET_STATS[0]
    if get_stats is None:
        pytest.skip("/root/reference is not mounted here")
    want = get_stats([list(r) for r in rows6], "ref", show_n, long_format)      # (get_stats pops from its input)
    got = ost.rows(rows6, "ref", show_n, long_format)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert typed_equal(a, b), (a, b)
