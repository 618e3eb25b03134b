"""Parity of the CUDA counting path (through the C ABI) against the oracle.  Bit-exact."""
import numpy as np
import pytest

from conftest import load_golden
from basecount_b200 import synth
from basecount_b200.records import ReadBatch, select_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from basecount_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def gpu_counts(eng, batches, ref_lens, mbq=0, variant=0, canonical=True):
    from basecount_b200.pack import pack_batches
    eng.set_count_variant(variant)
    eng.begin(ref_lens)
    eng.push(pack_batches(batches, mbq, canonical=canonical))
    eng.sync()
    return [eng.counts(r) for r in range(len(ref_lens))]


def oracle_counts(batch, ref_len, mbq=0):
    from oracle import bcount as obc
    return obc.bcount_flat(ref_len, mbq, batch).astype(np.int64)


def test_reference_kats_through_bcount():
    from basecount_b200.count import bcount
    kats = load_golden("bcount_kats.json.gz")
    for case in kats:
        args = (case["ref_len"], case["min_base_quality"], case["reads"], case["qualities"], case["starts"],
                [[tuple(t) for t in c] for c in case["ctuples"]])
        if "error" in case:
            with pytest.raises(IndexError):
                bcount(*args)
        else:
            assert bcount(*args) == case["counts"]


@pytest.mark.parametrize("variant", [0, 1, 2])      # fast kernel + walker (default), per-base atomics, walker alone
@pytest.mark.parametrize("mbq", [0, 20, 40])
def test_fuzz_vs_oracle(eng, mbq, variant):
    for seed in range(40, 52):
        b = synth.fuzz_batch(seed, n_reads=300, ref_len=700, sorted_by_pos=(seed % 2 == 0))
        got = gpu_counts(eng, b, [700], mbq, variant)[0]
        assert np.array_equal(got, oracle_counts(b, 700, mbq)), seed


@pytest.mark.parametrize("mbq", [0, 20])
def test_bam_native_cigars_vs_oracle(eng, mbq):
    """CIGARs handed over as the BAM holds them (no normal form: S/H/P, =/X, N, empty ops): the fast kernel takes
    the reads that happen to be in its two shapes and defers every other block to the general walker."""
    for seed in range(60, 66):
        b = synth.fuzz_batch(seed, n_reads=300, ref_len=700, sorted_by_pos=(seed % 2 == 0))
        got = gpu_counts(eng, b, [700], mbq, 0, canonical=False)[0]
        assert np.array_equal(got, oracle_counts(b, 700, mbq)), seed
    amp = select_reads(synth.amplicon_sample(seed=77, n_reads=20000, ref_len=6000, ref_name="x"), 0, 0)
    got = gpu_counts(eng, amp, [6000], mbq, 0, canonical=False)[0]       # deferred and fast blocks interleave
    assert np.array_equal(got, oracle_counts(amp, 6000, mbq))


@pytest.mark.parametrize("read_len,n_reads", [(150, 6000), (400, 4000), (700, 2500), (1000, 1500), (5000, 300)])
@pytest.mark.parametrize("mbq", [0, 20])
def test_group_widths_vs_oracle(eng, read_len, n_reads, mbq):
    """150 bp -> 4-lane read slots, 400 bp -> 8, 700 bp -> 16, >= 900 bp -> 32; 5 kb reads span several windows."""
    rec = synth.uniform_short_read_sample(seed=read_len, ref_len=20000, n_reads=n_reads, read_len=read_len, ref_name="x")
    b = select_reads(rec, 0, 0)
    got = gpu_counts(eng, b, [20000], mbq)[0]
    assert np.array_equal(got, oracle_counts(b, 20000, mbq))


@pytest.mark.parametrize("read_len,n_reads", [(150, 6000), (400, 4000), (1000, 1500), (5000, 300)])
def test_walker_alone_vs_oracle(eng, read_len, n_reads):
    """Count variant 2 (the general walker of csrc/k1_count.cuh without the fast kernel in front): every group width,
    quality masks on and off, the sample-batch shape."""
    rec = synth.uniform_short_read_sample(seed=read_len + 1, ref_len=20000, n_reads=n_reads, read_len=read_len, ref_name="x")
    b = select_reads(rec, 0, 0)
    for mbq in (0, 20):
        got = gpu_counts(eng, b, [20000], mbq, variant=2)[0]
        assert np.array_equal(got, oracle_counts(b, 20000, mbq)), mbq
    amp = [select_reads(synth.amplicon_sample(seed=40 + s, n_reads=9000, ref_len=6000, ref_name="x"), 0, 0) for s in range(3)]
    got = gpu_counts(eng, amp, [6000] * 3, 0, variant=2)
    for s in range(3):
        assert np.array_equal(got[s], oracle_counts(amp[s], 6000)), s


def test_amplicon_shapes_and_filter_matrix(eng):
    rec = synth.amplicon_sample(seed=9, n_reads=20000, ref_len=6000, ref_name="x")
    for mmq in (0, 30, 60):
        b = select_reads(rec, 0, mmq)
        for mbq in (0, 20, 40):
            got = gpu_counts(eng, b, [6000], mbq)[0]
            assert np.array_equal(got, oracle_counts(b, 6000, mbq)), (mbq, mmq)


def test_unsorted_and_chunk_invariance(eng):
    from basecount_b200.pack import pack_batches
    rec = synth.deep_short_read_sample(seed=4, n_reads=30000, ref_len=6000, ref_name="x")
    b = select_reads(rec, 0, 0)
    want = oracle_counts(b, 6000, 0)
    perm = np.random.default_rng(0).permutation(b.n)
    assert np.array_equal(gpu_counts(eng, synth.take_batch(b, perm), [6000])[0], want)
    # pushing the reads in several batches accumulates like np.add over chunks (main.py:155)
    for chunk in (1, 7, 1000, 12345):
        eng.begin([6000])
        for a in range(0, b.n, max(chunk, b.n // 40)):
            eng.push(pack_batches(synth.take_batch(b, np.arange(a, min(a + max(chunk, b.n // 40), b.n))), 0))
        eng.sync()
        assert np.array_equal(eng.counts(0), want), chunk


def test_multi_reference_batch(eng):
    lens = [700, 0, 1500, 64, 4097]
    batches, want = [], []
    for r, L in enumerate(lens):
        if L == 0:
            b = ReadBatch.from_lists([], [], [], [])
        else:
            b = synth.fuzz_batch(200 + r, n_reads=150 + 50 * r, ref_len=L, sorted_by_pos=True)
        batches.append(b)
        want.append(oracle_counts(b, L, 20))
    got = gpu_counts(eng, batches, lens, 20)
    for r in range(len(lens)):
        assert np.array_equal(got[r], want[r]), r


def test_index_error_semantics(eng):
    """IndexError only when an INCREMENT lands at refPos >= refLen (count.cpp .at())."""
    from basecount_b200.count import bcount
    q = [[30] * 4]
    with pytest.raises(IndexError):
        bcount(6, 0, ["ACGT"], q, [4], [[(0, 4)]])
    with pytest.raises(IndexError):
        bcount(6, 0, ["ACGN"], q, [3], [[(0, 4)]])            # the N past the end counts -> throws
    assert bcount(6, 0, ["ACRY"], q, [4], [[(0, 4)]])[5][1] == 1   # uncounted letters past the end: no error
    assert bcount(6, 31, ["ACGT"], q, [4], [[(0, 4)]])[4] == [0] * 6    # all filtered: no error
    with pytest.raises(IndexError):
        bcount(6, 31, ["ACGT"], q, [4], [[(0, 1), (2, 3)]])   # deletions are exempt from the filter
    with pytest.raises(TypeError):
        bcount(-1, 0, [], [], [], [])
    with pytest.raises(TypeError):
        bcount(6, 0, [None], q, [0], [[(0, 1)]])
    # the engine is usable again after an error
    assert bcount(3, 0, ["AC"], [[1, 1]], [1], [[(0, 2)]]) == [[0] * 6, [1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0]]


def test_config1_full_size_vs_oracle(eng):
    """BASELINE config 1/2 at full size: 124k reads x 400 bp on the 29,903 bp genome."""
    rec = synth.amplicon_sample(seed=1)
    b = select_reads(rec, 0, 0)
    assert b.n == 124000 and b.aligned_bases() == 49_600_000
    got = gpu_counts(eng, b, [synth.SARS2_LEN])[0]
    assert np.array_equal(got, oracle_counts(b, synth.SARS2_LEN))
    # size-independent property: every aligned base lands in exactly one cell, unless it is an
    # uncounted letter (none here: only ACGTN are generated)
    assert int(got.sum()) == b.aligned_bases()


def _reads_to_batch(reads):
    """reads: list of (start, [(op, len), ...]); random ACGT bases, quality 30."""
    rng = np.random.default_rng(len(reads))
    seqs, quals, starts, cts = [], [], [], []
    for start, ct in reads:
        qn = sum(l for o, l in ct if o in (0, 1, 7, 8))
        seqs.append("".join("ACGT"[i] for i in rng.integers(0, 4, size=qn)))
        quals.append([30] * qn)
        starts.append(start)
        cts.append(ct)
    return ReadBatch.from_lists(seqs, quals, starts, cts)


def test_long_reads_inside_short_read_batches(eng):
    """Reads far longer than the batch mean are not covered by the TMA stage of their block: they are
    copied into a free stage segment by segment (20 kb = several segments), with indels inside."""
    L = 60000
    rng = np.random.default_rng(3)
    reads = []
    for i in range(3000):
        s = int(rng.integers(0, L - 200))
        reads.append((s, [(0, 150)]))
        if i % 500 == 250:
            s2 = int(rng.integers(0, L - 26000))
            reads.append((s2, [(4, 10), (0, 9000), (1, 37), (0, 3000), (2, 5), (0, 8000), (3, 700), (8, 4000), (4, 3)]))
        if i % 700 == 350:
            reads.append((int(rng.integers(0, L - 3000)), [(0, 1400), (2, 1), (7, 1400)]))
    reads.sort(key=lambda r: r[0])
    b = _reads_to_batch(reads)
    for mbq in (0, 31):
        got = gpu_counts(eng, b, [L], mbq)[0]
        assert np.array_equal(got, oracle_counts(b, L, mbq)), mbq


def test_long_skips_many_ops_and_unsorted_windows(eng):
    """N skips longer than a window (spliced reads), D runs beyond the single-lane limit, reads with more
    than three ops next to plain ones, and a position order that forces windows to move backwards."""
    L = 50000
    rng = np.random.default_rng(11)
    reads = []
    for i in range(4000):
        s = int(rng.integers(0, L - 12000))
        kind = i % 8
        if kind == 0:
            reads.append((s, [(0, 60), (3, int(rng.integers(33, 9000))), (0, 90)]))
        elif kind == 1:
            reads.append((s, [(0, 40), (2, int(rng.integers(30, 200))), (0, 50), (1, 3), (0, 57)]))
        elif kind == 2:
            reads.append((s, [(4, 5), (7, 30), (8, 1), (7, 30), (1, 2), (0, 20), (2, 2), (0, 40), (5, 9)]))
        elif kind == 3:
            reads.append((s, [(3, 40), (0, 150)]))              # leading skip: run B only
        else:
            reads.append((s, [(0, 150)]))
    for order in ("sorted", "reverse", "random"):
        rs = sorted(reads, key=lambda r: r[0])
        if order == "reverse":
            rs = rs[::-1]
        elif order == "random":
            rs = [rs[i] for i in rng.permutation(len(rs))]
        b = _reads_to_batch(rs)
        got = gpu_counts(eng, b, [L])[0]
        assert np.array_equal(got, oracle_counts(b, L)), order


def test_fuzz_with_alignments_past_the_reference_end(eng):
    """Batches where some alignments run past ref_len: the GPU path must raise IndexError exactly when
    the reference does (count.cpp .at()), and agree on the counts otherwise."""
    from basecount_b200.pack import pack_batches
    raised = clean = 0
    for seed in range(300, 360):
        b = synth.fuzz_batch(seed, n_reads=5, ref_len=400, allow_overflow=True, sorted_by_pos=(seed % 2 == 0),
                             long_op_frac=0.02)
        for mbq in (0, 30):
            try:
                want = oracle_counts(b, 400, mbq)
            except IndexError:
                want = None
            eng.begin([400])
            eng.push(pack_batches(b, mbq))
            if want is None:
                with pytest.raises(IndexError):
                    eng.sync()
                raised += 1
            else:
                eng.sync()
                assert np.array_equal(eng.counts(0), want), (seed, mbq)
                clean += 1
    assert raised > 5 and clean > 5
