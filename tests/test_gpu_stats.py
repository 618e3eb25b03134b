"""Parity of K2 (per-position statistics) and K3 (summary, amplicon vectors) against the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from basecount_b200 import synth
from basecount_b200.records import ReadBatch, select_reads

pytestmark = pytest.mark.gpu
REL = 1e-6          # BASELINE.json north_star: pc_* and entropy within 1e-6 relative before rounding
TIGHT = 1e-12       # what we actually hold (only log2's last ulp differs)


@pytest.fixture(scope="module")
def eng():
    from basecount_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def load_counts(eng, counts):
    """Put an arbitrary count matrix on the device by pushing synthetic single-base reads."""
    from basecount_b200.pack import pack_batches
    counts = np.asarray(counts, dtype=np.int64)
    L = counts.shape[0]
    reads, quals, starts, ctuples = [], [], [], []
    letters = "ACGT?N"
    for pos in range(L):
        for col in range(6):
            k = int(counts[pos, col])
            if k == 0:
                continue
            if col == 4:
                reads += [""] * k
                quals += [[]] * k
                starts += [pos] * k
                ctuples += [[(2, 1)]] * k
            else:
                reads += [letters[col]] * k
                quals += [[30]] * k
                starts += [pos] * k
                ctuples += [[(0, 1)]] * k
    eng.begin([L])
    eng.push(pack_batches(ReadBatch.from_lists(reads, quals, starts, ctuples), 0))
    eng.sync()
    assert np.array_equal(eng.counts(0), counts)


def small_golden_counts():
    g = load_golden("stats.json.gz")
    return [c for c in g["counts"] if sum(c) < 5000]


@pytest.mark.parametrize("show_n", [False, True])
def test_stats_vs_oracle(eng, show_n):
    from oracle import stats as ost
    counts = small_golden_counts()
    load_counts(eng, counts)
    s = eng.stats(0, show_n)
    k = 6 if show_n else 5
    for pos, row in enumerate(counts):
        cov, c, pcs, ent, sec = ost.position_stats(row, show_n)
        assert int(s["coverage"][pos]) == cov
        f = int(s["flags"][pos])
        assert bool(f & 1) == (cov == 0)
        assert bool(f & 2) == isinstance(sec, int)
        for i in range(k):
            assert s["pc"][i, pos] == float(pcs[i])                 # divisions are IEEE-exact: bit-identical
        assert s["entropy"][pos] == pytest.approx(float(ent), rel=TIGHT, abs=1e-300)
        assert s["secondary"][pos] == pytest.approx(float(sec), rel=TIGHT, abs=1e-300)
        assert abs(s["entropy"][pos] - float(ent)) <= REL * abs(float(ent))


def test_summary_and_amplicons_vs_oracle(eng, tmp_path):
    from basecount_b200.pack import pack_batches
    from oracle import bcount as obc
    from oracle import stats as ost
    rec = synth.amplicon_sample(seed=21, n_reads=900, ref_len=3000, ref_name="toy")
    b = select_reads(rec, 0, 0)
    counts = obc.bcount_flat(3000, 0, b).astype(np.int64).tolist()
    eng.begin([3000])
    eng.push(pack_batches(b, 0))
    eng.sync()
    for show_n in (False, True):
        cov, ent, sec = ost.per_position_vectors(counts, show_n)
        nz, cs, es = eng.summary(show_n)
        assert int(nz[0]) == sum(1 for x in cov if x != 0)
        assert int(cs[0]) == sum(cov)
        assert float(es[0]) == pytest.approx(float(np.sum(ent)), rel=TIGHT)
        windows = [(-5, 10), (0, 0), (100, 99), (2990, 5000), (3000, 3100), (50, 449), (300, 700), (0, 2999), (17, 18)]
        want = ost.amplicon_vectors(cov, ent, sec, windows)
        got, empty = eng.amplicons(0, [w[0] for w in windows], [w[1] for w in windows], show_n)
        for t, (lo, hi) in enumerate(windows):
            is_empty = isinstance(want[0][t], int)
            assert bool(empty[t]) == is_empty, (lo, hi)
            for k in range(6):
                if is_empty:
                    assert got[k, t] == -1.0
                elif k < 2:
                    assert got[k, t] == float(want[k][t])          # coverage mean / median: exact
                else:
                    assert got[k, t] == pytest.approx(float(want[k][t]), rel=TIGHT, abs=1e-300)


def test_amplicon_window_larger_than_staging(eng):
    from basecount_b200.pack import pack_batches
    from oracle import bcount as obc
    from oracle import stats as ost
    rec = synth.uniform_short_read_sample(seed=8, ref_len=12000, n_reads=3000, read_len=150, ref_name="x")
    b = select_reads(rec, 0, 0)
    counts = obc.bcount_flat(12000, 0, b).astype(np.int64).tolist()
    eng.begin([12000])
    eng.push(pack_batches(b, 0))
    eng.sync()
    cov, ent, sec = ost.per_position_vectors(counts)
    windows = [(0, 11999), (100, 9000), (5, 4100), (0, 4095)]
    want = ost.amplicon_vectors(cov, ent, sec, windows)
    got, _ = eng.amplicons(0, [w[0] for w in windows], [w[1] for w in windows])
    for t in range(len(windows)):
        for k in range(6):
            if k in (1, 3, 5):
                assert got[k, t] == float(want[k][t])               # medians are exact selections
            else:
                assert got[k, t] == pytest.approx(float(want[k][t]), rel=TIGHT)


def test_queued_results_are_delivered_at_sync(eng):
    """bc_summary_async / bc_amplicons_async park their results in a device arena; several queued calls
    (more than the arena's first allocation holds between two syncs is exercised by the tile count) all
    land in the callers' arrays at the next sync and equal the synchronous calls."""
    from basecount_b200.pack import pack_batches
    rec = synth.amplicon_sample(seed=33, n_reads=1200, ref_len=5000, ref_name="toy")
    b = select_reads(rec, 0, 0)
    eng.begin([5000])
    eng.push(pack_batches(b, 0))
    eng.sync()
    lo = list(range(0, 4900, 7))
    hi = [x + 399 for x in lo]
    want, want_empty = eng.amplicons(0, lo, hi)
    want_sum = eng.summary(False)
    queued = []
    for _ in range(40):                                   # 40 x (700 windows x 49 B + 24 B) > 1 MB: forces an early delivery
        out = np.full((6, len(lo)), np.nan)
        empty = np.full(len(lo), 7, dtype=np.uint8)
        s = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
        eng.summary_async(s, False)
        eng.amplicons_async(0, lo, hi, out, empty)
        queued.append((out, empty, s))
    eng.sync()
    for out, empty, s in queued:
        assert np.array_equal(out, want) and np.array_equal(empty, want_empty)
        assert int(s[0][0]) == int(want_sum[0][0]) and int(s[1][0]) == int(want_sum[1][0]) and float(s[2][0]) == float(want_sum[2][0])


@pytest.mark.parametrize("shape", ["shallow", "medium", "deep"])
def test_summarise_entropy_paths_vs_oracle(eng, shape):
    """k2_summary takes its entropy terms from three places: a table of exact terms (coverage < 128), the
    exact log2 for the largest class plus a log2 table for the others (coverage < 16384), and the exact
    expression for everything (deeper).  Each against the oracle's get_stats restatement, and the
    coverage-threshold reduction behind BaseCount.mean_entropy against numpy on the oracle's vectors."""
    from basecount_b200.pack import pack_batches
    from oracle import bcount as obc
    from oracle import stats as ost
    if shape == "shallow":
        rec = synth.uniform_short_read_sample(seed=41, ref_len=6000, n_reads=1200, read_len=150, ref_name="s")
        L = 6000
    elif shape == "medium":
        rec = synth.amplicon_sample(seed=42, n_reads=4000, ref_len=1500, ref_name="m")
        L = 1500
    else:
        rec = synth.uniform_short_read_sample(seed=43, ref_len=400, n_reads=60000, read_len=150, ref_name="d")
        L = 400
    b = select_reads(rec, 0, 0)
    counts = obc.bcount_flat(L, 0, b).astype(np.int64)
    cmax = int(counts[:, :5].sum(axis=1).max())
    assert {"shallow": cmax < 128, "medium": 128 < cmax < 16384, "deep": cmax > 16384}[shape], cmax
    eng.begin([L])
    eng.push(pack_batches(b, 0))
    eng.sync()
    for show_n in (False, True):
        cov, ent, _ = ost.per_position_vectors(counts.tolist(), show_n)
        nz, cs, es = eng.summary(show_n)
        assert int(nz[0]) == sum(1 for x in cov if x != 0) and int(cs[0]) == sum(cov)
        assert float(es[0]) == pytest.approx(float(np.sum(np.asarray(ent, dtype=np.float64))), rel=TIGHT)
        cov_a, ent_a = np.asarray(cov), np.asarray(ent, dtype=np.float64)
        for thr in (0, 1, int(np.median(cov_a)), cmax, cmax + 1):
            sel, cs2, es2 = eng.summary_min_coverage(thr, show_n)
            keep = cov_a >= thr
            assert int(sel[0]) == int(keep.sum()) and int(cs2[0]) == int(cov_a.sum())
            assert float(es2[0]) == pytest.approx(float(ent_a[keep].sum()), rel=TIGHT, abs=1e-300)


def test_pipelined_summaries_see_their_own_step(eng):
    """bc_summary_async runs on its own stream, beside the counting kernel of the NEXT step (which, after
    bc_reset, writes the other set of accumulators).  A pipeline of steps over three different batches with no
    synchronisation in between must hand every step the summary of ITS batch; a push with no reset in between
    (same accumulators) must wait for the queued summary; a length change must wait for it too."""
    from basecount_b200.pack import pack_batches
    L = 4000
    batches = [select_reads(synth.amplicon_sample(seed=60 + k, n_reads=900 + 400 * k, ref_len=L, ref_name="p"), 0, 0)
               for k in range(3)]
    packed = [pack_batches(b, 0) for b in batches]
    want = []
    for p in packed:                                      # synchronous references, one batch at a time
        eng.begin([L])
        eng.push(p)
        eng.sync()
        want.append(tuple(np.array(a).copy() for a in eng.summary(False)))
    eng.begin([L])
    outs = []
    for i in range(12):
        s = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
        eng.reset()
        eng.push(packed[i % 3])
        eng.summary_async(s, False)
        outs.append(s)
    eng.sync()
    for i, s in enumerate(outs):
        w = want[i % 3]
        assert int(s[0][0]) == int(w[0][0]) and int(s[1][0]) == int(w[1][0]) and float(s[2][0]) == float(w[2][0]), i
    # no reset between the queued summary and the next push: the summary must not see the second batch
    eng.begin([L])
    eng.push(packed[0])
    s0 = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
    eng.summary_async(s0, False)
    eng.push(packed[1])
    s01 = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
    eng.summary_async(s01, False)
    eng.sync()
    assert int(s0[1][0]) == int(want[0][1][0]) and float(s0[2][0]) == float(want[0][2][0])
    assert int(s01[1][0]) == int(want[0][1][0]) + int(want[1][1][0])
    # a device-side length change behind a queued summary: the summary still covers the old length
    eng.begin([L])
    eng.push(packed[2])
    s2 = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
    eng.summary_async(s2, False)
    eng.set_length(0, L // 2)
    half = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.float64))
    eng.summary_async(half, False)
    eng.sync()
    assert int(s2[1][0]) == int(want[2][1][0]) and float(s2[2][0]) == float(want[2][2][0])
    assert int(half[1][0]) == int(eng.counts(0)[:, :5].sum()) < int(want[2][1][0])
    eng.set_length(0, L)
    eng.sync()
