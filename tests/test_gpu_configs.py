"""BASELINE.json's configs 3, 4 and 5 through the CUDA path (C ABI), at full size.

Config 1/2 at full size is tests/test_gpu_counts.py::test_config1_full_size_vs_oracle.  Here:

  * config 3 (2 M reads x 150 bp, ~10,000x, --summarise-with-bed): counts bit-exact against the C oracle
    (300 M aligned bases take it about two seconds), the summary scalars and the six amplicon vectors of
    the 98-amplicon scheme against the oracle's restatement of main.py:479-485,519-551;
  * config 4 (a batch of independent samples, one launch): every sample's counts and summary against the
    oracle run sample by sample -- samples never leak into each other -- and the sample shards of
    dist.shard_samples cover the batch;
  * config 5 (64.4 Mb, 12.9 M reads x 150 bp, region-sharded): a 1/8 region is checked bit-exact against
    the oracle, single pass and region-sharded 2, 4 and 8 ways (bc_halo_export / bc_halo_add / bc_truncate,
    the calls dist.count_region_sharded makes on each rank); the FULL-size reference is that region
    translated 8 times, so its counts must be 8 copies of the region's (translation invariance: a
    size-independent property that pins every cell of the 64.4 M x 6 matrix) and its summary sums 8 times
    the region's.
"""
import numpy as np
import pytest

from basecount_b200 import dist as bdist
from basecount_b200 import synth
from basecount_b200.records import ReadBatch, select_reads

pytestmark = pytest.mark.gpu
TIGHT = 1e-12
DEVICE = "cuda:0"                      # halo buffers of the sharded runs
CFG3_READS = 2_000_000


@pytest.fixture(scope="module")
def eng():
    from basecount_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def oracle_counts(batch, ref_len, mbq=0):
    from oracle import bcount as obc
    return obc.bcount_flat(ref_len, mbq, batch).astype(np.int64)


def entropy_sum_numpy(counts):
    """Sum over positions of the normalised entropy (main.py:10-11,24,42; zero coverage -> 1, main.py:35),
    vectorised for matrices too long for the oracle's per-position Python loop.  Test-side only."""
    c = counts[:, :5].astype(np.float64)
    cov = c.sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = c / cov[:, None]
        t = np.where(c > 0, -(p * np.log2(p)), 0.0)
    ent = np.where(cov > 0, (1 / np.log2(5)) * t.sum(axis=1), 1.0)
    return float(ent.sum()), int((cov > 0).sum()), int(counts[:, :5].sum())


# ----------------------------------------------------------------------------- config 3
def test_config3_full_size(eng, tmp_path):
    from basecount_b200.pack import pack_batches
    from oracle import stats as ost
    L = synth.SARS2_LEN
    b = select_reads(synth.deep_short_read_sample(seed=3, n_reads=CFG3_READS), 0, 0)
    assert b.n == CFG3_READS and b.aligned_bases() == 150 * CFG3_READS
    want = oracle_counts(b, L)
    eng.set_count_variant(0)
    eng.begin([L])
    eng.push(pack_batches(b, 0))
    eng.sync()
    got = eng.counts(0)
    assert np.array_equal(got, want)
    assert int(got.sum()) == b.aligned_bases()            # every aligned base lands in exactly one cell

    cov, ent, sec = ost.per_position_vectors(want.tolist())
    nz, cs, es = eng.summary(False)
    assert int(nz[0]) == sum(1 for x in cov if x != 0) and int(cs[0]) == sum(cov)
    assert float(es[0]) == pytest.approx(float(np.sum(ent)), rel=TIGHT)

    bed = str(tmp_path / "scheme.bed")
    synth.artic_like_bed(bed)
    tiles = ost.scheme_windows(bed)
    assert len(tiles) == 98
    windows = [(t[2]["inside_start"], t[2]["inside_end"]) for t in tiles]
    wantv = ost.amplicon_vectors(cov, ent, sec, windows)
    gotv, empty = eng.amplicons(0, [w[0] for w in windows], [w[1] for w in windows])
    assert not empty.any()
    for t in range(len(windows)):
        for k in range(6):
            if k < 2:
                assert gotv[k, t] == float(wantv[k][t]), (k, t)   # coverage mean / median: exact
            else:
                assert gotv[k, t] == pytest.approx(float(wantv[k][t]), rel=TIGHT, abs=1e-300), (k, t)


# ----------------------------------------------------------------------------- config 4
def test_config4_sample_batch(eng):
    """12 samples (one GPU's share of the 96) in one batch and one K1 launch."""
    from basecount_b200.pack import pack_batches
    from oracle import stats as ost
    L = synth.SARS2_LEN
    n_samples = 12
    batches = [select_reads(synth.amplicon_sample(seed=100 + s, n_reads=6000 + 500 * s), 0, 0) for s in range(n_samples)]
    eng.set_count_variant(0)
    eng.begin([L] * n_samples)
    eng.push(pack_batches(batches, 0))
    eng.sync()
    nz, cs, es = eng.summary(False)
    for s, b in enumerate(batches):
        want = oracle_counts(b, L)
        assert np.array_equal(eng.counts(s), want), s
        cov, ent, _ = ost.per_position_vectors(want.tolist())
        assert int(nz[s]) == sum(1 for x in cov if x != 0) and int(cs[s]) == sum(cov)
        assert float(es[s]) == pytest.approx(float(np.sum(ent)), rel=TIGHT)
    shards = [bdist.shard_samples(96, 8, r) for r in range(8)]
    assert sorted(i for sh in shards for i in sh) == list(range(96)) and all(len(sh) == n_samples for sh in shards)


# ----------------------------------------------------------------------------- config 5
REGION = 8_055_520                     # 8 regions = 64,444,160 of chr20's 64,444,167 columns
REGION_READS = 12_888_833 // 8


@pytest.fixture(scope="module")
def region():
    rec = synth.uniform_short_read_sample(seed=5, ref_len=REGION, n_reads=REGION_READS)
    b = select_reads(rec, 0, 0)
    return b, oracle_counts(b, REGION)


def _count_sharded(eng, b, ref_len, world):
    """What dist.count_region_sharded does on every rank, run rank after rank on one device: halos only
    travel to higher ranks, so the exports of rank q are ready when rank s > q adds them."""
    import torch
    from basecount_b200.pack import pack_batches
    bounds = bdist.region_bounds(ref_len, world)
    dev = torch.device(DEVICE)
    inbox = {s: [] for s in range(world)}
    out = np.zeros((ref_len, 6), dtype=np.int64)
    nz_t = cs_t = 0
    es_t = 0.0
    reads = 0
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        local = bdist.select_region(b, lo, hi)
        reads += local.n
        h = bdist.halo_columns(local, hi - lo, ref_len - hi)
        eng.begin([hi - lo + h])
        eng.push(pack_batches(local, 0))
        eng.sync()
        for s in range(r + 1, world):                     # dist.exchange_halos, sending side
            a, e = max(int(bounds[s]), hi), min(int(bounds[s + 1]), hi + h)
            if a < e:
                t = torch.empty(6 * (e - a), dtype=torch.int32, device=dev)
                eng.halo_export(0, a - lo, e - a, t.data_ptr())
                inbox[s].append((a - int(bounds[s]), e - a, t))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)               # the exports ran on the engine's stream; begin() reuses the slot
        for col, n, t in inbox[r]:                        # receiving side
            eng.halo_add(0, col, n, t.data_ptr())
        eng.truncate(0, hi - lo)
        out[lo:hi] = eng.counts(0)
        nz, cs, es = eng.summary(False)
        nz_t += int(nz[0])
        cs_t += int(cs[0])
        es_t += float(es[0])
    assert reads == b.n                                   # every read is counted on exactly one rank
    return out, (nz_t, cs_t, es_t)


def test_config5_region_single_pass_and_sharded(eng, region):
    from basecount_b200.pack import pack_batches
    b, want = region
    assert b.aligned_bases() == int(want.sum())
    eng.set_count_variant(0)
    eng.begin([REGION])
    eng.push(pack_batches(b, 0))
    eng.sync()
    assert np.array_equal(eng.counts(0), want)
    es_w, nz_w, cs_w = entropy_sum_numpy(want)
    nz, cs, es = eng.summary(False)
    assert int(nz[0]) == nz_w and int(cs[0]) == cs_w
    assert float(es[0]) == pytest.approx(es_w, rel=1e-9)  # numpy's log2 and summation order on the test side
    single = float(es[0])
    for world in (2, 4, 8):
        got, (nz_t, cs_t, es_t) = _count_sharded(eng, b, REGION, world)
        assert np.array_equal(got, want), world
        assert nz_t == nz_w and cs_t == cs_w
        assert es_t == pytest.approx(single, rel=TIGHT)   # the all-reduced scalars of summary_region_sharded


def test_config5_full_size_translation_invariance(eng, region):
    from basecount_b200.pack import pack_batches
    b, want = region
    copies, L = 8, synth.CHR20_LEN
    assert copies * REGION <= L
    starts = np.concatenate([b.starts.astype(np.int64) + k * REGION for k in range(copies)]).astype(np.uint32)
    nc, ns = int(b.cigar_off[-1]), int(b.seq_off[-1])
    cigar_off = np.concatenate([b.cigar_off[:-1].astype(np.int64) + k * nc for k in range(copies)] + [[copies * nc]])
    seq_off = np.concatenate([b.seq_off[:-1].astype(np.int64) + k * ns for k in range(copies)] + [[copies * ns]])
    full = ReadBatch(starts, np.tile(b.cigar, copies), cigar_off.astype(b.cigar_off.dtype), np.tile(b.seq, copies),
                     np.tile(b.qual, copies), seq_off.astype(b.seq_off.dtype))
    assert full.n == copies * b.n and full.aligned_bases() == copies * b.aligned_bases()
    assert full.n == 8 * REGION_READS and full.aligned_bases() == 150 * full.n
    eng.set_count_variant(0)
    eng.begin([L])
    eng.push(pack_batches(full, 0))
    eng.sync()
    nz, cs, es = eng.summary(False)
    got = eng.counts(0)
    for k in range(copies):
        assert np.array_equal(got[k * REGION:(k + 1) * REGION], want), k
    assert not got[copies * REGION:].any()
    del got
    es_w, nz_w, cs_w = entropy_sum_numpy(want)
    assert int(nz[0]) == copies * nz_w and int(cs[0]) == copies * cs_w      # (coverage leaves the N column out)
    assert float(es[0]) == pytest.approx(copies * es_w + (L - copies * REGION), rel=1e-9)
