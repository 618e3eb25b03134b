"""Host-side logic of the multi-GPU paths, run on CPU with gloo (world_size 2 and 3).

The exchange code under test is basecount_b200.dist (the same functions the GPU ranks run);
the counting backend is a CPU stand-in over the oracle so no GPU is needed here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basecount_b200 import dist as bdist
from basecount_b200 import synth
from basecount_b200.records import ReadBatch, select_reads


class OracleBackend:
    """Same interface as dist.GpuBackend, counts held in numpy (6 planes x columns)."""

    def __init__(self):
        self.planes = None
        self.length = 0

    def begin(self, lens):
        self.length = int(lens[0])
        self.planes = np.zeros((6, self.length), dtype=np.uint32)

    def count(self, batch, mbq):
        from oracle import bcount as obc
        self.planes += obc.bcount_flat(self.length, mbq, batch).T

    def halo_export(self, ref, col_lo, n):
        return torch.from_numpy(self.planes[:, col_lo:col_lo + n].astype(np.int32).reshape(-1).copy())

    def halo_buffer(self, n):
        return torch.empty(6 * n, dtype=torch.int32)

    def halo_add(self, ref, col_lo, n, t):
        self.planes[:, col_lo:col_lo + n] += t.numpy().reshape(6, n).astype(np.uint32)

    def truncate(self, ref, n):
        self.planes = self.planes[:, :n]
        self.length = n

    def summary(self, show_n=False):
        from oracle import stats as ost
        cov, ent, _ = ost.per_position_vectors(self.planes.T.astype(np.int64).tolist(), show_n)
        return (np.array([sum(1 for x in cov if x)]), np.array([sum(cov)]), np.array([float(np.sum(ent))]))

    def scalar_tensor(self, values, dtype):
        return torch.tensor(values, dtype=dtype)


def _reads(seed, ref_len):
    rec = synth.uniform_short_read_sample(seed=seed, ref_len=ref_len, n_reads=1500, read_len=150, ref_name="x")
    b = select_reads(rec, 0, 0)
    # a few reads with long reference skips so a halo crosses more than one region
    extra = ReadBatch.from_lists(["ACGT" * 5] * 3, [[30] * 20] * 3, [10, ref_len // 3 - 5, ref_len // 2],
                                 [[(0, 10), (3, ref_len // 2), (0, 10)], [(0, 10), (3, ref_len // 3 + 40), (0, 10)],
                                  [(0, 8), (2, 3), (0, 12)]])
    return synth.concat_batches([b, extra])


def _worker(rank, world, port, ref_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = _reads(7, ref_len)
        be = OracleBackend()
        bounds = bdist.count_region_sharded(be, dist, rank, world, batch, ref_len, min_base_quality=20)
        pc, depth, ent = bdist.summary_region_sharded(be, dist, world, ref_len)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), planes=be.planes, lo=bounds[rank], hi=bounds[rank + 1],
                 summary=np.array([pc, depth, ent]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_region_sharding_matches_single_pass(tmp_path, world):
    from oracle import bcount as obc
    from oracle import stats as ost
    ref_len = 6001
    mp.spawn(_worker, args=(world, _free_port(), ref_len, str(tmp_path)), nprocs=world, join=True)
    want = obc.bcount_flat(ref_len, 20, _reads(7, ref_len))
    got = np.zeros_like(want)
    summaries = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert z["planes"].shape[1] == z["hi"] - z["lo"]
        got[int(z["lo"]):int(z["hi"])] = z["planes"].T
        summaries.append(z["summary"])
    assert np.array_equal(got, want)                       # halo merge leaves exactly the single-pass counts
    cov, ent, _ = ost.per_position_vectors(want.astype(np.int64).tolist())
    pc, depth, avg_ent = ost.summary(cov, ent, ref_len)
    for s in summaries:                                    # every rank holds the all-reduced summary
        assert s[0] == pc and s[1] == depth
        assert s[2] == pytest.approx(float(avg_ent), rel=1e-12)


def _bam_worker(rank, world, port, bam, ref_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = OracleBackend()
        bounds, n = bdist.count_region_sharded_bam(be, dist, rank, world, bam, 0, ref_len, min_base_quality=0,
                                                   min_mapping_quality=30, threads=2)
        total = torch.tensor([n], dtype=torch.int64)
        dist.all_reduce(total)
        np.savez(os.path.join(out_dir, f"b{rank}.npz"), planes=be.planes, lo=bounds[rank], hi=bounds[rank + 1],
                 total=int(total.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_region_sharding_from_an_indexed_bam(tmp_path, world):
    """Every rank fetches only its own region through the BAI; the merged counts are
    the single-pass counts of the whole file under the same MAPQ filter (main.py:165)."""
    from basecount_b200 import bamio
    from oracle import bcount as obc
    ref_len = 50_000                                       # spans four 16 kbp index windows
    rec = synth.uniform_short_read_sample(seed=9, ref_len=ref_len, n_reads=4000, read_len=150, ref_name="x")
    bam = str(tmp_path / "x.bam")
    bamio.write_bam(bam, rec)
    bamio.write_bai(bam)
    mp.spawn(_bam_worker, args=(world, _free_port(), bam, ref_len, str(tmp_path)), nprocs=world, join=True)
    whole = select_reads(rec, 0, 30)
    want = obc.bcount_flat(ref_len, 0, whole)
    got = np.zeros_like(want)
    for r in range(world):
        z = np.load(tmp_path / f"b{r}.npz")
        got[int(z["lo"]):int(z["hi"])] = z["planes"].T
        assert int(z["total"]) == whole.n                  # num_reads: every kept read counted on exactly one rank
    assert np.array_equal(got, want)


def test_sample_sharding_covers_every_sample_once():
    for n, world in ((96, 8), (12, 5), (3, 8)):
        seen = sorted(i for r in range(world) for i in bdist.shard_samples(n, world, r))
        assert seen == list(range(n))
    assert [len(bdist.shard_samples(96, 8, r)) for r in range(8)] == [12] * 8


def test_region_helpers():
    b = _reads(3, 4000)
    ends = bdist.ref_ends(b)
    op, ln = b.cigar & 0xF, (b.cigar >> 4).astype(np.int64)
    i = b.n - 3                                            # the first long-skip read
    c0, c1 = int(b.cigar_off[i]), int(b.cigar_off[i + 1])
    assert ends[i] == int(b.starts[i]) + sum(int(l) for o, l in zip(op[c0:c1], ln[c0:c1]) if o in (0, 2, 3, 7, 8))
    bounds = bdist.region_bounds(4000, 8)
    assert bounds[0] == 0 and bounds[-1] == 4000 and np.array_equal(bounds, np.linspace(0, 4000, 9, dtype=int))
    parts = [bdist.select_region(b, int(bounds[r]), int(bounds[r + 1])) for r in range(8)]
    assert sum(p.n for p in parts) == b.n
    assert bdist.halo_columns(parts[0], int(bounds[1]), 4000 - int(bounds[1])) > 0


def _failing_worker(rank, world, port, ref_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = _reads(7, ref_len)
        # one read that runs past the end of the reference: only the LAST rank's count raises (count.cpp .at())
        bad = ReadBatch.from_lists(["ACGT" * 5], [[30] * 20], [ref_len - 5], [[(0, 20)]])
        batch = synth.concat_batches([batch, bad])

        class Raising(OracleBackend):
            def count(self, b, mbq):
                if int((b.starts.astype(np.int64) + 20 > self.length).any()) and rank == world - 1:
                    raise IndexError("alignment counted past the end of the reference")
                super().count(b, mbq)

        try:
            bdist.count_region_sharded(Raising(), dist, rank, world, batch, ref_len)
            outcome = "returned"
        except IndexError:
            outcome = "IndexError"
        open(os.path.join(out_dir, f"r{rank}.txt"), "w").write(outcome)
    finally:
        dist.destroy_process_group()


def test_index_error_on_one_rank_raises_on_every_rank(tmp_path):
    """An alignment past the reference end is an IndexError in the reference (count.cpp:60-64,85); in a
    region-sharded run only the rank that holds the read sees it, and the others must fail too instead of
    waiting in the halo exchange."""
    world = 3
    mp.spawn(_failing_worker, args=(world, _free_port(), 6001, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"r{r}.txt").read() for r in range(world)] == ["IndexError"] * world
