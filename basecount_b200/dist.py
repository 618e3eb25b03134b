"""Multi-GPU partitioning of the counting path: one process per GPU.

Two ways the path shards:

  * sample sharding (config 4): whole samples are independent -> `shard_samples`; every rank
    runs K1 -> K2 -> K3 on its own samples, no data-path collective.
  * region sharding (config 5): the reference is cut at np.linspace boundaries (the split the
    reference's own tests use, tests/test_basecount.py:146-150); a read belongs to the rank that
    owns its start.  Reads that run past the rank's right boundary leave counts in a halo of
    H columns; those columns are sent to the ranks that own them (send/recv over NCCL on
    device buffers) and added there before the statistics, and the three --summarise scalars
    are all-reduced.  The count matrix itself is never reduced across GPUs.

On GPUs the exchange lives in the C-ABI library (bc_comm_init / bc_halo_merge / bc_summary_allreduce_async: NCCL
send/recv and all-reduce enqueued on the engine's compute stream, no host synchronisation between a batch and its
summary).  The host only has to carry the 128-byte communicator id to every rank once -- `engine_comm` does it
over whatever process group is around -- so torch is plumbing here, not part of the data path.  The same logic
with the host moving the halo (exchange_halos over torch.distributed send/recv) is kept for a *backend* without
the library's communicator; tests plug in a CPU stand-in over gloo.
"""
from __future__ import annotations

import numpy as np

from .records import OP_D, OP_EQ, OP_M, OP_N, OP_X, ReadBatch


def shard_samples(n_samples: int, world: int, rank: int):
    """Round-robin sample indices of this rank."""
    return list(range(rank, n_samples, world))


def region_bounds(ref_len: int, world: int) -> np.ndarray:
    return np.linspace(0, ref_len, num=world + 1, dtype=np.int64)


def ref_ends(batch: ReadBatch) -> np.ndarray:
    """reference_end (exclusive) of every read: start + sum of M,=,X,D,N lengths."""
    op = batch.cigar & 0xF
    ln = (batch.cigar >> 4).astype(np.int64)
    consumes = (op == OP_M) | (op == OP_EQ) | (op == OP_X) | (op == OP_D) | (op == OP_N)
    per = np.where(consumes, ln, 0)
    csum = np.concatenate([[0], np.cumsum(per)])
    co = batch.cigar_off.astype(np.int64)
    return batch.starts.astype(np.int64) + (csum[co[1:]] - csum[co[:-1]])


def select_region(batch: ReadBatch, lo: int, hi: int) -> ReadBatch:
    """Reads whose start lies in [lo, hi), with starts made relative to lo."""
    from .synth import take_batch
    idx = np.flatnonzero((batch.starts >= lo) & (batch.starts < hi))
    b = take_batch(batch, idx)
    return ReadBatch((b.starts.astype(np.int64) - lo).astype(np.uint32), b.cigar, b.cigar_off, b.seq, b.qual, b.seq_off)


def halo_columns(local: ReadBatch, region_len: int, cols_right: int) -> int:
    """How many columns past the region's right edge this rank's reads touch (clipped to the
    columns that exist to the right; an alignment past the reference end stays an IndexError)."""
    if local.n == 0:
        return 0
    over = int(ref_ends(local).max()) - region_len
    return int(min(max(over, 0), cols_right))


_FAILED = 0xFFFFFFFF          # a rank's halo width when its count raised


def engine_comm(engine, dist, rank: int, world: int):
    """Join `engine` to a communicator of its own over the ranks of the process group `dist` (any backend: the
    group only broadcasts the 128-byte id).  Collective."""
    box = [engine.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    engine.comm_init(world, rank, box[0])


class GpuBackend:
    """Engine (+ torch CUDA buffers when the host moves the halo itself).  native_comm: the engine holds the
    library's own communicator (engine_comm) and the exchange is bc_halo_merge."""

    def __init__(self, engine, device, native_comm=False):
        import torch
        self.torch = torch
        self.engine = engine
        self.device = device
        self.native_comm = native_comm

    def begin(self, lens):
        self.engine.begin(lens)

    def count(self, batch: ReadBatch, min_base_quality: int):
        from .pack import pack_batches
        self.engine.push(pack_batches(batch, min_base_quality))
        self.engine.sync()

    def halo_export(self, ref, col_lo, n_cols):
        t = self.torch.empty(6 * n_cols, dtype=self.torch.int32, device=self.device)
        self.engine.halo_export(ref, col_lo, n_cols, t.data_ptr())
        return t

    def halo_buffer(self, n_cols):
        return self.torch.empty(6 * n_cols, dtype=self.torch.int32, device=self.device)

    def halo_add(self, ref, col_lo, n_cols, t):
        self.torch.cuda.synchronize(self.device)
        self.engine.halo_add(ref, col_lo, n_cols, t.data_ptr())

    def truncate(self, ref, n):
        self.engine.truncate(ref, n)

    def summary(self, show_n_bases=False):
        return self.engine.summary(show_n_bases)

    def scalar_tensor(self, values, dtype):
        return self.torch.tensor(values, dtype=dtype, device=self.device)


def exchange_halos(backend, dist, rank, world, bounds, halos, ref=0):
    """Send every rank's halo columns to the ranks that own them; add what we receive.

    bounds: region boundaries (world+1); halos[r]: halo columns of rank r (all-gathered)."""
    lo_r, hi_r = int(bounds[rank]), int(bounds[rank + 1])
    ops, recv_bufs, keep = [], [], []
    # what this rank sends: its halo covers global columns [hi_r, hi_r + halos[rank])
    for s in range(rank + 1, world):
        a = max(int(bounds[s]), hi_r)
        b = min(int(bounds[s + 1]), hi_r + int(halos[rank]))
        if a < b:
            t = backend.halo_export(ref, a - lo_r, b - a)
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, s))
    # what this rank receives: halos of ranks to the left that reach into [lo_r, hi_r)
    for q in range(rank):
        hq = int(bounds[q + 1])
        a = max(lo_r, hq)
        b = min(hi_r, hq + int(halos[q]))
        if a < b:
            t = backend.halo_buffer(b - a)
            recv_bufs.append((a - lo_r, b - a, t))
            ops.append(dist.P2POp(dist.irecv, t, q))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for col, n, t in recv_bufs:
        backend.halo_add(ref, col, n, t)


def count_region_sharded(backend, dist, rank, world, batch: ReadBatch, ref_len: int, min_base_quality: int = 0,
                         local_reads: ReadBatch | None = None):
    """Count this rank's region of one reference and merge the boundary halos.

    `batch` holds all reads (each rank selects its own by start), or pass `local_reads`
    (already selected and made relative) to skip the selection.  Afterwards the backend's
    slot 0 holds exactly the owned columns [bounds[rank], bounds[rank+1]).  Returns bounds."""
    import torch
    bounds = region_bounds(ref_len, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    local = local_reads if local_reads is not None else select_region(batch, lo, hi)
    h = halo_columns(local, hi - lo, ref_len - hi)
    backend.begin([hi - lo + h])
    # An alignment past the reference end raises IndexError on the rank that holds it (count.cpp .at()); the
    # other ranks must not be left waiting in the collective that follows: the failure travels with the halo
    # widths and every rank raises.
    failed = None
    try:
        backend.count(local, min_base_quality)
    except IndexError as e:
        failed = e
    mine_h = _FAILED if failed is not None else h
    if getattr(backend, "native_comm", False):
        halos = backend.engine.allgather_u32(mine_h)
        if _FAILED in halos:
            raise failed if failed is not None else IndexError("another rank counted past the end of the reference")
        backend.engine.halo_merge(0, bounds, halos)          # asynchronous; the slot now holds the owned columns
        return bounds
    if world > 1:
        mine = backend.scalar_tensor([mine_h], torch.int64)
        gathered = [backend.scalar_tensor([0], torch.int64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        halos = [int(t.item()) for t in gathered]
        if _FAILED in halos:
            raise failed if failed is not None else IndexError("another rank counted past the end of the reference")
        exchange_halos(backend, dist, rank, world, bounds, halos)
    elif failed is not None:
        raise failed
    backend.truncate(0, hi - lo)
    return bounds


def load_region_reads(bam_path: str, ref_id: int, lo: int, hi: int, min_mapping_quality: int = 0,
                      index: str | None = None, threads: int = 0) -> ReadBatch:
    """Kept reads of reference `ref_id` whose start lies in [lo, hi), starts relative to lo, decoded through
    the BAI index (csrc/bam_index.h): only this region's BGZF blocks are read and inflated, so with one
    process per GPU the host decode scales with the number of ranks instead of every rank reading the file."""
    from .bamio import NativeBam
    nb = NativeBam(bam_path, threads=threads, region=(ref_id, lo, hi), index=index)
    try:
        b = nb.select(ref_id, min_mapping_quality)
    finally:
        nb.close()
    return ReadBatch((b.starts.astype(np.int64) - lo).astype(np.uint32), b.cigar, b.cigar_off, b.seq, b.qual, b.seq_off)


def count_region_sharded_bam(backend, dist, rank, world, bam_path: str, ref_id: int, ref_len: int,
                             min_base_quality: int = 0, min_mapping_quality: int = 0, index: str | None = None,
                             threads: int = 0):
    """count_region_sharded with every rank fetching its own region of an indexed, coordinate-sorted BAM.
    Returns (bounds, number of reads this rank counted)."""
    bounds = region_bounds(ref_len, world)
    local = load_region_reads(bam_path, ref_id, int(bounds[rank]), int(bounds[rank + 1]), min_mapping_quality, index,
                              threads)
    count_region_sharded(backend, dist, rank, world, None, ref_len, min_base_quality, local_reads=local)
    return bounds, local.n


def summary_region_sharded(backend, dist, world, ref_len: int, show_n_bases: bool = False):
    """(pc_reference_coverage, avg_depth, avg_entropy) of the whole reference (main.py:479-485)
    from per-rank K3 partials: all-reduce of {nonzero, coverage sum} (int64) and entropy sum (f64)."""
    import torch
    if getattr(backend, "native_comm", False):
        from . import _lib
        out = (_lib.pinned_empty(1, np.int64), _lib.pinned_empty(1, np.int64), _lib.pinned_empty(1, np.float64))
        backend.engine.summary_allreduce_async(out, show_n_bases)
        backend.engine.sync()
        return 100 * (int(out[0][0]) / ref_len), np.float64(int(out[1][0])) / ref_len, np.float64(out[2][0]) / ref_len
    nz, cs, es = backend.summary(show_n_bases)
    # one collective: the two integers ride as float64 (sums below 2^53 are exact)
    assert int(cs[0]) < 2 ** 53 // max(world, 1)
    t = backend.scalar_tensor([float(int(nz[0])), float(int(cs[0])), float(es[0])], torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    nonzero, cov_sum, ent_sum = t.tolist()
    return 100 * (int(nonzero) / ref_len), np.float64(int(cov_sum)) / ref_len, np.float64(ent_sum) / ref_len
