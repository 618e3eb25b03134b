"""Python face of the C ABI: one Engine = one GPU's accumulators, streams and kernels."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib
from .pack import PackedBatch


def norm_factors(show_n_bases: bool):
    """(1/log2(K), 1/log2(K-1)) computed exactly as the reference does (main.py:22-25)."""
    k = 6 if show_n_bases else 5
    return 1 / math.log2(k), 1 / math.log2(k - 1)


class ResidentBatch:
    """A packed batch kept in HBM (bench kernel-only timing; repeated counting)."""

    def __init__(self, engine: "Engine", host: PackedBatch):
        self.engine = engine
        self.host = host
        self.struct = _lib.BcBatch()
        _lib.check(engine._h, _lib.lib().bc_batch_upload(engine._h, ctypes.byref(host.as_struct()),
                                                         ctypes.byref(self.struct)))

    def free(self):
        if self.struct is not None and self.engine._h:
            _lib.lib().bc_batch_free(self.engine._h, ctypes.byref(self.struct))
        self.struct = None


class Engine:
    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        self._h = ctypes.c_void_p()
        rc = self._L.bc_create(int(device), ctypes.byref(self._h))
        if rc != _lib.BC_OK:
            msg = self._L.bc_last_error(None)
            self._h = None
            raise RuntimeError(f"bc_create(device={device}) failed: {msg.decode() if msg else rc}")
        self.device = device
        self.ref_lens = []
        self._keepalive = []
        self.comm_world, self.comm_rank = 1, 0

    # -- lifecycle
    def close(self):
        if self._h:
            self._L.bc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- accumulators
    def begin(self, ref_lens):
        lens = np.ascontiguousarray(ref_lens, dtype=np.int64)
        if lens.ndim != 1 or lens.size == 0 or (lens < 0).any():
            raise TypeError("ref_lens must be a non-empty list of unsigned ints")
        arr = lens.astype(np.uint32)
        _lib.check(self._h, self._L.bc_begin(self._h, arr.size, _lib.ptr(arr)))
        self.ref_lens = [int(x) for x in lens]
        self._keepalive.clear()

    def reset(self):
        """Zero the accumulators again, keeping the slots (asynchronous)."""
        _lib.check(self._h, self._L.bc_reset(self._h))

    # -- the operator
    def push(self, batch, keep: int = 0):
        """Queue one batch (PackedBatch in host memory, or ResidentBatch in HBM). Asynchronous.

        Host buffers must outlive the asynchronous copy.  By default every pushed batch is held until sync();
        keep=2 holds only the last two: the library has two staging sets and bc_push_batch waits for the set it
        is about to overwrite, so when push k returns, batch k - 2 has been copied and counted -- a stream of
        chunks then needs host memory for two chunks, not for the file."""
        if isinstance(batch, ResidentBatch):
            _lib.check(self._h, self._L.bc_push_batch(self._h, ctypes.byref(batch.struct)))
            return
        self._keepalive.append(batch)           # host buffers must outlive the async copies
        _lib.check(self._h, self._L.bc_push_batch(self._h, ctypes.byref(batch.as_struct())))
        if keep > 0 and len(self._keepalive) > keep:
            del self._keepalive[:-keep]

    def sync(self):
        try:
            _lib.check(self._h, self._L.bc_sync(self._h))
        finally:
            self._keepalive.clear()

    def upload(self, packed: PackedBatch) -> ResidentBatch:
        return ResidentBatch(self, packed)

    # -- results
    def counts(self, ref: int = 0) -> np.ndarray:
        out = np.empty((self.ref_lens[ref], 6), dtype=np.int64)
        _lib.check(self._h, self._L.bc_counts(self._h, ref, _lib.ptr(out)))
        return out

    def stats(self, ref: int = 0, show_n_bases: bool = False, want_pc: bool = True):
        L = self.ref_lens[ref]
        k = 6 if show_n_bases else 5
        n1, n2 = norm_factors(show_n_bases)
        cov = np.empty(L, dtype=np.int64)
        pc = np.empty((k, L), dtype=np.float64) if want_pc else None
        ent = np.empty(L, dtype=np.float64)
        sec = np.empty(L, dtype=np.float64)
        flags = np.empty(L, dtype=np.uint8)
        _lib.check(self._h, self._L.bc_stats(self._h, ref, int(show_n_bases), n1, n2, _lib.ptr(cov), _lib.ptr(pc),
                                             _lib.ptr(ent), _lib.ptr(sec), _lib.ptr(flags)))
        return {"coverage": cov, "pc": pc, "entropy": ent, "secondary": sec, "flags": flags}

    def rows_window(self, ref: int, lo: int, n: int, show_n_bases: bool = False, bufs=None):
        """Counts and per-position statistics of columns [lo, lo + n) of a slot: (counts n x 6, stats dict as stats()).
        `bufs` (from a previous call with the same n and show_n_bases) is reused instead of allocating."""
        k = 6 if show_n_bases else 5
        n1, n2 = norm_factors(show_n_bases)
        if bufs is None or bufs[0].shape[0] != n:
            bufs = (np.empty((n, 6), np.int64), np.empty(n, np.int64), np.empty((k, n), np.float64), np.empty(n, np.float64),
                    np.empty(n, np.float64), np.empty(n, np.uint8))
        cnt, cov, pc, ent, sec, flags = bufs
        _lib.check(self._h, self._L.bc_rows_window(self._h, ref, int(show_n_bases), n1, n2, int(lo), int(n), _lib.ptr(cnt),
                                                   _lib.ptr(cov), _lib.ptr(pc), _lib.ptr(ent), _lib.ptr(sec), _lib.ptr(flags)))
        return cnt, {"coverage": cov, "pc": pc, "entropy": ent, "secondary": sec, "flags": flags}, bufs

    def summary(self, show_n_bases: bool = False):
        r = len(self.ref_lens)
        n1, n2 = norm_factors(show_n_bases)
        nz = np.empty(r, dtype=np.int64)
        cs = np.empty(r, dtype=np.int64)
        es = np.empty(r, dtype=np.float64)
        _lib.check(self._h, self._L.bc_summary(self._h, int(show_n_bases), n1, n2, _lib.ptr(nz), _lib.ptr(cs), _lib.ptr(es)))
        return nz, cs, es

    def summary_min_coverage(self, min_coverage: int = 0, show_n_bases: bool = False):
        """Per slot: positions with coverage >= min_coverage, coverage sum over all positions, entropy sum
        over the selected positions (device reductions behind mean_coverage / mean_entropy)."""
        r = len(self.ref_lens)
        n1, _ = norm_factors(show_n_bases)
        sel = np.empty(r, dtype=np.int64)
        cs = np.empty(r, dtype=np.int64)
        es = np.empty(r, dtype=np.float64)
        _lib.check(self._h, self._L.bc_summary_min_coverage(self._h, int(show_n_bases), n1, int(min_coverage),
                                                            _lib.ptr(sel), _lib.ptr(cs), _lib.ptr(es)))
        return sel, cs, es

    def summary_async(self, out, show_n_bases: bool = False):
        """Queue the summarise reductions; `out` = (nonzero, cov_sum, ent_sum) pinned arrays
        (see _lib.pinned_empty), valid after sync()."""
        n1, n2 = norm_factors(show_n_bases)
        _lib.check(self._h, self._L.bc_summary_async(self._h, int(show_n_bases), n1, n2, _lib.ptr(out[0]),
                                                     _lib.ptr(out[1]), _lib.ptr(out[2])))

    def amplicons(self, ref: int, lo, hi, show_n_bases: bool = False):
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        t = int(lo.shape[0])
        out = np.empty((6, t), dtype=np.float64)
        empty = np.zeros(t, dtype=np.uint8)
        n1, n2 = norm_factors(show_n_bases)
        _lib.check(self._h, self._L.bc_amplicons(self._h, ref, int(show_n_bases), n1, n2, t, _lib.ptr(lo), _lib.ptr(hi),
                                                 _lib.ptr(out), _lib.ptr(empty)))
        return out, empty

    def amplicons_async(self, ref: int, lo, hi, out, empty, show_n_bases: bool = False):
        """Queue the amplicon reductions; `out` (6 x T float64) and `empty` (T uint8) are caller-owned
        arrays that hold the results after sync()."""
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        t = int(lo.shape[0])
        assert out.dtype == np.float64 and out.size == 6 * t and empty.dtype == np.uint8 and empty.size == t
        n1, n2 = norm_factors(show_n_bases)
        _lib.check(self._h, self._L.bc_amplicons_async(self._h, ref, int(show_n_bases), n1, n2, t, _lib.ptr(lo),
                                                       _lib.ptr(hi), _lib.ptr(out), _lib.ptr(empty)))

    # -- region sharding
    def halo_export(self, ref: int, col_lo: int, n_cols: int, dev_ptr: int):
        _lib.check(self._h, self._L.bc_halo_export(self._h, ref, col_lo, n_cols, ctypes.c_void_p(dev_ptr)))

    def halo_add(self, ref: int, col_lo: int, n_cols: int, dev_ptr: int):
        _lib.check(self._h, self._L.bc_halo_add(self._h, ref, col_lo, n_cols, ctypes.c_void_p(dev_ptr)))

    def truncate(self, ref: int, new_len: int):
        _lib.check(self._h, self._L.bc_truncate(self._h, ref, new_len))
        self.ref_lens[ref] = int(new_len)

    def set_length(self, ref: int, new_len: int):
        """Asynchronous bc_truncate: a device-side write on the compute stream."""
        _lib.check(self._h, self._L.bc_set_length(self._h, ref, new_len))
        self.ref_lens[ref] = int(new_len)

    # -- the library's own NCCL communicator (region sharding without torch on the data path)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = _lib.lib().bc_comm_unique_id(buf)
        if rc != _lib.BC_OK:
            msg = _lib.lib().bc_last_error(None)
            raise RuntimeError(f"bc_comm_unique_id failed: {msg.decode() if msg else rc}")
        return buf.raw

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        _lib.check(self._h, self._L.bc_comm_init(self._h, int(world), int(rank), ctypes.c_char_p(unique_id)))
        self.comm_world, self.comm_rank = int(world), int(rank)

    def comm_destroy(self):
        if self._h:
            self._L.bc_comm_destroy(self._h)

    def allgather_u32(self, value: int):
        out = np.zeros(self.comm_world, dtype=np.uint32)
        _lib.check(self._h, self._L.bc_comm_allgather_u32(self._h, int(value), _lib.ptr(out)))
        return [int(x) for x in out]

    def halo_merge(self, ref: int, bounds, halos):
        """Collective and asynchronous: send this rank's halo columns to their owners, add what arrives, cut the
        slot to the owned columns (csrc/bc_api.cu: bc_halo_merge)."""
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        hl = np.ascontiguousarray(halos, dtype=np.uint32)
        assert b.shape[0] == self.comm_world + 1 and hl.shape[0] == self.comm_world
        _lib.check(self._h, self._L.bc_halo_merge(self._h, ref, _lib.ptr(b), _lib.ptr(hl)))
        self.ref_lens[ref] = int(b[self.comm_rank + 1] - b[self.comm_rank])

    def summary_allreduce_async(self, out, show_n_bases: bool = False):
        """summary_async + all-reduce over the communicator; `out` as in summary_async, valid after sync()."""
        n1, n2 = norm_factors(show_n_bases)
        _lib.check(self._h, self._L.bc_summary_allreduce_async(self._h, int(show_n_bases), n1, n2, _lib.ptr(out[0]),
                                                               _lib.ptr(out[1]), _lib.ptr(out[2])))

    # -- instrumentation
    def timer_start(self):
        _lib.check(self._h, self._L.bc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        _lib.check(self._h, self._L.bc_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def last_count_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        _lib.check(self._h, self._L.bc_last_count_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def count_kernel_ms_history(self, n: int):
        """Device times (ms) of the last n counting-kernel launches, most recent first."""
        buf = (ctypes.c_float * n)()
        got = self._L.bc_count_kernel_ms_history(self._h, buf, n)
        if got < 0:
            raise RuntimeError("bc_count_kernel_ms_history failed")
        return [float(buf[i]) for i in range(got)]

    def h2d_probe(self, nbytes: int, reps: int = 8) -> float:
        """GB/s of plain pinned host-to-device copies of nbytes on this engine's copy stream (diagnostic)."""
        g = ctypes.c_double()
        _lib.check(self._h, self._L.bc_h2d_probe(self._h, int(nbytes), int(reps), ctypes.byref(g)))
        return float(g.value)

    def kernel_launches(self) -> int:
        return int(self._L.bc_kernel_launches(self._h))

    def set_count_variant(self, variant: int):
        _lib.check(self._h, self._L.bc_set_count_variant(self._h, variant))
