"""ctypes binding of libbasecount_b200.so (include/basecount_b200.h).

The shared library is built in-tree by `python -m basecount_b200.build` (nvcc, sm_100a).
There is no fallback: if the library is missing, or no CUDA device is present when an
engine is created, this module raises -- it never computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BASECOUNT_B200_LIB: another build of the same library (kernel A/B experiments); still no fallback
LIB_PATH = os.environ.get("BASECOUNT_B200_LIB") or os.path.join(_HERE, "csrc", "libbasecount_b200.so")

BC_OK, BC_ERR_ARG, BC_ERR_INDEX, BC_ERR_CUDA, BC_ERR_STATE, BC_ERR_READ_OVERRUN, BC_ERR_MISSING_QUAL = range(7)

_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


class BcBatch(ctypes.Structure):
    """struct bc_batch"""
    _fields_ = [
        ("n_reads", ctypes.c_uint32), ("n_refs", ctypes.c_uint32),
        ("ref_read_off", ctypes.c_void_p), ("starts", ctypes.c_void_p), ("cigar_off", ctypes.c_void_p),
        ("cigar", ctypes.c_void_p), ("seq_woff", ctypes.c_void_p), ("planes", ctypes.c_void_p),
        ("okmask", ctypes.c_void_p),
        ("n_exc", ctypes.c_uint32),
        ("exc_read", ctypes.c_void_p), ("exc_pos", ctypes.c_void_p),
        ("on_device", ctypes.c_uint32), ("sorted_hint", ctypes.c_uint32), ("mean_read_len", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


# name -> (restype, argtypes); every symbol the header declares
_SIGNATURES = {
    "bc_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_destroy": (None, [ctypes.c_void_p]),
    "bc_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "bc_device_count": (ctypes.c_int, []),
    "bc_device_pci_bus_id": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "bc_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_host_free": (None, [ctypes.c_void_p]),
    "bc_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "bc_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "bc_push_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BcBatch)]),
    "bc_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "bc_batch_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BcBatch), ctypes.POINTER(BcBatch)]),
    "bc_batch_free": (None, [ctypes.c_void_p, ctypes.POINTER(BcBatch)]),
    "bc_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "bc_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_summary": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_summary_min_coverage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_summary_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_amplicons_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_amplicons": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_halo_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    "bc_halo_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    "bc_truncate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]),
    "bc_pack_words": (ctypes.c_uint64, [ctypes.c_uint32, ctypes.c_void_p]),
    "bc_pack_reads": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]),
    "bc_timer_start": (ctypes.c_int, [ctypes.c_void_p]),
    "bc_timer_stop": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "bc_last_count_kernel_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "bc_count_kernel_ms_history": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
    "bc_kernel_launches": (ctypes.c_uint64, [ctypes.c_void_p]),
    "bc_h2d_probe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "bc_set_count_variant": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bc_rows_window": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "bc_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "bc_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bc_comm_allgather_u32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "bc_halo_merge": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_summary_allreduce_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_set_length": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]),
    "bc_canonical_cigars": (ctypes.c_uint64, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    # exact native TSV rows (host code)
    "bc_format_tsv": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]),
    "bc_free_text": (None, [ctypes.c_void_p]),
    # native BAM decode (host code in the same library)
    "bc_bam_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_bam_stream_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_bam_stream_next": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_bam_stream_close": (None, [ctypes.c_void_p]),
    "bc_bam_last_error": (ctypes.c_char_p, []),
    "bc_bgzf_crc32": (ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_uint64]),
    "bc_inflate_raw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]),
    "bc_bam_pack_sizes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32,
                                         ctypes.c_void_p]),
    "bc_bam_pack_fill": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_uint64)]),
    "bc_bam_index_build": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]),
    "bc_bam_open_region": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bc_bam_close": (None, [ctypes.c_void_p]),
    "bc_bam_num_records": (ctypes.c_uint64, [ctypes.c_void_p]),
    "bc_bam_num_refs": (ctypes.c_uint32, [ctypes.c_void_p]),
    "bc_bam_ref_name": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_uint32]),
    "bc_bam_ref_len": (ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_uint32]),
    "bc_bam_core": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bc_bam_select_sizes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32,
                                           ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                           ctypes.POINTER(ctypes.c_uint64)]),
    "bc_bam_select_fill": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_LIB = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once).  Raises loudly if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m basecount_b200.build` "
                "(nvcc, sm_100a).  basecount_b200 has no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = l
    return _LIB


def ptr(a):
    """Address of a numpy array's data, or None."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


def check(handle, rc):
    """Map a bc_status to the exception the reference raises at the same point."""
    if rc == BC_OK:
        return
    msg = lib().bc_last_error(handle)
    msg = msg.decode() if msg else f"bc_status {rc}"
    if rc == BC_ERR_INDEX:
        raise IndexError(msg)                 # pybind11's translation of std::out_of_range (count.cpp .at())
    if rc == BC_ERR_ARG:
        raise TypeError(msg)                  # pybind11 caster failure
    if rc == BC_ERR_READ_OVERRUN:
        raise ValueError(msg or "CIGAR consumes more bases than the read holds")
    raise RuntimeError(msg)


def pinned_empty(n, dtype):
    """A numpy array backed by pinned (page-locked) host memory from bc_host_alloc."""
    dtype = np.dtype(dtype)
    nbytes = max(int(n) * dtype.itemsize, 1)
    p = ctypes.c_void_p()
    if lib().bc_host_alloc(nbytes, ctypes.byref(p)) != BC_OK or not p.value:
        raise MemoryError(f"bc_host_alloc({nbytes}) failed")
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    weakref.finalize(buf, lib().bc_host_free, ctypes.c_void_p(p.value))
    return arr
