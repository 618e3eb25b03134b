/*
 * basecount_b200.h -- C ABI of the B200-native pileup counting path.
 *
 * This is the drop-in boundary for the ONE native operator of tombch/basecount,
 *     count.bcount(refLen, minBaseQuality, reads, qualities, starts, ctuples)
 *     (reference: basecount/count.cpp:7-14, bound at count.cpp:102-105, called from
 *      basecount/main.py:146-153 and main.py:179-186)
 * plus the numeric work the reference does around it in Python:
 *     int64 accumulation across chunks          main.py:132,155,188
 *     per-position statistics (get_stats)       main.py:10-79
 *     --summarise reductions                    main.py:479-485
 *     --summarise-with-bed amplicon vectors     main.py:519-551
 *
 * Plain C: pointers and sizes only, no torch / pybind types.  Every function
 * returns a bc_status (0 = OK).  The Python side (basecount_b200/_lib.py) loads
 * libbasecount_b200.so with ctypes and maps the codes to the exceptions the
 * reference raises (BC_ERR_INDEX -> IndexError as pybind11 does for the
 * std::out_of_range thrown by .at() at count.cpp:60-64,85; BC_ERR_ARG -> TypeError).
 *
 * There is no CPU implementation behind this ABI: without a CUDA device every
 * compute entry point fails with BC_ERR_CUDA.
 */
#ifndef BASECOUNT_B200_H
#define BASECOUNT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bc_status {
    BC_OK = 0,
    BC_ERR_ARG = 1,        /* bad argument / malformed batch           -> TypeError  */
    BC_ERR_INDEX = 2,      /* a counted event fell at refPos >= refLen -> IndexError */
    BC_ERR_CUDA = 3,       /* CUDA runtime failure (message in bc_last_error)        */
    BC_ERR_STATE = 4,      /* call out of order (no bc_begin, ...)                   */
    BC_ERR_READ_OVERRUN = 5, /* CIGAR consumes more bases than the read holds (UB in the reference) */
    BC_ERR_MISSING_QUAL = 6 /* a kept read has QUAL '*': pysam hands the reference None     -> TypeError  */
} bc_status;

typedef struct bc_handle bc_handle;

/*
 * One batch of reads in structure-of-arrays form.  This replaces the four Python
 * lists of count.cpp:10-13 (reads, qualities, starts, ctuples):
 *
 *   starts[i]      reference_start of read i, relative to ITS reference  (count.cpp:35)
 *   cigar[...]     BAM-native words  len << 4 | op ; read i owns
 *                  cigar[cigar_off[i] .. cigar_off[i+1])                  (count.cpp:40-46).
 *                  Any CIGAR is counted correctly.  The packers of this library hand over a normal form
 *                  with the same meaning (bc_canonical_cigars: =/X spelled M, N spelled D, S/H/P and
 *                  zero-length operations dropped, equal neighbours merged -- count.cpp:51,74,80,92 treats
 *                  them alike); reads whose CIGAR is "M" or "M, I or D, M" in that form take the fast
 *                  counting kernel, everything else the general one.
 *   planes[...]    the query_alignment_sequence, 2 bits per base, bit-planar:
 *                  base j of read i is bit (j & 31) of word
 *                  planes[seq_woff[i] + (j >> 5)]; the low 32 bits of the 64-bit
 *                  word hold bit 0 of the code, the high 32 bits hold bit 1
 *                  (A=0 C=1 G=2 T=3; anything else is stored as 0 and listed in exc_*).
 *                  Every read starts on a fresh 64-bit word; unused bits are 0.
 *   okmask[...]    optional (NULL when min_base_quality == 0): one 32-bit word per
 *                  planes word, bit set <=> quality >= minBaseQuality AND letter in ACGT
 *                  (the test at count.cpp:56 evaluated by the packer).
 *   exc_read/exc_pos  sparse list of bases the 2-bit code cannot express, sorted by
 *                  (read, pos):  exc_pos = pos << 2 | flags,
 *                  flag 1 = count column N (letter 'N' that passes the quality test, count.cpp:64),
 *                  flag 2 = the main pass counted this base as 'A' and must be undone
 *                           (only ever set when okmask == NULL).
 *   ref_read_off   reads are grouped by reference slot: slot r owns reads
 *                  [ref_read_off[r], ref_read_off[r+1]); n_refs+1 entries, matching bc_begin.
 *
 * Reads sorted by start within a slot are the fast case (coordinate-sorted BAM);
 * any order is counted correctly.
 */
typedef struct bc_batch {
    uint32_t n_reads;
    uint32_t n_refs;
    const uint32_t *ref_read_off;   /* n_refs + 1 */
    const uint32_t *starts;         /* n_reads */
    const uint32_t *cigar_off;      /* n_reads + 1 */
    const uint32_t *cigar;          /* cigar_off[n_reads] */
    const uint32_t *seq_woff;       /* n_reads + 1, in 64-bit words */
    const uint64_t *planes;         /* seq_woff[n_reads] */
    const uint32_t *okmask;         /* seq_woff[n_reads] or NULL */
    uint32_t n_exc;
    const uint32_t *exc_read;       /* n_exc */
    const uint32_t *exc_pos;        /* n_exc */
    uint32_t on_device;             /* 0: host pointers (copied by bc_push_batch); 1: device pointers */
    uint32_t sorted_hint;           /* 1 if starts are non-decreasing within every slot */
    uint32_t mean_read_len;         /* packer's estimate, picks the lane-group width; 0 = unknown */
    uint32_t reserved;
} bc_batch;

/* ---- lifecycle ------------------------------------------------------------ */
int bc_create(int device, bc_handle **out);
void bc_destroy(bc_handle *h);
const char *bc_last_error(bc_handle *h);     /* h may be NULL: last create error */
int bc_device_count(void);
/* PCI bus id of a device ("0000:1b:00.0", lower case as under /sys/bus/pci/devices) so the host can
 * place its threads and pinned buffers on the device's NUMA node.  len >= 16. */
int bc_device_pci_bus_id(int device, char *out, int len);

/* Pinned host memory for batches and results (cudaHostAlloc). */
int bc_host_alloc(size_t bytes, void **out);
void bc_host_free(void *p);

/* ---- accumulators (replaces np.zeros((L,6)) at main.py:132) ----------------- */
/* Allocate and zero one refLen x 6 accumulator per reference slot. */
int bc_begin(bc_handle *h, uint32_t n_refs, const uint32_t *ref_lens);
/* Zero the accumulators of the current slots again (asynchronous, compute stream). */
int bc_reset(bc_handle *h);

/* ---- the operator (replaces bcount + np.add, main.py:146-157 / 179-189) ----- */
/* Asynchronous: H2D on the copy stream (host batches), counting kernels on the
 * compute stream.  Host buffers must stay valid until bc_sync. */
int bc_push_batch(bc_handle *h, const bc_batch *b);
/* Wait for everything pushed so far; returns BC_ERR_INDEX / BC_ERR_READ_OVERRUN if any
 * batch hit the reference's error conditions (accumulators are then unspecified). */
int bc_sync(bc_handle *h);

/* Keep a batch resident in HBM (bench: kernel-only timing).  dev receives device
 * pointers and on_device = 1; release with bc_batch_free. */
int bc_batch_upload(bc_handle *h, const bc_batch *host, bc_batch *dev);
void bc_batch_free(bc_handle *h, bc_batch *dev);

/* ---- results ---------------------------------------------------------------- */
/* refLen x 6 int64, row-major, columns A,C,G,T,DS,N (count.cpp:16-17; int64 as main.py:132). */
int bc_counts(bc_handle *h, uint32_t ref, int64_t *out);

/* Per-position statistics (get_stats, main.py:14-53).  K = show_n ? 6 : 5.
 * norm = 1/log2(K), norm2 = 1/log2(K-1) (main.py:24-25, computed by the caller so the
 * constants are bit-identical to the reference's).
 * Outputs (any may be NULL): coverage[L]; pc[K*L] (pc[k*L + pos]); entropy[L];
 * secondary[L]; flags[L]: bit0 = coverage 0 (reference emits int -1 / 1 / 1, main.py:34-36),
 * bit1 = secondary coverage 0 (secondary_entropy stays int 1, main.py:47).  The float arrays
 * hold -1.0 / 1.0 / 1.0 at flagged positions. */
int bc_stats(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2,
             int64_t *coverage, double *pc, double *entropy, double *secondary, uint8_t *flags);

/* bc_counts and bc_stats for the window [lo, lo + n) of a slot only (any output may be NULL; counts is n x 6 int64,
 * pc is pc[k * n + i]): what lets `basecount BAM_FILE` print a 64 Mb reference window by window (main.py:456-466)
 * without the host ever holding more than one window of rows. */
int bc_rows_window(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t lo, uint32_t n,
                   int64_t *counts, int64_t *coverage, double *pc, double *entropy, double *secondary, uint8_t *flags);

/* --summarise reductions for ALL slots in one pass (main.py:479-485): per slot the
 * number of positions with coverage != 0, the sum of coverage and the sum of entropy
 * (entropy = 1 at zero coverage). */
int bc_summary(bc_handle *h, int show_n, double norm, double norm2,
               int64_t *nonzero, int64_t *cov_sum, double *entropy_sum);
/* Same, without waiting: the three outputs (any host memory) are filled in by the next bc_sync
 * and must stay alive until then.  Lets a stream of batches pipeline without a host sync each. */
int bc_summary_async(bc_handle *h, int show_n, double norm, double norm2,
                     int64_t *nonzero, int64_t *cov_sum, double *entropy_sum);

/* BaseCount.mean_entropy(min_coverage) / mean_coverage (main.py:325-359) served from the device:
 * per slot the number of positions with coverage >= min_coverage, the sum of their entropies and
 * the sum of coverage over ALL positions.  Synchronous. */
int bc_summary_min_coverage(bc_handle *h, int show_n, double norm, int64_t min_coverage,
                            int64_t *selected, int64_t *cov_sum, double *entropy_sum_selected);

/* --summarise-with-bed amplicon vectors (main.py:519-551) for one slot.
 * Window t covers 0-based positions lo[t]..hi[t] inclusive, clipped to the reference.
 * out[6*n_tiles]: rows mean_cov, median_cov, mean_ent, median_ent, mean_sec, median_sec;
 * empty[t] = 1 where the window holds no position (reference prints int -1). */
int bc_amplicons(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2,
                 uint32_t n_tiles, const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty);

/* Same, without waiting: out / empty (any host memory) are filled in by the next bc_sync and must stay
 * alive until then; lo / hi are consumed before the call returns. */
int bc_amplicons_async(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2,
                       uint32_t n_tiles, const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty);

/* ---- region sharding (config 5): one process per GPU, boundary-column halo merge over NCCL ---------------
 * The reference is cut at `bounds` (the np.linspace split of the reference's tests, tests/test_basecount.py:146-150);
 * rank r owns global columns [bounds[r], bounds[r+1]) and counts the reads that START there into a slot of
 * own + halos[r] columns (halos[r] = how far its reads run past its right boundary, known when they are packed).
 * NCCL is bound at run time (libnccl.so.2, or BASECOUNT_B200_NCCL); nothing here needs PyTorch. */
/* Rank 0 makes a 128-byte id (ncclGetUniqueId) and hands it to the other ranks by whatever the host has. */
int bc_comm_unique_id(void *id128);
/* Collective: every rank's handle joins the communicator (ncclCommInitRank on the handle's device). */
int bc_comm_init(bc_handle *h, int world, int rank, const void *id128);
int bc_comm_destroy(bc_handle *h);
/* Collective, synchronous: out[r] = rank r's value (how the ranks learn each other's halo widths). */
int bc_comm_allgather_u32(bc_handle *h, uint32_t mine, uint32_t *out);
/* Collective, ASYNCHRONOUS (compute stream, no host synchronisation): the halo columns of slot `ref` go straight
 * from the accumulators to the ranks that own them (ncclSend per plane segment; a skip may reach past several
 * regions), what arrives from the ranks to the left is added to this rank's first columns, and the slot is cut
 * to the owned columns (bc_set_length) -- so the statistics that follow cover exactly [bounds[rank], bounds[rank+1]).
 * bounds: world + 1 entries, halos: world entries (bc_comm_allgather_u32). */
int bc_halo_merge(bc_handle *h, uint32_t ref, const uint32_t *bounds, const uint32_t *halos);
/* bc_summary_async followed by an all-reduce (sum) of the per-slot scalars over the communicator, on the
 * compute stream: every rank gets the whole reference's numbers at the next bc_sync. */
int bc_summary_allreduce_async(bc_handle *h, int show_n, double norm, double norm2,
                               int64_t *nonzero, int64_t *cov_sum, double *entropy_sum);
/* Set a slot's length to its first new_len columns (new_len <= its length at bc_begin) ASYNCHRONOUSLY, as a
 * device-side write on the compute stream: drop the halo columns after a merge, restore them before the next
 * batch of a stream of samples.  Counts are not touched. */
int bc_set_length(bc_handle *h, uint32_t ref, uint32_t new_len);

/* Building blocks of the same exchange for a host that moves the halo itself (e.g. torch.distributed): copy / add
 * the u32 counts of columns [col_lo, col_lo + n_cols) of a slot, all six planes (6 * n_cols values, plane-major);
 * buf is a DEVICE pointer.  These three wait for the stream. */
int bc_halo_export(bc_handle *h, uint32_t ref, uint32_t col_lo, uint32_t n_cols, uint32_t *dev_buf);
int bc_halo_add(bc_handle *h, uint32_t ref, uint32_t col_lo, uint32_t n_cols, const uint32_t *dev_buf);
int bc_truncate(bc_handle *h, uint32_t ref, uint32_t new_len);

/* ---- host-side packer (replaces pybind11's list -> std::vector casters) -------- */
/* Sizes needed for bc_pack_reads outputs, from the per-read base counts. */
uint64_t bc_pack_words(uint32_t n_reads, const uint64_t *seq_off);
/* ASCII bases + phred bytes -> planes / okmask / exception list.
 * okmask_out may be NULL iff min_base_quality == 0.  exc arrays must hold exc_cap
 * entries; *n_exc receives the number needed (call again with more room if > exc_cap).
 * Also validates that no CIGAR consumes more bases than its read holds. */
int bc_pack_reads(uint32_t n_reads, const uint8_t *seq, const uint8_t *qual, const uint64_t *seq_off,
                  const uint32_t *cigar, const uint32_t *cigar_off, uint32_t min_base_quality,
                  uint32_t *seq_woff_out, uint64_t *planes_out, uint32_t *okmask_out,
                  uint32_t *exc_read_out, uint32_t *exc_pos_out, uint32_t exc_cap, uint32_t *n_exc);

/* ---- instrumentation --------------------------------------------------------- */
/* CUDA-event timing on the compute stream (bench.py: kernel-only figures). */
int bc_timer_start(bc_handle *h);
int bc_timer_stop(bc_handle *h, float *ms);
/* Diagnostic: GB/s of `reps` plain cudaMemcpyAsync copies of `bytes` from pinned host memory to the device on the
 * handle's copy stream (CUDA events) -- what the platform gives the end-to-end path of bc_push_batch to work with. */
int bc_h2d_probe(bc_handle *h, uint64_t bytes, int reps, double *gb_per_s);
/* Device time of the last counting-kernel launch alone (events around K1), and launches so far. */
int bc_last_count_kernel_ms(bc_handle *h, float *ms);
/* Device times of the last n counting-kernel launches (ms[0] = most recent; ring of 256).
 * Returns how many were written, or -1. */
int bc_count_kernel_ms_history(bc_handle *h, float *ms, int n);
uint64_t bc_kernel_launches(bc_handle *h);
/* Debug / cross-check: 0 = the lean bit-sliced kernel (csrc/k1_fast.cuh; straight-line decode of reads with
 * at most three CIGAR ops) followed by the general walker (csrc/k1_count.cuh) over the blocks it deferred
 * (default), 1 = one-thread-per-read per-base atomics, 2 = the general walker alone
 * (BASECOUNT_B200_K1=walker makes it what 0 selects).
 * All are CUDA; there is no host path.  Set it before bc_batch_upload: a resident batch keeps
 * the chunking of the variant it was uploaded under. */
int bc_set_count_variant(bc_handle *h, int variant);

/* The packers' CIGAR normal form (see bc_batch.cigar) of n_reads reads: out_off[n_reads + 1] and, when out is not
 * NULL, the words (at most as many as went in).  Returns the number of words, or ~0 on bad arguments. */
uint64_t bc_canonical_cigars(uint32_t n_reads, const uint32_t *cigar, const uint64_t *cigar_off, uint32_t *out,
                             uint32_t *out_off);

/* ---- native BAM decode (host code) -------------------------------------------
 * Replaces what the reference does per read through pysam: open (basecount/main.py:97-99),
 * fetch(until_eof=True) (main.py:127), the read filter `not is_unmapped and mapping_quality >=
 * min_mapping_quality` (main.py:165) and the four per-read attributes (main.py:166-173:
 * reference_start, cigartuples, query_alignment_sequence, query_alignment_qualities -- soft clips
 * trimmed).  The file is inflated block-parallel and indexed once; a selection is copied into
 * caller-owned flat arrays (the inputs of bc_pack_reads) by a pool of threads. */
typedef struct bc_bam bc_bam;
/* threads <= 0: one per host core (at most 32).  Errors: bc_bam_last_error() (thread-local text). */
int bc_bam_open(const char *path, int threads, bc_bam **out);
/* The same file span by span, for files whose inflated size exceeds host memory (the reference's
 * fetch(until_eof=True) loop, main.py:127, is O(chunk) too): bc_bam_stream_next hands out an ordinary bc_bam
 * holding the whole records that start in the next ~max_inflated_bytes of the inflated stream (*out = NULL at the
 * end of the file; the first span is handed out even if it holds no record, so the header is always available
 * through bc_bam_num_refs / bc_bam_ref_name / bc_bam_ref_len).  Close every span with bc_bam_close. */
typedef struct bc_bam_stream bc_bam_stream;
int bc_bam_stream_open(const char *path, int threads, bc_bam_stream **out);
int bc_bam_stream_next(bc_bam_stream *s, uint64_t max_inflated_bytes, bc_bam **out);
void bc_bam_stream_close(bc_bam_stream *s);
const char *bc_bam_last_error(void);
/* The CRC-32 every BGZF block is checked with on the way in (RFC 1952; carry-less-multiply folding where the
 * CPU has it, zlib's crc32 otherwise).  Exported so the decoder's check can be held to zlib's on any bytes. */
uint32_t bc_bgzf_crc32(const uint8_t *data, uint64_t n);
/* The block decoder that runs before zlib (csrc/inflate_fast.h), on its own: raw DEFLATE `in` -> exactly
 * out_len bytes.  Returns 1 if it produced them, 0 if it declined (the readers then let zlib decide).
 * Exported so it can be held to zlib on arbitrary streams, valid and damaged. */
int bc_inflate_raw(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_len);
void bc_bam_close(bc_bam *b);
uint64_t bc_bam_num_records(const bc_bam *b);
uint32_t bc_bam_num_refs(const bc_bam *b);
const char *bc_bam_ref_name(const bc_bam *b, uint32_t i);
uint32_t bc_bam_ref_len(const bc_bam *b, uint32_t i);
/* refID, pos, MAPQ and FLAG of every record in file order (any pointer may be NULL). */
int bc_bam_core(const bc_bam *b, int32_t *ref_id, int32_t *pos, uint8_t *mapq, uint16_t *flag);
/* Selection = records [rec_a, rec_b) with refID == ref_id, FLAG & 4 == 0 and MAPQ >= min_mapq.
 * Sizes first, then fill: starts[n], cigar[n_cigar] (BAM-native len << 4 | op), cigar_off[n + 1],
 * seq[n_bases] (ASCII, soft clips trimmed), qual[n_bases] (phred bytes; may be NULL when
 * min_base_quality is 0 and nothing reads them), seq_off[n + 1]. */
int bc_bam_select_sizes(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                        uint64_t *n_reads, uint64_t *n_cigar, uint64_t *n_bases);
int bc_bam_select_fill(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                       uint32_t *starts, uint32_t *cigar, uint64_t *cigar_off, uint8_t *seq, uint8_t *qual,
                       uint64_t *seq_off);

/* The same selection straight into the packed batch of ONE reference slot (bc_bam_select_fill + bc_pack_reads
 * in one pass: the 4-bit bases go directly to the 2-bit bit-planar words).  out6 = {n_reads, n_cigar, n_words,
 * n_bases, aligned_bases, starts_sorted}.  Arrays: starts[n], cigar[n_cigar], cigar_off[n + 1], seq_woff[n + 1],
 * planes[n_words], okmask[n_words] (NULL unless min_base_quality > 0), exc_read / exc_pos [exc_cap];
 * *n_exc may exceed exc_cap (call again with room).  BC_ERR_READ_OVERRUN as bc_pack_reads. */
int bc_bam_pack_sizes(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq, uint64_t *out6);
int bc_bam_pack_fill(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                     uint32_t min_base_quality, uint32_t *starts, uint32_t *cigar, uint32_t *cigar_off,
                     uint32_t *seq_woff, uint64_t *planes, uint32_t *okmask, uint32_t *exc_read, uint32_t *exc_pos,
                     uint64_t exc_cap, uint64_t *n_exc);

/* ---- index-aware region fetch (host code) -----------------------------------
 * The region-sharded path (config 5) gives each rank the reads that START in its region of the
 * reference (the np.linspace split of tests/test_basecount.py:146-150).  bc_bam_index_build writes
 * the BAI (SAM spec 5.2) of a coordinate-sorted BAM -- the job of `pysam.index`, which the
 * reference's tests call at tests/test_basecount.py:343-344.  bc_bam_open_region reads and inflates
 * only the BGZF blocks that hold the records of reference `ref_id` starting in [beg, end) (0-based,
 * half open) and returns a handle that behaves like bc_bam_open on a file of just those records. */
int bc_bam_index_build(const char *bam_path, const char *bai_path, int threads);
int bc_bam_open_region(const char *bam_path, const char *bai_path, int32_t ref_id, int64_t beg, int64_t end,
                       int threads, bc_bam **out);

/* ---- exact native TSV rows (host code) --------------------------------------
 * Replaces the row loop of basecount/main.py:456-466: one line per position (wide) or per
 * position and base (long), cells joined by tabs, each cell `str(round(x, decimal_places))` --
 * ints stay ints (incl. the zero-coverage sentinels -1 / 1 / 1, from `flags` as in bc_stats),
 * floats are rounded on their exact binary value, ties to even, and printed as their shortest
 * repr.  Exact for 0 <= decimal_places <= 4 and |x| < 1e11; returns BC_ERR_ARG otherwise (callers
 * fall back to Python's own formatting).  counts: n_pos x 6 int64 row-major; pc: k planes of
 * pc_stride doubles; first_pos: 1-based position of row 0.  *text is malloc'ed (rows joined by
 * '\n', no trailing newline, NUL-terminated); release it with bc_free_text. */
int bc_format_tsv(const char *ref_name, uint64_t n_pos, uint64_t first_pos, int k, int long_format, int decimal_places,
                  const int64_t *counts, const int64_t *coverage, const double *pc, uint64_t pc_stride,
                  const double *entropy, const double *secondary, const uint8_t *flags, int threads, char **text,
                  uint64_t *len);
void bc_free_text(char *text);

#ifdef __cplusplus
}
#endif
#endif /* BASECOUNT_B200_H */
