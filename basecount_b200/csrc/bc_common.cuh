// bc_common.cuh -- shared definitions for the sm_100a pileup-counting kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bc {

// Count matrix layout in HBM: six planes (A,C,G,T,DS,N -- count.cpp:16-17), each
// `stride` uint32 columns long.  Reference slot r owns columns
// [col_base[r], col_base[r] + ref_len[r]); col_base is a multiple of 128 so every
// slot starts on a 512-byte boundary and on a 32-column window word.
constexpr int kPlanes = 6;
constexpr int kPlaneDS = 4;
constexpr int kPlaneN = 5;
constexpr uint32_t kColAlign = 128;

// status words written by the kernels (device memory, read back in bc_sync)
enum StatusWord : int {
    kStatMaybeOverflow = 0,   // an M/=/X piece crossed ref_len: run the exact check
    kStatIndexError = 1,      // a counted event at refPos >= ref_len  (count.cpp .at())
    kStatDeferredReads = 2,   // reads k1_count_fast left to the general walker (diagnostic; sizes the walker's grid)
    kStatWords = 4
};

struct Chunk {                // one warp's share of a batch: consecutive reads of ONE slot
    uint32_t read_begin;
    uint32_t read_end;
    uint32_t col_base;
    uint32_t ref_len;
};

struct BatchView {            // device pointers of one packed batch (see bc_batch in the header)
    uint32_t n_reads;
    uint32_t n_refs;
    const uint32_t *ref_read_off;
    const uint32_t *starts;
    const uint32_t *cigar_off;
    const uint32_t *cigar;
    const uint32_t *seq_woff;
    const uint2 *planes;      // .x = low bit plane, .y = high bit plane
    const uint32_t *okmask;
    uint32_t n_exc;
    const uint32_t *exc_read;
    const uint32_t *exc_pos;
};

struct CountView {            // the accumulators
    uint32_t *counts;         // kPlanes * stride
    uint64_t stride;
    const uint32_t *col_base; // n_refs
    const uint32_t *ref_len;  // n_refs
    uint32_t *status;         // kStatWords
};

__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == 0u || op == 7u || op == 8u; }
__device__ __forceinline__ bool op_is_refskip(uint32_t op) { return op == 2u || op == 3u; }

// slot (reference) of read i: largest r with ref_read_off[r] <= i
__device__ __forceinline__ uint32_t slot_of_read(const uint32_t *ref_read_off, uint32_t n_refs, uint32_t i)
{
    uint32_t lo = 0, hi = n_refs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (ref_read_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace bc
