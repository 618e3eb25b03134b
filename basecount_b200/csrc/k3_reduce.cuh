// k3_reduce.cuh -- K3: amplicon (BED window) mean / median vectors.
//
// Replaces the O(tiles x L) Python membership loop and the per-tile np.mean / np.median
// of the --summarise-with-bed block (basecount/main.py:519-551).  One CTA per
// (window, metric): the window's values are staged in shared memory, sorted with a
// bitonic network (exact selection: the median of an even count is the mean of the two
// middle values, as np.median does), and summed in numpy's own pairwise order so the
// mean is bit-identical to np.mean for windows that fit the staging buffer.
// Windows larger than the buffer fall back to an 8-pass radix select over HBM.
#pragma once
#include "bc_common.cuh"
#include <math.h>

namespace bc {

constexpr int kK3Threads = 256;

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, pairwise_sum):
// < 8 elements: running sum from -0.0; <= 128: eight strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail; otherwise split at n/2 rounded down
// to a multiple of 8.  np.mean over a list is np.add.reduce over the float64 array / n.
__device__ __forceinline__ double np_pairwise_leaf(const double *a, uint32_t n)   // n <= 128
{
    if (n < 8) {
        double r = -0.0;
        for (uint32_t i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    uint32_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

// The recursion "sum(a, n) = sum(a, n2) + sum(a + n2, n - n2)" unrolled onto an explicit
// stack (device recursion would need a run-time stack size).  Depth <= log2(n / 64) < 32.
__device__ double np_pairwise(const double *a, uint32_t n)
{
    uint32_t off_s[32], n_s[32];
    double left_s[32];
    uint8_t state_s[32];                    // 0: fresh, 1: left child pending, 2: right child pending
    int sp = 0;
    off_s[0] = 0;
    n_s[0] = n;
    state_s[0] = 0;
    double ret = 0.0;
    for (;;) {
        const uint32_t off = off_s[sp], m = n_s[sp];
        if (state_s[sp] == 0) {
            if (m <= 128) {
                ret = np_pairwise_leaf(a + off, m);
                if (sp == 0) return ret;
                sp--;
                continue;
            }
            uint32_t n2 = m / 2;
            n2 -= n2 % 8;
            state_s[sp] = 1;
            sp++;
            off_s[sp] = off;
            n_s[sp] = n2;
            state_s[sp] = 0;
        } else if (state_s[sp] == 1) {
            left_s[sp] = ret;
            uint32_t n2 = m / 2;
            n2 -= n2 % 8;
            state_s[sp] = 2;
            sp++;
            off_s[sp] = off + n2;
            n_s[sp] = m - n2;
            state_s[sp] = 0;
        } else {
            ret = left_s[sp] + ret;
            if (sp == 0) return ret;
            sp--;
        }
    }
}

__device__ __forceinline__ double window_value(int metric, const long long *cov, const double *ent,
                                               const double *sec, uint32_t pos)
{
    if (metric == 0) return (double)cov[pos];
    return metric == 1 ? ent[pos] : sec[pos];
}

__device__ __forceinline__ unsigned long long order_key(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k)
{
    unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}

// k-th smallest (0-based) of the window, by radix select over HBM (CTA-wide).
__device__ double radix_select(int metric, const long long *cov, const double *ent, const double *sec,
                               uint32_t a, uint32_t n, uint32_t kth, uint32_t *hist /* 256 shared */)
{
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_k;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_k = kth; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = order_key(window_value(metric, cov, ent, sec, a + i));
            if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t k = s_k, d = 0;
            for (; d < 256; d++) {
                if (k < hist[d]) break;
                k -= hist[d];
            }
            s_k = k;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    return key_to_double(s_prefix);
}

// grid = (n_tiles, 3); dynamic shared memory = cap doubles (cap a power of two).
__global__ void __launch_bounds__(kK3Threads)
k3_amplicons(const long long *__restrict__ cov, const double *__restrict__ ent, const double *__restrict__ sec,
             uint32_t L, const int32_t *__restrict__ lo, const int32_t *__restrict__ hi, uint32_t n_tiles,
             uint32_t cap, double *__restrict__ out, uint8_t *__restrict__ empty)
{
    extern __shared__ double k3_vals[];
    __shared__ uint32_t hist[256];
    __shared__ double red[kK3Threads];
    const uint32_t tile = blockIdx.x;
    const int metric = blockIdx.y;
    // members: 0-based positions j with start <= j <= end (main.py:523)
    const long long a = max((long long)lo[tile], 0ll);
    const long long b = min((long long)hi[tile], (long long)L - 1);
    double *mean_out = out + (uint64_t)(2 * metric) * n_tiles + tile;
    double *median_out = out + (uint64_t)(2 * metric + 1) * n_tiles + tile;
    if (a > b) {                                   // empty window -> int -1 (main.py:531-533)
        if (threadIdx.x == 0) {
            *mean_out = -1.0;
            *median_out = -1.0;
            if (metric == 0) empty[tile] = 1;
        }
        return;
    }
    const uint32_t n = (uint32_t)(b - a + 1);
    if (threadIdx.x == 0 && metric == 0) empty[tile] = 0;

    if (n <= cap) {
        uint32_t np2 = 1;
        while (np2 < n) np2 <<= 1;
        for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x)
            k3_vals[i] = i < n ? window_value(metric, cov, ent, sec, (uint32_t)a + i) : INFINITY;
        __syncthreads();
        if (threadIdx.x == 0) *mean_out = np_pairwise(k3_vals, n) / (double)n;     // before sorting: input order
        __syncthreads();
        for (uint32_t k = 2; k <= np2; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                    const uint32_t ixj = i ^ j;
                    if (ixj > i) {
                        const double x = k3_vals[i], y = k3_vals[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { k3_vals[i] = y; k3_vals[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0)
            *median_out = (n & 1u) ? k3_vals[n / 2] : (k3_vals[n / 2 - 1] + k3_vals[n / 2]) / 2.0;
        return;
    }
    // large window: block-tree sum (not numpy's order; < 1e-15 relative) + radix select
    double s = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s += window_value(metric, cov, ent, sec, (uint32_t)a + i);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int d = kK3Threads / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    const double total = red[0];
    const double m1 = radix_select(metric, cov, ent, sec, (uint32_t)a, n, n / 2, hist);
    double med = m1;
    if ((n & 1u) == 0u) {
        const double m0 = radix_select(metric, cov, ent, sec, (uint32_t)a, n, n / 2 - 1, hist);
        med = (m0 + m1) / 2.0;
    }
    if (threadIdx.x == 0) {
        *mean_out = total / (double)n;
        *median_out = med;
    }
}

}  // namespace bc
