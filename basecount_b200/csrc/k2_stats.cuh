// k2_stats.cuh -- K2: fused per-position statistics, and the --summarise reduction.
//
// Replaces get_stats / get_entropy (basecount/main.py:10-53) and the three reductions
// of the summarise block (main.py:479-485).  One thread per reference position reads the
// six count planes (coalesced, 4 B per lane per plane) and produces coverage, the K
// percentages, normalised entropy and secondary entropy in float64 with the reference's
// exact operation order:
//     p_i  = c_i / coverage                         (main.py:40)
//     pc_i = 100 * p_i                              (main.py:41)  -- divide, then scale
//     H    = norm * sum_i( -(p_i * log2 p_i) )      (main.py:11,42)
// `sum` is Python's builtin: on CPython >= 3.12 a Neumaier-compensated float sum that
// starts from int 0; neumaier_sum() below restates it so the only difference left between
// the two implementations is log2 itself (CUDA libdevice vs glibc, both < 1 ulp).
// Compiled with -fmad=false so nothing is contracted.
#pragma once
#include "bc_common.cuh"

namespace bc {

struct PosStats {
    long long cov;
    double pc[6];
    double ent, sec;
    uint32_t flags;     // bit0: coverage == 0, bit1: secondary coverage == 0
};

__device__ __forceinline__ double neumaier_entropy(const long long *c, int n, long long total, int skip)
{
    double hi = 0.0, lo = 0.0;
    const double tot = (double)total;
    for (int i = 0; i < n; i++) {
        if (i == skip || c[i] == 0) continue;            // p == 0 contributes int 0 (main.py:11)
        if (c[i] == total) continue;                     // p == 1: -(1 * log2 1) = -0.0, and x + -0.0 == x exactly
        const double p = (double)c[i] / tot;
        const double x = -(p * log2(p));
        const double t = hi + x;
        if (fabs(hi) >= fabs(x)) lo += (hi - t) + x; else lo += (x - t) + hi;
        hi = t;
    }
    if (lo != 0.0 && isfinite(lo)) return hi + lo;
    return hi;
}

// log2 of the integers below kLog2Tab, filled once per handle with the same log2() the exact path calls.
constexpr int kLog2Tab = 16384;
// ... followed by the exact entropy terms -(p * log2 p), p = c / cov, for every 0 < c < cov < kTermTab (the
// very expression of main.py:11, so a looked-up term is bit-identical to a computed one): shallow coverage
// (whole-genome 30x) then needs no division and no log2 at all.
constexpr int kTermTab = 128;
constexpr int kSummaryTabDoubles = kLog2Tab + kTermTab * kTermTab;
__global__ void k_fill_log2(double *__restrict__ tab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kLog2Tab) {
        tab[i] = i ? log2((double)i) : 0.0;
    } else if (i < kSummaryTabDoubles) {
        const int cov = (i - kLog2Tab) / kTermTab, c = (i - kLog2Tab) % kTermTab;
        double x = 0.0;
        if (c > 0 && c < cov) {
            const double p = (double)c / (double)cov;
            x = -(p * log2(p));
        }
        tab[i] = x;
    }
}

// Entropy for the --summarise reductions (sums of entropies only; K2 is bound by the FP64 pipe and nearly all
// of that was log2).  The largest class of a position keeps the exact expression of main.py:11 -- its p is close
// to 1, where a difference of logs would cancel -- and it is computed ONCE, outside the class loop, so the lanes
// of a warp call log2 together whatever letter their largest class is.  Every other class has 2c <= coverage
// (|log2 p| >= 1), and log2(c / cov) is taken as tab[c] - tab[cov]: two table loads, relative error of the
// term <= ~4e-15.  The terms are still added in class order with the Neumaier compensation of CPython's sum().
// Coverage beyond the table takes the exact path for every class.  Per-position outputs (k2_stats_rows) never
// use the table.
__device__ __forceinline__ double neumaier_entropy_tab(const long long *c, int n, long long total,
                                                       const double *__restrict__ tab)
{
    if (total >= (long long)kLog2Tab) return neumaier_entropy(c, n, total, -1);
    if (total < (long long)kTermTab) {                     // shallow: every term straight from the table
        const double *__restrict__ row = tab + kLog2Tab + (int)total * kTermTab;
        double hi = 0.0, lo = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            if (i >= n || c[i] == 0 || c[i] == total) continue;
            const double x = row[c[i]];
            const double t = hi + x;
            if (fabs(hi) >= fabs(x)) lo += (hi - t) + x; else lo += (x - t) + hi;
            hi = t;
        }
        if (lo != 0.0 && isfinite(lo)) return hi + lo;
        return hi;
    }
    int top = 0;
    long long ctop = c[0];
#pragma unroll
    for (int i = 1; i < 6; i++)
        if (i < n && c[i] > ctop) {
            ctop = c[i];
            top = i;
        }
    if (ctop == total) return 0.0;                         // one class only: -(1 * log2 1) contributes -0.0
    const double tot = (double)total;
    const double ptop = (double)ctop / tot;
    const double xtop = -(ptop * log2(ptop));
    const double ltot = tab[total];
    double hi = 0.0, lo = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i >= n || c[i] == 0) continue;
        double x = xtop;
        if (i != top) {
            const double p = (double)c[i] / tot;
            x = -(p * (tab[c[i]] - ltot));
        }
        const double t = hi + x;
        if (fabs(hi) >= fabs(x)) lo += (hi - t) + x; else lo += (x - t) + hi;
        hi = t;
    }
    if (lo != 0.0 && isfinite(lo)) return hi + lo;
    return hi;
}

// The same for uint32 counts (no int64 fold: a position's coverage is at most the reads since the last fold, < 2^32)
// and without the compensated sum: a SUM of entropies tolerates the ~1e-16 that plain addition of <= 6 positive
// terms and one reciprocal (c * (1 / cov) instead of c / cov for the minor classes) cost, and the kernel was bound
// by instruction issue -- 260 instructions per position at whole-genome depth, mostly 64-bit integer compares,
// branches around empty classes and the compensation.  Terms of empty classes and of a class that holds the whole
// coverage are 0.0 in both tables (k_fill_log2), so nothing is skipped and nothing branches.
// cov in (0, kLog2Tab).
__device__ __forceinline__ double summary_entropy_u32(const uint32_t (&c)[6], int K, uint32_t cov, const double *__restrict__ tab)
{
    if (cov < (uint32_t)kTermTab) {                        // shallow: every term straight from the table
        const double *__restrict__ row = tab + kLog2Tab + cov * (uint32_t)kTermTab;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i < K) s += row[c[i]];
        return s;
    }
    uint32_t ctop = c[0];
#pragma unroll
    for (int i = 1; i < 6; i++)
        if (i < K) ctop = max(ctop, c[i]);
    const double tot = (double)cov, inv = 1.0 / tot, ltot = tab[cov];
    double s = 0.0;
    bool top_pending = true;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i >= K) continue;
        if (top_pending && c[i] == ctop) {                 // the first maximum keeps the exact expression of main.py:11
            top_pending = false;
            if (ctop != cov) {
                const double p = (double)ctop / tot;
                s += -(p * log2(p));
            }
        } else if (c[i] != 0u) {                           // 2c <= cov: |log2 p| >= 1, no cancellation in the difference
            s += -(((double)c[i] * inv) * (tab[c[i]] - ltot));
        }
    }
    return s;
}

// Coverage and entropy only (what the --summarise reductions need, main.py:479-485).
__device__ __forceinline__ void position_cov_entropy(const long long *c, int K, double norm, long long &cov_out,
                                                     double &ent_out, const double *__restrict__ tab)
{
    long long cov = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (i < K) cov += c[i];
    cov_out = cov;
    ent_out = cov == 0 ? 1.0 : norm * neumaier_entropy_tab(c, K, cov, tab);
}

__device__ __forceinline__ void position_stats(const long long *c, int K, double norm, double norm2, PosStats &o)
{
    long long cov = 0;
    for (int i = 0; i < K; i++) cov += c[i];                          // main.py:37
    o.cov = cov;
    o.flags = 0;
    if (cov == 0) {                                                   // main.py:34-36
        for (int i = 0; i < K; i++) o.pc[i] = -1.0;
        o.ent = 1.0;
        o.sec = 1.0;
        o.flags = 3u;
        return;
    }
    const double dcov = (double)cov;
    for (int i = 0; i < K; i++) o.pc[i] = 100.0 * ((double)c[i] / dcov);
    o.ent = norm * neumaier_entropy(c, K, cov, -1);
    int top = 0;
    for (int i = 1; i < K; i++) if (c[i] > c[top]) top = i;           // first maximum, np.argmax (main.py:45)
    const long long rest = cov - c[top];
    if (rest == 0) {                                                  // main.py:47
        o.sec = 1.0;
        o.flags = 2u;
    } else {
        o.sec = norm2 * neumaier_entropy(c, K, rest, top);            // main.py:48-53
    }
}

__device__ __forceinline__ void load_counts(const uint32_t *__restrict__ c32, const unsigned long long *__restrict__ c64,
                                            uint64_t stride, uint64_t col, int K, long long *c)
{
#pragma unroll
    for (int p = 0; p < kPlanes; p++) {
        if (p < K) {
            const uint64_t a = (uint64_t)p * stride + col;
            c[p] = (long long)c32[a] + (c64 ? (long long)c64[a] : 0ll);
        }
    }
}

// Per-position outputs for one slot (TSV / records / amplicon inputs).  Any output may be NULL.
__global__ void __launch_bounds__(256)
k2_stats_rows(const uint32_t *__restrict__ c32, const unsigned long long *__restrict__ c64, uint64_t stride,
              uint64_t col_base, uint32_t ref_len, int K, double norm, double norm2,
              long long *__restrict__ coverage, double *__restrict__ pc, double *__restrict__ entropy,
              double *__restrict__ secondary, uint8_t *__restrict__ flags)
{
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= ref_len) return;
    long long c[6];
    load_counts(c32, c64, stride, col_base + pos, K, c);
    PosStats s;
    position_stats(c, K, norm, norm2, s);
    if (coverage) coverage[pos] = s.cov;
    if (pc) for (int i = 0; i < K; i++) pc[(uint64_t)i * ref_len + pos] = s.pc[i];
    if (entropy) entropy[pos] = s.ent;
    if (secondary) secondary[pos] = s.sec;
    if (flags) flags[pos] = (uint8_t)s.flags;
}

struct SummaryPartial {
    long long nonzero;
    long long cov_sum;
    double ent_sum;
};

// Partials per slot: a fixed function of the slot length (-> deterministic reduction order): one CTA per
// kSummaryPerCta positions (4 per thread: one resident wave for a batch of viral samples), up to kSummaryMaxBlocks CTAs per slot.
constexpr int kSummaryMaxBlocks = 2048;
#ifndef BC_K2_PER_CTA
#define BC_K2_PER_CTA 1024
#endif
constexpr uint32_t kSummaryPerCta = BC_K2_PER_CTA;
__host__ __device__ inline uint32_t summary_blocks(uint32_t ref_len)
{
    const uint32_t b = (ref_len + kSummaryPerCta - 1u) / kSummaryPerCta;
    return b < 1u ? 1u : (b > (uint32_t)kSummaryMaxBlocks ? (uint32_t)kSummaryMaxBlocks : b);
}

// Sum (nz, cs, es) over the CTA in a fixed order: xor-shuffle tree inside each warp, then thread 0 adds the
// eight warp totals in warp order.  Result valid in thread 0.
__device__ __forceinline__ void summary_cta_sum(long long &nz, long long &cs, double &es, long long *s_nz, long long *s_cs,
                                                double *s_es)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nz += __shfl_xor_sync(0xFFFFFFFFu, nz, d);
        cs += __shfl_xor_sync(0xFFFFFFFFu, cs, d);
        es += __shfl_xor_sync(0xFFFFFFFFu, es, d);
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();                                       // earlier readers of the staging arrays are done
    if ((threadIdx.x & 31) == 0) {
        s_nz[w] = nz;
        s_cs[w] = cs;
        s_es[w] = es;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        nz = s_nz[0];
        cs = s_cs[0];
        es = s_es[0];
        for (int i = 1; i < 8; i++) {
            nz += s_nz[i];
            cs += s_cs[i];
            es += s_es[i];
        }
    }
}

// Fused stats + summarise reduction for ALL slots in ONE launch: grid = (max blocks over slots, n_refs);
// partials of slot r sit at part_off[r] .. part_off[r] + summary_blocks(ref_len[r]).  The CTA of a slot that
// arrives last (arrive[r], reset for the next launch) sums the slot's partials: every thread its partials in
// index order, then the fixed-order CTA sum -- the result does not depend on which CTA was last.
// min_cov < 0: nonzero = positions with coverage != 0, ent_sum over all positions; min_cov >= 0: nonzero =
// positions with coverage >= min_cov, ent_sum over those (BaseCount.mean_entropy, main.py:342-359).
// FINAL = false: the kernel stops at the per-CTA partials (written to `partials`, which then lies in the result arena
// the host fetches at the next synchronisation) and the HOST adds a slot's partials in index order -- a launch then
// has no __threadfence / atomic round trip per CTA and no last-CTA phase, which were a fifth of its duration on the
// viral-sample shape.  FINAL = true keeps the sums on the device (the all-reduce over ranks needs them there).
template <bool HAS64, bool FINAL>
__global__ void __launch_bounds__(256)
k2_summary(const uint32_t *__restrict__ c32, const unsigned long long *__restrict__ c64, uint64_t stride,
           const uint32_t *__restrict__ col_base, const uint32_t *__restrict__ ref_len, int K, double norm,
           long long min_cov, const double *__restrict__ log2_tab, const uint32_t *__restrict__ part_off,
           SummaryPartial *partials, uint32_t *arrive, long long *__restrict__ nonzero, long long *__restrict__ cov_sum,
           double *__restrict__ ent_sum)
{
    const uint32_t r = blockIdx.y;
    const uint32_t L = ref_len[r];
    const uint32_t nb = summary_blocks(L);
    if (blockIdx.x >= nb) return;
    const uint64_t base = col_base[r];
    long long nz = 0, cs = 0;
    double es = 0.0;
    const uint32_t *__restrict__ const col0 = c32 + base;
    // Four positions per trip with all their count loads issued first: a thread's positions are independent, and one
    // after the other each paid the full load -> table look-up latency chain (the kernel is latency-bound on the
    // viral-sample shape: 4 positions per thread, 5 warps per scheduler).  Same positions in the same order.
    constexpr int kPP = 4;
    const uint32_t pstride = nb * blockDim.x;
    for (uint32_t pos0 = blockIdx.x * blockDim.x + threadIdx.x; pos0 < L; pos0 += kPP * pstride) {
      uint32_t ub[kPP][6];
      if (!HAS64) {
#pragma unroll
          for (int j = 0; j < kPP; j++) {
              const uint32_t pj = pos0 + (uint32_t)j * pstride;
#pragma unroll
              for (int p = 0; p < kPlanes; p++) ub[j][p] = (p < K && pj < L && pj >= pos0) ? col0[(uint64_t)p * stride + pj] : 0u;
          }
      }
#pragma unroll
      for (int j = 0; j < kPP; j++) {
        const uint32_t pos = pos0 + (uint32_t)j * pstride;
        if (pos >= L || pos < pos0) continue;
        long long cov;
        double ent;
        if (!HAS64) {
            uint32_t u[6], ucov = 0u;
#pragma unroll
            for (int p = 0; p < kPlanes; p++) {
                u[p] = ub[j][p];
                ucov += u[p];
            }
            cov = (long long)ucov;
            if (ucov == 0u) {
                ent = 1.0;                                 // main.py:34-36
            } else if (ucov < (uint32_t)kLog2Tab) {
                ent = norm * summary_entropy_u32(u, K, ucov, log2_tab);
            } else {                                       // beyond the tables: the exact expression for every class
                long long c[6];
#pragma unroll
                for (int p = 0; p < kPlanes; p++) c[p] = (long long)u[p];
                ent = norm * neumaier_entropy(c, K, cov, -1);
            }
        } else {
            long long c[6];
#pragma unroll
            for (int p = 0; p < kPlanes; p++) {
                c[p] = 0;
                if (p < K) {
                    const uint64_t a = (uint64_t)p * stride + base + pos;
                    c[p] = (long long)c32[a] + (long long)c64[a];
                }
            }
            position_cov_entropy(c, K, norm, cov, ent, log2_tab);
        }
        if (min_cov < 0) {                                 // --summarise (main.py:479-485)
            nz += (cov != 0);
            es += ent;
        } else if (cov >= min_cov) {                       // mean_entropy(min_coverage) (main.py:342-359)
            nz += 1;
            es += ent;
        }
        cs += cov;
      }
    }
    __shared__ long long s_nz[8], s_cs[8];
    __shared__ double s_es[8];
    __shared__ uint32_t s_last;
    summary_cta_sum(nz, cs, es, s_nz, s_cs, s_es);
    SummaryPartial *p = partials + part_off[r];
    if (threadIdx.x == 0) {
        p[blockIdx.x].nonzero = nz;
        p[blockIdx.x].cov_sum = cs;
        p[blockIdx.x].ent_sum = es;
        if (FINAL) {
            __threadfence();                               // the partial is visible before the arrival is
            s_last = atomicAdd(arrive + r, 1u) == nb - 1u;
        }
    }
    if (!FINAL) return;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile SummaryPartial *vp = p;                 // written by other CTAs: no cached copies
    nz = 0;
    cs = 0;
    es = 0.0;
    for (uint32_t i = threadIdx.x; i < nb; i += 256u) {
        nz += vp[i].nonzero;
        cs += vp[i].cov_sum;
        es += vp[i].ent_sum;
    }
    summary_cta_sum(nz, cs, es, s_nz, s_cs, s_es);
    if (threadIdx.x == 0) {
        nonzero[r] = nz;
        cov_sum[r] = cs;
        ent_sum[r] = es;
        arrive[r] = 0u;
    }
}

}  // namespace bc
