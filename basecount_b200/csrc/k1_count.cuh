// k1_count.cuh -- K1: CIGAR walk + per-position base counting on sm_100a.
//
// Replaces the loop nest of the reference operator (basecount/count.cpp:22-97).
//
// Design (B200-first, nothing here mirrors the reference's scalar loop).  The kernel is bound by
// the ALU pipe (LOP3 / SHF issue at one warp instruction per two cycles per SM sub-partition), so
// everything is arranged to spend as few logic instructions per aligned base as possible:
//   * A warp owns a chunk of consecutive reads of one reference slot.  Its 32 lanes form
//     S = 32/G "read slots" of G lanes; a lane covers kW = 2 consecutive 32-column window words,
//     so all slots share ONE window of 64*G reference columns and window moves / flushes are
//     warp-uniform.
//   * Sequence data is 2-bit, bit-planar (lo / hi plane), staged in shared memory by 1-D TMA bulk
//     copies three blocks of reads ahead.  A funnel shift aligns 32 bases of a read onto a window
//     word; no per-base work anywhere.
//   * CIGAR walk: one read per lane, all lanes step through their own CIGAR in lock step
//     (count.cpp:40-96).  Every M/=/X run that fits the window becomes a 16-byte *piece* entry
//     {first column, end column, bit offset of the data} in a shared-memory ring; D/N runs add to
//     the DS plane directly; I advances the read.  Runs that do not fit wait (their lane stalls)
//     until the window moves; runs longer than a window are pushed window by window.
//   * Counting: the ring is consumed four pieces per read slot at a time, straight-line code with
//     no votes or branches inside.  Per window word four quantities are counted VERTICALLY in
//     bit-sliced carry-save counters (Harley-Seal): n(lo), n(hi), n(lo & hi), n(valid), each masked
//     with the piece's column range (a 64-bit mask looked up in shared memory).  That is ~2.1
//     LOP3 per counted quantity per 32 bases; letters follow at flush time:
//     T = n(lo&hi), C = n(lo) - T, G = n(hi) - T, A = n(valid) - n(lo) - n(hi) + T.
//   * A flush (every 252 pieces per slot, or on a window move) converts the counters to integers
//     (byte-packed extraction, 16-bit slot sums, shared-memory staging) and adds them to the HBM
//     count planes with coalesced 128-byte RED.ADD -- one atomic per column per flush instead of
//     one per base.
//   * Letters the 2-bit code cannot express (N, IUPAC, lower case) are a sparse correction pass
//     (k1_exceptions); reads whose data was not staged (far longer than the batch mean) are copied
//     into a free stage segment by segment and take the same path.
#pragma once
#include "bc_common.cuh"

namespace bc {

#ifndef BC_K1_SEQCAP
#define BC_K1_SEQCAP 448
#define BC_K1_CIGCAP 96
#endif
#ifndef BC_K1_WARPS
#define BC_K1_WARPS 4
#define BC_K1_MINCTAS 3
#endif
#ifndef BC_K1_META_SMEM
#define BC_K1_META_SMEM 0       // 1: block metadata through cp.async into a shared-memory ring instead of registers
#endif
constexpr int kK1WarpsPerCta = BC_K1_WARPS;
constexpr int kK1MinCtas = BC_K1_MINCTAS;
constexpr int kK1Threads = kK1WarpsPerCta * 32;
constexpr int kNB = 8;                        // bit planes per vertical counter (counts to 255)
constexpr int kNR = 5;                        // planes 0..4 live in registers; planes 5..7 and the weight-16 pending
                                              // carry are touched every 8th / 4th trip only and live in shared memory
constexpr uint32_t kSpillWords = 32;          // per lane: 8 counters x (pending + 3 planes); aliases the flush rows
constexpr int kNC = 4;                        // counted quantities per window word: lo, hi, lo&hi, valid
constexpr int kW = 2;                         // window words per lane
constexpr int kStages = 3;                    // TMA pipeline depth (blocks of reads)
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kSeqCap = BC_K1_SEQCAP;             // staged 64-bit plane words per stage (31 reads x 13 words + slack)
constexpr uint32_t kCigCap = BC_K1_CIGCAP;              // staged CIGAR words per stage; the rest is read from HBM
constexpr uint32_t kCntMax = 252;             // pieces per slot between flushes (8-plane counters, 4 per trip)
constexpr uint32_t kFlushStride = 34;         // uint16 per staged window word (32 + 2 pad: conflict-free stores)
constexpr uint32_t kMaxRpb = 31;              // reads per block: lane i+1 holds the end offsets of read i
constexpr uint32_t kLaneSkipMax = 32;         // longest D/N run a single lane adds itself
constexpr uint32_t kOpClass = 0x140F9u;       // 2 bits per CIGAR op: 1 = M/=/X, 2 = I, 3 = D/N, 0 = S/H/P/B/other

static_assert(kW == 2, "the piece masks are 64-bit (uint2)");

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b)
{
    uint32_t t = a + b;
    return t < a ? 0xFFFFFFFFu : t;
}

// ---- TMA (1-D bulk async copy) + mbarrier helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- geometry + shared-memory layout --------------------------------------------------------
// Per CTA: the 65-entry mask table, then one region per warp.  Inside a warp's region the
// sequence stages sit between the ring / flush rows (below) and the CIGAR stages (above), so the
// unclamped word index of a piece (up to 64*G columns before or after its data) always lands in
// the CTA's own shared memory; whatever it reads there is masked away.
template <int G, bool HAS_OK, uint32_t RING = 0>
struct K1Cfg {
    static constexpr int S = 32 / G;                        // read slots per warp
    static constexpr int Q = 4 * S;                         // ring entries consumed per trip
    static constexpr uint32_t kRing = RING ? RING : (Q >= 32 ? 128u : 64u); // piece ring entries (uint4 each): a leftover + one block
    static constexpr uint32_t kWin = 32u * kW * G;          // window columns
    static constexpr uint32_t kMaxFit = kWin - 31u;         // a piece this long fits a fresh window at any alignment
    static constexpr uint32_t kCols = kFlushStride * kW * G;   // uint16 per flush row
    static constexpr uint32_t lut_bytes = 528;              // 65 x uint2, padded to 16 B
    static constexpr uint32_t ring_off = 0;
    static constexpr uint32_t frow_off = ring_off + kRing * 16u;
    static constexpr uint32_t frow_bytes = kNC * kCols * 2u > kSpillWords * 128u ? kNC * kCols * 2u : kSpillWords * 128u;
    static constexpr uint32_t seq_off = frow_off + frow_bytes;
    static constexpr uint32_t ok_off = seq_off + kStages * kSeqCap * 8u;
    static constexpr uint32_t cig_off = ok_off + (HAS_OK ? kStages * kSeqCap * 4u : 0u);
    static constexpr uint32_t bar_off = cig_off + kStages * kCigCap * 4u;
    static constexpr uint32_t rng_off = bar_off + 32u;      // per stage: what it holds (uint4)
    static constexpr uint32_t meta_off = rng_off + kStages * 16u;   // BC_K1_META_SMEM: kStages x {cigar_off, seq_woff, start} x 32 lanes
    static constexpr uint32_t warp_bytes = meta_off + (BC_K1_META_SMEM ? kStages * 384u : 0u);
    static constexpr uint32_t cta_bytes = lut_bytes + kK1WarpsPerCta * warp_bytes;
    static_assert(seq_off % 16 == 0 && ok_off % 16 == 0 && cig_off % 16 == 0 && bar_off % 16 == 0 && warp_bytes % 16 == 0,
                  "TMA destinations are 16-byte aligned");
    static_assert(seq_off >= (kWin / 32 + 4) * 8 && kStages * kCigCap * 4u >= (kWin / 32 + 4) * 8,
                  "guard bands around the sequence stages");
};
template <int G, bool HAS_OK>
__host__ __device__ constexpr uint32_t k1_cta_smem_bytes()
{
    return K1Cfg<G, HAS_OK>::cta_bytes;
}

// Unconditional: a window that overhangs the reference end adds zeros to the slack columns behind
// it (bc_begin pads every plane by a full window), which is cheaper than testing every value.
__device__ __forceinline__ void red_add(uint32_t *p, uint32_t v)
{
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Convert the warp's vertical counters to integers and add them to the HBM planes.
//   1. every lane rebuilds its eight full 8-plane counters (planes 5..7 and the weight-16 pending
//      carry come back from shared memory; pending carries are folded in);
//   2. the S read slots that cover the same window words add their counters IN BIT-SLICED FORM
//      (a ripple-carry adder over the planes, operands exchanged by shuffle), each keeping half
//      of the counters per round, so after log2(S) rounds a lane holds 8/S counters of the
//      window totals and only those are extracted;
//   3. byte-packed extraction (the loop over the shift amount jj is deliberately NOT unrolled:
//      registers stay statically indexed while the code stays small), staged as uint16 rows;
//   4. letters from the four counted quantities, written with coalesced RED.ADD (128 B per
//      warp instruction).
//   frow: kNC rows x (kW*G window words x kFlushStride) uint16, aliasing the spill words
//   counts: plane A at window column 0
template <int G>
__device__ __forceinline__ void flush_counters(uint32_t (&pl)[kW][kNC][kNR], uint32_t (&pa)[kW][kNC],
                                               uint32_t (&pb)[kW][kNC], uint32_t cnt, uint16_t *frow,
                                               uint32_t *__restrict__ counts, uint64_t stride, int lane)
{
    constexpr int S = 32 / G;
    constexpr int R = S == 8 ? 3 : (S == 4 ? 2 : (S == 2 ? 1 : 0));      // combine rounds
    constexpr int kCols = (int)kFlushStride * kW * G;
    constexpr int kN = kW * kNC;                                          // counters per lane (8)
    const int slot = lane / G, wl = lane % G;
    const uint32_t *sp = reinterpret_cast<const uint32_t *>(frow);
    uint32_t A[kN][kNB + 3];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
            const int i = w * kNC + k;
#pragma unroll
            for (int p = 0; p < kNR; p++) A[i][p] = pl[w][k][p];
#pragma unroll
            for (int p = kNR; p < kNB; p++) A[i][p] = cnt >= 32u ? sp[(i * 4 + (p - kNR + 1)) * 32 + lane] : 0u;
            // pending carries of weight 4 / 8 / 16 are live iff that bit of cnt is set
            uint32_t c = (cnt & 4u) ? pa[w][k] : 0u;
            {
                const uint32_t t = A[i][2] & c;
                A[i][2] ^= c;
                c = t;
            }
            {
                const uint32_t d = (cnt & 8u) ? pb[w][k] : 0u;            // two carries into plane 3: full adder
                const uint32_t t = maj3(A[i][3], c, d);
                A[i][3] ^= c ^ d;
                c = t;
            }
            {
                const uint32_t d = (cnt & 16u) ? sp[(i * 4) * 32 + lane] : 0u;
                const uint32_t t = maj3(A[i][4], c, d);
                A[i][4] ^= c ^ d;
                c = t;
            }
#pragma unroll
            for (int p = 5; p < kNB; p++) {
                const uint32_t t = A[i][p] & c;
                A[i][p] ^= c;
                c = t;
            }
#pragma unroll
            for (int p = kNB; p < kNB + 3; p++) A[i][p] = 0u;
        }
    }
    __syncwarp();                                     // all spill words are read before the rows are written
    // ---- bit-sliced sum over the read slots
    int first = 0;                                    // counter index (w * kNC + k) of A[0] after the rounds
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int d = G << r;
        const int h = kN >> (r + 1);
        const bool up = (slot >> r) & 1;
        if (up) first += h;
#pragma unroll
        for (int i = 0; i < h; i++) {
            uint32_t carry = 0u;
#pragma unroll
            for (int p = 0; p < kNB + r; p++) {
                const uint32_t mine = up ? A[h + i][p] : A[i][p];
                const uint32_t give = up ? A[i][p] : A[h + i][p];
                const uint32_t got = __shfl_xor_sync(kFull, give, d);
                A[i][p] = mine ^ got ^ carry;
                carry = maj3(mine, got, carry);
            }
            A[i][kNB + r] = carry;
        }
    }
    constexpr int kLeft = kN >> R;                    // counters this lane extracts
    const bool high = cnt * (uint32_t)S >= 16u;       // planes 4.. can only be set once 16 inputs went in
#pragma unroll 1
    for (int jj = 0; jj < 8; jj++) {
#pragma unroll
        for (int i = 0; i < kLeft; i++) {
            uint32_t acc = 0u, acc2 = 0u;             // byte t = count of column jj + 8t: low 8 bits / bits 8..
#pragma unroll
            for (int p = 0; p < 4; p++) acc += ((A[i][p] >> jj) & 0x01010101u) << p;
            if (high) {
#pragma unroll
                for (int p = 4; p < kNB; p++) acc += ((A[i][p] >> jj) & 0x01010101u) << p;
#pragma unroll
                for (int p = 0; p < R; p++) acc2 += ((A[i][kNB + p] >> jj) & 0x01010101u) << p;
            }
            const uint32_t ev = (acc & 0x00FF00FFu) + ((acc2 & 0x00FF00FFu) << 8);               // columns jj, jj+16
            const uint32_t od = ((acc >> 8) & 0x00FF00FFu) + (((acc2 >> 8) & 0x00FF00FFu) << 8);  // columns jj+8, jj+24
            const int ci = first + i, w = ci / kNC, k = ci % kNC;
            uint16_t *dst = frow + k * kCols + (kW * wl + w) * (int)kFlushStride + jj;
            dst[0] = (uint16_t)ev;
            dst[16] = (uint16_t)(ev >> 16);
            dst[8] = (uint16_t)od;
            dst[24] = (uint16_t)(od >> 16);
        }
    }
    __syncwarp();
    uint32_t *const pA = counts + lane, *const pC = pA + stride, *const pG = pC + stride, *const pT = pG + stride;
#pragma unroll
    for (int w = 0; w < kW * G; w++) {
        const int at = (int)kFlushStride * w + lane;
        const uint32_t nlo = frow[at], nhi = frow[kCols + at], nb = frow[2 * kCols + at], nv = frow[3 * kCols + at];
        red_add(pA + 32 * w, nv + nb - nlo - nhi);           // RED.ADD, 128 B per warp instruction
        red_add(pC + 32 * w, nlo - nb);
        red_add(pG + 32 * w, nhi - nb);
        red_add(pT + 32 * w, nb);
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
#pragma unroll
            for (int p = 0; p < kNR; p++) pl[w][k][p] = 0u;
            pa[w][k] = 0u;
            pb[w][k] = 0u;
        }
    }
}

// ---- shared-memory accessors on 32-bit shared addresses.  The warp's base address is made
//      opaque to the compiler once (so it lives in a register instead of being recomputed from
//      special registers in every loop) and all hot accesses are base + immediate.
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint2 v)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t opaque(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void mbar_init_s(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Masked words of one ring entry (a piece) for a lane's two window words.  e.x / e.y = first / end column of the
// piece relative to the window, e.z = bit index of window column 0 in the staged data (only its low 5 bits, the
// funnel-shift amount, are used), e.w = shared address of the plane word that holds window column 0.
template <bool HAS_OK>
__device__ __forceinline__ void piece_words(const uint4 e, uint32_t (&x)[kW][kNC], int L0, uint32_t lutb,
                                            uint32_t lane_seq_off, uint32_t seqb, uint32_t okb)
{
    const int a_c = __viaddmin_s32_relu((int)e.x, -L0, 64);      // clamp(first - L0, 0, 64), one VIADDMNMX
    const int e_c = __viaddmin_s32_relu((int)e.y, -L0, 64);
    const uint2 ga = lds64(lutb + 8u * (uint32_t)a_c), ge = lds64(lutb + 8u * (uint32_t)e_c);
    uint32_t m[kW] = {ga.x & ~ge.x, ga.y & ~ge.y};
    const uint32_t wa = e.w + lane_seq_off;
    const uint2 r0 = lds64(wa), r1 = lds64(wa + 8u), r2 = lds64(wa + 16u);
    const uint32_t lo[kW] = {__funnelshift_r(r0.x, r1.x, e.z), __funnelshift_r(r1.x, r2.x, e.z)};
    const uint32_t hi[kW] = {__funnelshift_r(r0.y, r1.y, e.z), __funnelshift_r(r1.y, r2.y, e.z)};
    if (HAS_OK) {
        const uint32_t oa = okb + (uint32_t)((int)(wa - seqb) >> 1);
        const uint32_t o0 = lds32(oa), o1 = lds32(oa + 4u), o2 = lds32(oa + 8u);
        m[0] &= __funnelshift_r(o0, o1, e.z);
        m[1] &= __funnelshift_r(o1, o2, e.z);
    }
#pragma unroll
    for (int w = 0; w < kW; w++) {
        x[w][0] = lo[w] & m[w];
        x[w][1] = hi[w] & m[w];
        x[w][2] = lo[w] & hi[w] & m[w];
        x[w][3] = m[w];
    }
}

// One trip of the vertical counters: four masked inputs per counter go through carry-save levels 0 and 1 every
// trip, level 2 every 2nd, level 3 every 4th and level 4 (with planes 5..7 in shared memory) every 8th trip.
// cnt = pieces per slot counted since the last flush (a multiple of 4, < kCntMax); the caller adds 4 afterwards.
__device__ __forceinline__ void csa_trip(uint32_t (&x)[4][kW][kNC], uint32_t (&pl)[kW][kNC][kNR], uint32_t (&pa)[kW][kNC],
                                         uint32_t (&pb)[kW][kNC], uint32_t cnt, uint32_t spb)
{
    uint32_t c2[kW][kNC];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
            const uint32_t c1a = maj3(pl[w][k][0], x[0][w][k], x[1][w][k]);
            const uint32_t t = pl[w][k][0] ^ x[0][w][k] ^ x[1][w][k];
            const uint32_t c1b = maj3(t, x[2][w][k], x[3][w][k]);
            pl[w][k][0] = t ^ x[2][w][k] ^ x[3][w][k];
            c2[w][k] = maj3(pl[w][k][1], c1a, c1b);
            pl[w][k][1] ^= c1a ^ c1b;
        }
    }
    if (!(cnt & 4u)) {
#pragma unroll
        for (int w = 0; w < kW; w++)
#pragma unroll
            for (int k = 0; k < kNC; k++) pa[w][k] = c2[w][k];
    } else {
        uint32_t c3[kW][kNC];
#pragma unroll
        for (int w = 0; w < kW; w++) {
#pragma unroll
            for (int k = 0; k < kNC; k++) {
                c3[w][k] = maj3(pl[w][k][2], pa[w][k], c2[w][k]);
                pl[w][k][2] ^= pa[w][k] ^ c2[w][k];
            }
        }
        if (!(cnt & 8u)) {
#pragma unroll
            for (int w = 0; w < kW; w++)
#pragma unroll
                for (int k = 0; k < kNC; k++) pb[w][k] = c3[w][k];
        } else {
            uint32_t c4[kW][kNC];
#pragma unroll
            for (int w = 0; w < kW; w++) {
#pragma unroll
                for (int k = 0; k < kNC; k++) {
                    c4[w][k] = maj3(pl[w][k][3], pb[w][k], c3[w][k]);
                    pl[w][k][3] ^= pb[w][k] ^ c3[w][k];
                }
            }
            if (!(cnt & 16u)) {                           // every 8th trip: park the weight-16 carry
#pragma unroll
                for (int w = 0; w < kW; w++)
#pragma unroll
                    for (int k = 0; k < kNC; k++) sts32(spb + 128u * (uint32_t)((w * kNC + k) * 4), c4[w][k]);
            } else {                                      // every 8th trip: planes 5..7 in shared memory
                const bool have = cnt >= 32u;             // (they were never written before trip 8)
#pragma unroll
                for (int w = 0; w < kW; w++) {
#pragma unroll
                    for (int k = 0; k < kNC; k++) {
                        const uint32_t qa = spb + 128u * (uint32_t)((w * kNC + k) * 4);
                        const uint32_t pcv = lds32(qa);
                        uint32_t c = maj3(pl[w][k][4], pcv, c4[w][k]);
                        pl[w][k][4] ^= pcv ^ c4[w][k];
#pragma unroll
                        for (int p = 1; p < 4; p++) {     // ripple the weight-32 carry upwards
                            const uint32_t v = have ? lds32(qa + 128u * p) : 0u;
                            sts32(qa + 128u * p, v ^ c);
                            c &= v;
                        }
                    }
                }
            }
        }
    }
}

// The counting kernel.  One warp = one chunk of consecutive reads, processed in blocks of
// `rpb` <= 31 reads, one read per lane.
template <int G, bool HAS_OK>
#ifdef BC_K1_MAXNREG
__global__ void __maxnreg__(BC_K1_MAXNREG)
#else
__global__ void __launch_bounds__(kK1Threads, kK1MinCtas)
#endif
k1_count_tiled(BatchView bv, CountView cv, const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t rpb,
               const uint32_t *__restrict__ n_chunks_dev)
{
    using C = K1Cfg<G, HAS_OK>;
    constexpr int S = C::S, Q = C::Q;
    constexpr uint32_t kWin = C::kWin, kRing = C::kRing;
    extern __shared__ __align__(128) unsigned char k1_smem[];

    uint32_t lane_u;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
    const int lane = (int)lane_u;
    const int warp_in_cta = threadIdx.x >> 5;

    // lut[v] = the 64 window columns of a lane at or above column v (v in [0, 64])
    uint2 *lut = reinterpret_cast<uint2 *>(k1_smem);
    for (int v = threadIdx.x; v <= 64; v += kK1Threads)
        lut[v] = make_uint2(v < 32 ? 0xFFFFFFFFu << v : 0u, v <= 32 ? 0xFFFFFFFFu : (v < 64 ? 0xFFFFFFFFu << (v - 32) : 0u));
    __syncthreads();                                // the only CTA-wide barrier: warps are independent from here on

    // n_chunks_dev != nullptr: the chunk list was written by an earlier kernel (k1_count_fast defers the blocks its
    // straight-line decode does not take) and its length lives in device memory; the grid is then a fixed number of
    // warps that stride over the list.  Otherwise one chunk per warp.
    const uint32_t warp_id = blockIdx.x * kK1WarpsPerCta + warp_in_cta;
    if (n_chunks_dev) n_chunks = __ldcg(n_chunks_dev);
    else if (warp_id >= n_chunks) return;
    const uint32_t warp_stride = gridDim.x * kK1WarpsPerCta;

    unsigned char *wsm = k1_smem + C::lut_bytes + (size_t)warp_in_cta * C::warp_bytes;
    uint16_t *frow = reinterpret_cast<uint16_t *>(wsm + C::frow_off);          // flush only
    const uint32_t lutb = opaque(smem_u32(k1_smem));
    const uint32_t wb = opaque(smem_u32(wsm));                                 // the warp's region
    const uint32_t ringb = wb + C::ring_off, seqb = wb + C::seq_off, okb = wb + C::ok_off, cigb = wb + C::cig_off,
                   barb = wb + C::bar_off, rngb = wb + C::rng_off;

    const int slot = lane / G, wl = lane % G;
    const int L0 = 32 * kW * wl;                    // window column of this lane's bit 0
    const uint32_t lt_mask = opaque((1u << lane) - 1u);
    const uint32_t spb = opaque(wb + C::frow_off + 4u * (uint32_t)lane);       // this lane's spill words, 128 B apart
    const uint32_t trip_ringb = opaque(ringb + 16u * (uint32_t)slot);          // ring entry of this slot in a trip
    const uint32_t lane_seq_off = 8u * kW * (uint32_t)wl;                      // byte offset of this lane's window words

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init_s(barb + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phases = 0;                            // bit s: parity to wait for on stage s (persists across chunks)

    for (uint32_t chunk_id = warp_id; chunk_id < n_chunks; chunk_id += warp_stride) {
    const Chunk ch = chunks[chunk_id];
    const uint32_t ref_len = ch.ref_len;
    const uint32_t rb = ch.read_begin, re = ch.read_end;
    if (re <= rb) continue;
    const uint32_t nblk = (re - rb + rpb - 1) / rpb;
    uint32_t *const plane0 = cv.counts + ch.col_base;                       // plane A, column 0 of this slot
    uint32_t *const ds_plane = plane0 + (uint64_t)kPlaneDS * cv.stride;

#if BC_K1_META_SMEM
    // ---- block metadata: {cigar_off, seq_woff, start} of block b go to slot b % kStages of a small shared-memory
    //      ring with cp.async (no register is held while the loads are in flight), three blocks ahead; word l + 1
    //      of an offset array is read l's end.  A block's metadata slot is its stage number.
    const uint32_t metab = wb + C::meta_off;
    auto fetch_meta = [&](uint32_t blk, uint32_t slot) {
        const uint32_t a = metab + slot * 384u + 4u * (uint32_t)lane;
        const uint32_t idx = rb + blk * rpb + (uint32_t)lane;         // (the host keeps n_reads < 2^32 - 64)
        if (blk < nblk && idx <= re) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(bv.cigar_off + idx) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a + 128u), "l"(bv.seq_woff + idx) : "memory");
            if (idx < re)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a + 256u), "l"(bv.starts + idx) : "memory");
            else
                sts32(a + 256u, 0u);
        } else {
            sts32(a, 0u);
            sts32(a + 128u, 0u);
            sts32(a + 256u, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto meta_landed = [&]() {                      // every lane's copies are done and visible to the warp
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    };
    auto issue_block = [&](uint32_t blk, uint32_t stg) {
        if (lane == 0) {
            const uint32_t nvalid = min(rpb, re - (rb + blk * rpb));
            const uint32_t ma = metab + stg * 384u;
            const uint32_t c0 = lds32(ma), c1 = lds32(ma + 4u * nvalid);
            const uint32_t s0 = lds32(ma + 128u), s1 = lds32(ma + 128u + 4u * nvalid);
            const uint32_t s_lo = s0 & ~3u, s_n = min(((s1 + 3u) & ~3u) - s_lo, kSeqCap);     // 32 B / 16 B aligned sources
            const uint32_t c_lo = c0 & ~3u, c_n = min(((c1 + 3u) & ~3u) - c_lo, kCigCap);
            sts128(rngb + 16u * stg, make_uint4(s_lo, s_n, c_lo, c_n));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of this stage
            const uint32_t bar = barb + 8u * stg;
            mbar_expect_tx_s(bar, s_n * 8u + (HAS_OK ? s_n * 4u : 0u) + c_n * 4u);
            if (s_n) {
                bulk_g2s_s(seqb + stg * (kSeqCap * 8u), bv.planes + s_lo, s_n * 8u, bar);
                if (HAS_OK) bulk_g2s_s(okb + stg * (kSeqCap * 4u), bv.okmask + s_lo, s_n * 4u, bar);
            }
            if (c_n) bulk_g2s_s(cigb + stg * (kCigCap * 4u), bv.cigar + c_lo, c_n * 4u, bar);
        }
    };
    fetch_meta(0, 0);
    fetch_meta(1, 1);
    fetch_meta(2, 2);
    meta_landed();
    issue_block(0, 0);
    if (nblk > 1) issue_block(1, 1);
    if (nblk > 2) issue_block(2, 2);
#else
    // ---- block metadata: lane l holds read (block_first + l); lane l+1's offsets are read l's ends
    struct Meta { uint32_t start, cbase, wbase; };
    auto load_meta = [&](uint32_t blk) {
        Meta m = {0u, 0u, 0u};
        const uint32_t idx = rb + blk * rpb + (uint32_t)lane;         // (the host keeps n_reads < 2^32 - 64)
        if (blk < nblk && idx <= re) {
            m.cbase = __ldg(bv.cigar_off + idx);
            m.wbase = __ldg(bv.seq_woff + idx);
            if (idx < re) m.start = __ldg(bv.starts + idx);
        }
        return m;
    };
    // Stage block `blk` (metadata m) into pipeline stage stg; what the stage holds
    // {first plane word, plane words, first CIGAR word, CIGAR words} is left in shared memory for the block's turn.
    auto issue_block = [&](uint32_t blk, const Meta &m, uint32_t stg) {
        const uint32_t nvalid = min(rpb, re - (rb + blk * rpb));
        const uint32_t s0 = __shfl_sync(kFull, m.wbase, 0), s1 = __shfl_sync(kFull, m.wbase, (int)nvalid);
        const uint32_t c0 = __shfl_sync(kFull, m.cbase, 0), c1 = __shfl_sync(kFull, m.cbase, (int)nvalid);
        if (lane == 0) {
            const uint32_t s_lo = s0 & ~3u, s_n = min(((s1 + 3u) & ~3u) - s_lo, kSeqCap);     // 32 B / 16 B aligned sources
            const uint32_t c_lo = c0 & ~3u, c_n = min(((c1 + 3u) & ~3u) - c_lo, kCigCap);
            sts128(rngb + 16u * stg, make_uint4(s_lo, s_n, c_lo, c_n));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of this stage
            const uint32_t bar = barb + 8u * stg;
            mbar_expect_tx_s(bar, s_n * 8u + (HAS_OK ? s_n * 4u : 0u) + c_n * 4u);
            if (s_n) {
                bulk_g2s_s(seqb + stg * (kSeqCap * 8u), bv.planes + s_lo, s_n * 8u, bar);
                if (HAS_OK) bulk_g2s_s(okb + stg * (kSeqCap * 4u), bv.okmask + s_lo, s_n * 4u, bar);
            }
            if (c_n) bulk_g2s_s(cigb + stg * (kCigCap * 4u), bv.cigar + c_lo, c_n * 4u, bar);
        }
    };

    Meta m0 = load_meta(0), m1 = load_meta(1), m2 = load_meta(2), m3 = load_meta(3);
    issue_block(0, m0, 0);
    if (nblk > 1) issue_block(1, m1, 1);
    if (nblk > 2) issue_block(2, m2, 2);
#endif
    uint32_t st = 0;                                // stage of the current block

    // ---- warp-uniform state that persists across blocks
    uint32_t ring_head = 0, ring_tail = 0, mark = 0;      // mark: entries before it come from earlier blocks
    uint32_t win_lo = 0, cnt = 0;
    bool win_valid = false;
    uint32_t pl[kW][kNC][kNR], pa[kW][kNC], pb[kW][kNC];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
#pragma unroll
            for (int p = 0; p < kNR; p++) pl[w][k][p] = 0u;
            pa[w][k] = 0u;
            pb[w][k] = 0u;
        }
    }

    // A ring entry: x / y = first / end column of the piece relative to the window, z = bit index
    // of window column 0 in the staged data (only its low 5 bits, the funnel-shift amount, are
    // used), w = shared address of the plane word that holds window column 0.
    auto make_entry = [&](uint32_t rel, uint32_t n, int qbit) {
        const int z = qbit - (int)rel;
        return make_uint4(rel, rel + n, (uint32_t)z, seqb + (uint32_t)((z >> 5) * 8));
    };
    auto piece = [&](const uint4 e, uint32_t (&x)[kW][kNC]) { piece_words<HAS_OK>(e, x, L0, lutb, lane_seq_off, seqb, okb); };

    // j == nblk is a virtual empty block: it drains the ring and does the final flush in the
    // same (single) trip / flush code as everything else.
    for (uint32_t j = 0; j <= nblk; j++) {
        const bool last = (j == nblk);
#if BC_K1_META_SMEM
        // this block's metadata out of its slot, then the slot goes to block j+3
        meta_landed();
        const uint32_t ma = metab + st * 384u + 4u * (uint32_t)lane, ma1 = metab + st * 384u + 4u * (uint32_t)((lane + 1) & 31);
        const uint32_t cbase = lds32(ma), cend_all = lds32(ma1);
        const uint32_t wbase = lds32(ma + 128u), wend = lds32(ma1 + 128u);
        const uint32_t start0 = lds32(ma + 256u);
        __syncwarp();
        fetch_meta(j + 3u, st);
#else
        const Meta m4 = load_meta(j + 4u);                           // rotated in at the end of this block
#endif
        uint32_t nvalid = 0;
        uint4 rg = make_uint4(0u, 0u, 0u, 0u);                       // s_lo, s_n, c_lo, c_n of this stage
        if (!last) {
            mbar_wait_s(barb + 8u * st, (phases >> st) & 1u);
            phases ^= 1u << st;
            nvalid = min(rpb, re - (rb + j * rpb));
            __syncwarp();                                            // lane 0 wrote the range when it issued the block
            rg = lds128(rngb + 16u * st);
        }
        const uint32_t cg = cigb + st * (kCigCap * 4u) - rg.z * 4u;  // shared address of CIGAR word 0
        const int seg_bit0 = (int)(st * kSeqCap * 32u);              // bit index of the stage's first word

        // ---- per-lane read state (count.cpp:35-38)
        const bool valid = (uint32_t)lane < nvalid;
#if !BC_K1_META_SMEM
        const uint32_t cbase = m0.cbase, cend_all = __shfl_down_sync(kFull, m0.cbase, 1);
        const uint32_t wbase = m0.wbase, wend = __shfl_down_sync(kFull, m0.wbase, 1);
        const uint32_t start0 = m0.start;
#endif
        const bool staged = valid && (wend - rg.x) <= rg.y;
        const bool cig_staged = (cend_all - rg.z) <= rg.w;           // this read's CIGAR words are in shared memory
        uint32_t unst = __ballot_sync(kFull, valid && !staged && cend_all > cbase);   // reads to stage by hand
        uint32_t cur = cbase, cend = staged ? cend_all : cbase;
        uint32_t rpos = min(start0, ref_len), rem = 0u, ds_pos = 0u, ds_n = 0u;
        int qb = seg_bit0 + (int)((wbase - rg.x) * 32u);             // bit index of the next read base
        int qend = qb + (int)((wend - wbase) * 32u);                 // end of the staged data of this read
        bool issue_pending = (j >= 1u) && (j + 2u < nblk);           // block j+2 goes into block j-1's stage
        int slow_lane = -1;                                          // lane whose read is staged by hand right now
        uint32_t seg_w = 0u, seg_end = 0u;                           // its current segment / end (plane word indices)

        // ---- fast mode: reads with at most three CIGAR ops (which is nearly all short reads: M, M-I-M,
        //      M-D-M, clipped or =/X spellings) are decoded in straight-line code, no loop and no
        //      votes: a missing op reads as a zero-length M, which changes nothing.  Ops of one match
        //      run (nothing but zero-length or S/H/P ops between them) merge into one piece; a run
        //      starts at the read start (run A) or right after the first non-empty I/D/N (run B).
        //      If every read of the block qualifies, the block runs in fast mode: pieces that fit the
        //      window are pushed with two votes, the rest after a window move.  Anything else (more
        //      ops, clipping at the reference end, long D/N runs, runs longer than a window,
        //      hand-staged reads) takes the lock-step walker, which starts the block from scratch:
        //      the decode has no side effects.
        uint32_t nA = 0u, nB = 0u, ppB = 0u;                          // pending pieces of this lane (length 0 = none)
        int pqB = 0;
        bool fast = false;
        if (!last) {
            const uint32_t ncig = cend_all - cbase;
            bool bad = unst != 0u || (valid && (!cig_staged || ncig > 3u));
            uint32_t sk_pos = 0u, sk_n = 0u;
            if (staged) {
                const uint32_t ca = cg + cbase * 4u;
                uint32_t cw[3];
                cw[0] = ncig > 0u ? lds32(ca) : 0u;
                cw[1] = ncig > 1u ? lds32(ca + 4u) : 0u;
                cw[2] = ncig > 2u ? lds32(ca + 8u) : 0u;
                uint32_t r[4], q[4], mlen[3], dlen[3];
                bool brk[3];
                r[0] = rpos;
                q[0] = (uint32_t)qb;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const uint32_t len = cw[k] >> 4;
                    const uint32_t cls = (kOpClass >> ((cw[k] << 1) & 30u)) & 3u;
                    r[k + 1] = r[k] + ((cls & 1u) ? len : 0u);               // M/=/X (1), D/N (3): count.cpp:67-68, 87
                    q[k + 1] = q[k] + (((cls + 1u) & 2u) ? len : 0u);        // M/=/X (1), I (2): count.cpp:67, 75
                    mlen[k] = cls == 1u ? len : 0u;
                    dlen[k] = cls == 3u ? len : 0u;
                    brk[k] = cls >= 2u && len != 0u;                         // a non-empty I/D/N ends the match run
                }
                // run of op 1: brk0; run of op 2: brk0 + brk1; an M in a third run is left to the walker
                nA = mlen[0] + (brk[0] ? 0u : mlen[1]) + ((brk[0] || brk[1]) ? 0u : mlen[2]);
                nB = (brk[0] ? mlen[1] : 0u) + ((brk[0] != brk[1]) ? mlen[2] : 0u);
                ppB = brk[0] ? r[1] : r[2];
                pqB = (int)(brk[0] ? q[1] : q[2]);
                sk_n = dlen[0] + dlen[1] + dlen[2];
                sk_pos = dlen[0] ? r[0] : (dlen[1] ? r[1] : r[2]);
                const uint32_t nsk = (dlen[0] ? 1u : 0u) + (dlen[1] ? 1u : 0u) + (dlen[2] ? 1u : 0u);
                // reference end, data bounds, piece lengths
                bad = bad || r[3] > ref_len || q[3] > (uint32_t)qend || (brk[0] && brk[1] && mlen[2] != 0u) || nsk > 1u ||
                      sk_n > kLaneSkipMax || nA > C::kMaxFit || nB > C::kMaxFit;
            }
            fast = !__any_sync(kFull, bad);
            if (fast) {
                if (sk_n) {                                          // count.cpp:80-87; short deletions are the rule
                    uint32_t *const dp = ds_plane + sk_pos;
                    red_add(dp, 1u);
                    if (sk_n > 1u) red_add(dp + 1, 1u);
                    if (sk_n > 2u) red_add(dp + 2, 1u);
#pragma unroll 1
                    for (uint32_t t = 3u; t < sk_n; t++) red_add(dp + t, 1u);
                }
                cur = cend;                                          // the walker has nothing to do
            } else {
                nA = 0u;
                nB = 0u;
            }
        }

        for (;;) {
            int action = 2;                 // 0: keep going, 1: everyone waits for a window move, 2: this pass is over
            uint32_t new_lo = 0u;
            if (fast) {
                // ---- push the pending pieces that fit the window: rel + n <= kWin with rel = pp - win_lo
                //      as unsigned (pp below the window wraps)
                const uint32_t relA = rpos - win_lo, relB = ppB - win_lo;
                const bool fitA = nA != 0u && win_valid && relA <= kWin - nA;
                const bool fitB = nB != 0u && win_valid && relB <= kWin - nB;
                const uint32_t mA = __ballot_sync(kFull, fitA), mB = __ballot_sync(kFull, fitB);
                const uint32_t npA = __popc(mA), npB = __popc(mB);
                // the ring always has room for one piece per lane; run B waits a round if both do not fit
                const bool allB = npA + npB <= kRing - (ring_tail - ring_head);
                const uint32_t mBp = allB ? mB : 0u;
                if (mA | mBp) {
                    __syncwarp();                                    // earlier ring reads are done
                    uint32_t at = ring_tail + __popc(mA & lt_mask) + __popc(mBp & lt_mask);
                    if (fitA) {
                        sts128(ringb + 16u * (at & (kRing - 1u)), make_entry(relA, nA, qb));
                        at++;
                        nA = 0u;
                    }
                    if (fitB && allB) {
                        sts128(ringb + 16u * (at & (kRing - 1u)), make_entry(relB, nB, pqB));
                        nB = 0u;
                    }
                    ring_tail += npA + (allB ? npB : 0u);
                    __syncwarp();
                }
                const bool left = (nA | nB) != 0u;
                if (__any_sync(kFull, left)) {
                    if (mB != 0u && !allB) {
                        action = 0;                                  // run B pieces fit, they only waited for ring room
                    } else {                                         // move the window to the lowest pending piece
                        action = 1;
                        new_lo = __reduce_min_sync(kFull, min(nA ? rpos : 0xFFFFFFFFu, nB ? ppB : 0xFFFFFFFFu)) & ~31u;
                    }
                }
            } else {
                // ---- F: fetch CIGAR ops until an M/=/X run is open (count.cpp:40-96)
                bool moved = false;
                while (rem == 0u && ds_n == 0u && cur < cend) {
                    const uint32_t cw = (cur - rg.z) < rg.w ? lds32(cg + cur * 4u) : __ldg(bv.cigar + cur);
                    cur++;
                    moved = true;
                    const uint32_t len = cw >> 4;
                    const uint32_t cls = (kOpClass >> ((cw << 1) & 30u)) & 3u;
                    if (cls == 1u) {                                     // M / = / X, count.cpp:51
                        const uint32_t lim = ref_len - rpos;
                        if (len > lim) cv.status[kStatMaybeOverflow] = 1u;   // would index past the matrix: exact check later
                        rem = min(len, lim);
                    } else if (cls == 2u) {                              // insertion, count.cpp:74
                        qb = (int)min((uint32_t)qb + len, 1u << 30);
                    } else if (cls == 3u) {                              // deletion / skip, count.cpp:80-87
                        const uint32_t lim = ref_len - rpos, n = min(len, lim);
                        if (len > lim) cv.status[kStatIndexError] = 1u;
                        if (n <= kLaneSkipMax) {
                            for (uint32_t t = 0; t < n; t++) atomicAdd(ds_plane + rpos + t, 1u);
                        } else {
                            ds_pos = rpos;
                            ds_n = n;
                        }
                        rpos += n;
                    }                                                    // S, H, P, B: ignored, count.cpp:92-95
                }
                // ---- D: long D/N runs, all lanes help
                uint32_t dsm = __ballot_sync(kFull, ds_n != 0u);
                while (dsm) {
                    const int src = __ffs((int)dsm) - 1;
                    dsm &= dsm - 1u;
                    const uint32_t p = __shfl_sync(kFull, ds_pos, src), n = __shfl_sync(kFull, ds_n, src);
                    for (uint32_t t = lane; t < n; t += 32u) atomicAdd(ds_plane + p + t, 1u);
                }
                ds_n = 0u;
                // ---- P: the part of the open run that fits the window becomes a piece
                const uint32_t relp = rpos - win_lo;
                uint32_t n1 = 0u;
                if (rem != 0u && win_valid && relp < kWin && qb < qend)
                    n1 = min(min(rem, kWin - relp), (uint32_t)(qend - qb));
                const uint32_t pm = __ballot_sync(kFull, n1 != 0u);
                if (pm) {
                    __syncwarp();                                        // earlier ring reads are done
                    if (n1) {
                        sts128(ringb + 16u * ((ring_tail + __popc(pm & lt_mask)) & (kRing - 1u)), make_entry(relp, n1, qb));
                        rpos += n1;
                        qb += (int)n1;
                        rem -= n1;
                        moved = true;
                    }
                    ring_tail += __popc(pm);
                    __syncwarp();
                }
                if (__any_sync(kFull, cur < cend || rem != 0u)) {
                    action = 0;
                    if (!__any_sync(kFull, moved)) {
                        const bool wst = rem != 0u && qb < qend;         // waits for the window (not for data)
                        if (__any_sync(kFull, wst)) {
                            action = 1;
                            new_lo = __reduce_min_sync(kFull, wst ? rpos : 0xFFFFFFFFu) & ~31u;
                        } else {
                            action = 2;
                        }
                    }
                }
            }
            // ---- the one trip site and the one flush site
            for (;;) {
                const uint32_t avail = ring_tail - ring_head;
                if (avail < (uint32_t)Q || cnt == kCntMax) {         // rare: everything but a plain trip
                    bool want_flush = true;                          // a full trip is waiting but the counters are full
                    if (avail < (uint32_t)Q) {
                        const bool drain = action == 1 || (action == 2 && (last || unst != 0u || slow_lane >= 0 ||
                                                                           (issue_pending && (int)(ring_head - mark) < 0)));
                        if (avail != 0u && drain) {                  // pad the ring with empty pieces to a full trip
                            if ((uint32_t)lane < (uint32_t)Q - avail)
                                sts128(ringb + 16u * ((ring_tail + lane) & (kRing - 1u)), make_uint4(0u, 0u, 0u, seqb));
                            ring_tail += (uint32_t)Q - avail;
                            __syncwarp();
                            continue;
                        }
                        want_flush = cnt != 0u && (action == 1 || (action == 2 && last));
                    }
                    if (!want_flush) break;
                    flush_counters<G>(pl, pa, pb, cnt, frow, plane0 + win_lo, cv.stride, lane);
                    cnt = 0u;
                    continue;
                }
                // -- trip: four pieces per read slot, straight-line.  ring_head is a multiple of Q
                //    and Q divides kRing, so the Q entries of a trip never wrap.
                const uint32_t ea = trip_ringb + 16u * (ring_head & (kRing - 1u));
                uint4 e[4];
#pragma unroll
                for (int q = 0; q < 4; q++) e[q] = lds128(ea + 16u * (uint32_t)(q * S));
                ring_head += (uint32_t)Q;
                uint32_t x[4][kW][kNC];
#pragma unroll
                for (int q = 0; q < 4; q++) piece(e[q], x[q]);
                csa_trip(x, pl, pa, pb, cnt, spb);
                cnt += 4u;
            }
            if (issue_pending && (int)(ring_head - mark) >= 0) {     // block j-1's pieces are all counted
#if BC_K1_META_SMEM
                issue_block(j + 2u, st == 0u ? 2u : st - 1u);        // (its metadata landed before this block started)
#else
                issue_block(j + 2u, m2, st == 0u ? 2u : st - 1u);
#endif
                issue_pending = false;
            }
            if (action == 1) {
                win_lo = new_lo;
                win_valid = true;
            }
            if (action != 2) continue;
            if (fast) break;

            // ---- pass over: reads that were not staged are copied into this stage segment by segment
            if (slow_lane >= 0) {
                const int done = __shfl_sync(kFull, (int)(cur >= cend && rem == 0u), slow_lane);
                const int qb_u = __shfl_sync(kFull, qb, slow_lane);
                const uint32_t adv = (uint32_t)(qb_u - seg_bit0) >> 5;      // whole words already consumed
                if (done || adv == 0u || adv >= seg_end - seg_w) {          // finished (or the CIGAR overruns the read)
                    if (lane == slow_lane) {
                        cur = cend;
                        rem = 0u;
                    }
                    slow_lane = -1;
                } else {
                    seg_w += adv;
                    if (lane == slow_lane) qb -= (int)(adv * 32u);
                }
            }
            if (slow_lane < 0) {
                if (unst == 0u) break;
                slow_lane = __ffs((int)unst) - 1;
                unst &= unst - 1u;
                seg_w = __shfl_sync(kFull, wbase, slow_lane);
                seg_end = __shfl_sync(kFull, wend, slow_lane);
                if (lane == slow_lane) {
                    cur = cbase;
                    cend = cend_all;
                    rpos = min(start0, ref_len);
                    rem = 0u;
                    qb = seg_bit0;
                }
            }
            {
                const uint32_t nw = min(seg_end - seg_w, kSeqCap);
                for (uint32_t i = lane; i < nw; i += 32u) {
                    sts64(seqb + (st * kSeqCap + i) * 8u, __ldg(bv.planes + seg_w + i));
                    if (HAS_OK) sts32(okb + (st * kSeqCap + i) * 4u, __ldg(bv.okmask + seg_w + i));
                }
                if (lane == slow_lane) qend = seg_bit0 + (int)(nw * 32u);
                __syncwarp();
            }
        }

        // ---- rotate the pipelines
        mark = ring_tail;
        st = (st == (uint32_t)kStages - 1u) ? 0u : st + 1u;
#if !BC_K1_META_SMEM
        m0 = m1;
        m1 = m2;
        m2 = m3;
        m3 = m4;
#endif
    }
    __syncwarp();
    }   // chunk loop
    if (n_chunks_dev && lane == 0) {
        // the last warp to get here empties the list for the next launch (everything is stream-ordered)
        uint32_t *ctrl = const_cast<uint32_t *>(n_chunks_dev);
        __threadfence();
        if (atomicAdd(ctrl + 1, 1u) == warp_stride - 1u) {
            ctrl[2] = n_chunks;                     // what this launch processed: the host learns whether a batch needs the walker
            ctrl[0] = 0u;
            ctrl[1] = 0u;
        }
    }
}

// Cross-check variant: one thread per read, one RED per base.  Same inputs, same planes.
__global__ void k1_count_per_base(BatchView bv, CountView cv)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bv.n_reads) return;
    const uint32_t r = slot_of_read(bv.ref_read_off, bv.n_refs, i);
    const uint32_t ref_len = cv.ref_len[r];
    const uint64_t base = cv.col_base[r];
    uint32_t ref_pos = bv.starts[i], read_pos = 0;
    const uint32_t wbase = bv.seq_woff[i], nwords = bv.seq_woff[i + 1] - wbase;
    for (uint32_t c = bv.cigar_off[i]; c < bv.cigar_off[i + 1]; c++) {
        const uint32_t op = bv.cigar[c] & 0xFu, len = bv.cigar[c] >> 4;
        if (op_is_match(op)) {
            for (uint32_t j = 0; j < len; j++) {
                const uint32_t rp = read_pos + j, col = ref_pos + j;
                if ((rp >> 5) >= nwords) break;
                const uint2 w = bv.planes[wbase + (rp >> 5)];
                const uint32_t code = ((w.x >> (rp & 31)) & 1u) | (((w.y >> (rp & 31)) & 1u) << 1);
                const bool ok = bv.okmask ? ((bv.okmask[wbase + (rp >> 5)] >> (rp & 31)) & 1u) : true;
                if (!ok) continue;
                if (col >= ref_len || col < ref_pos) { cv.status[kStatMaybeOverflow] = 1u; continue; }
                atomicAdd(cv.counts + (uint64_t)code * cv.stride + base + col, 1u);
            }
            ref_pos = sat_add(ref_pos, len);
            read_pos = sat_add(read_pos, len);
        } else if (op == 1u) {
            read_pos = sat_add(read_pos, len);
        } else if (op_is_refskip(op)) {
            for (uint32_t j = 0; j < len; j++) {
                const uint32_t col = ref_pos + j;
                if (col >= ref_len || col < ref_pos) { cv.status[kStatIndexError] = 1u; break; }
                atomicAdd(cv.counts + (uint64_t)kPlaneDS * cv.stride + base + col, 1u);
            }
            ref_pos = sat_add(ref_pos, len);
        }
    }
}

// Map a read position to its reference column by walking the read's CIGAR.
// Returns false if the position lies in an insertion (or past the alignment).
__device__ __forceinline__ bool read_pos_to_col(const BatchView &bv, uint32_t i, uint32_t pos, uint32_t *col)
{
    uint32_t ref_pos = bv.starts[i], read_pos = 0;
    for (uint32_t c = bv.cigar_off[i]; c < bv.cigar_off[i + 1]; c++) {
        const uint32_t op = bv.cigar[c] & 0xFu, len = bv.cigar[c] >> 4;
        if (op_is_match(op)) {
            if (pos >= read_pos && pos - read_pos < len) {
                const uint32_t t = ref_pos + (pos - read_pos);
                if (t < ref_pos) return false;      // wrapped: far past any reference
                *col = t;
                return true;
            }
            ref_pos = sat_add(ref_pos, len);
            read_pos = sat_add(read_pos, len);
        } else if (op == 1u) {
            if (pos >= read_pos && pos - read_pos < len) return false;
            read_pos = sat_add(read_pos, len);
        } else if (op_is_refskip(op)) {
            ref_pos = sat_add(ref_pos, len);
        }
    }
    return false;
}

// Sparse corrections for letters outside ACGT (see bc_batch.exc_* in the header).  One thread per
// exception; the chain of dependent loads is what it costs, so the slot tables (a binary search plus two
// look-ups per thread) are staged in shared memory when they fit.
constexpr uint32_t kExcSlotsInSmem = 256;
__global__ void __launch_bounds__(128)
k1_exceptions(BatchView bv, CountView cv)
{
    __shared__ uint32_t s_off[kExcSlotsInSmem + 1], s_base[kExcSlotsInSmem], s_len[kExcSlotsInSmem];
    const bool staged = bv.n_refs <= kExcSlotsInSmem;
    if (staged) {
        for (uint32_t t = threadIdx.x; t <= bv.n_refs; t += blockDim.x) s_off[t] = bv.ref_read_off[t];
        for (uint32_t t = threadIdx.x; t < bv.n_refs; t += blockDim.x) {
            s_base[t] = cv.col_base[t];
            s_len[t] = cv.ref_len[t];
        }
        __syncthreads();
    }
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= bv.n_exc) return;
    const uint32_t i = bv.exc_read[e];
    const uint32_t ep = bv.exc_pos[e];
    const uint32_t pos = ep >> 2, flags = ep & 3u;
    if (i >= bv.n_reads) return;
    uint32_t col;
    if (!read_pos_to_col(bv, i, pos, &col)) return;
    const uint32_t r = slot_of_read(staged ? s_off : bv.ref_read_off, bv.n_refs, i);
    const uint64_t base = staged ? s_base[r] : cv.col_base[r];
    if (col >= (staged ? s_len[r] : cv.ref_len[r])) {
        if (flags & 1u) cv.status[kStatIndexError] = 1u;   // an N that counts, past the end (count.cpp:64)
        return;                                            // flag 2: the main pass clipped it already
    }
    if (flags & 2u) atomicAdd(cv.counts + base + col, 0xFFFFFFFFu);                              // undo the 'A'
    if (flags & 1u) atomicAdd(cv.counts + (uint64_t)kPlaneN * cv.stride + base + col, 1u);     // count.cpp:64
}

__device__ __forceinline__ uint32_t find_exception(const BatchView &bv, uint32_t i, uint32_t pos)
{
    // exceptions are sorted by (read, pos); returns flags or 0xFFFFFFFF if absent
    uint32_t lo = 0, hi = bv.n_exc;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t r = bv.exc_read[mid], p = bv.exc_pos[mid] >> 2;
        if (r < i || (r == i && p < pos)) lo = mid + 1; else hi = mid;
    }
    if (lo < bv.n_exc && bv.exc_read[lo] == i && (bv.exc_pos[lo] >> 2) == pos) return bv.exc_pos[lo] & 3u;
    return 0xFFFFFFFFu;
}

// Runs only when a piece crossed ref_len: decides, base by base, whether the reference
// would have thrown (an INCREMENT at refPos >= refLen, count.cpp:60-64) -- bases that fail
// the quality test or are not in ACGTN never index the matrix and never throw.
__global__ void k1_check_overflow(BatchView bv, CountView cv)
{
    if (cv.status[kStatMaybeOverflow] == 0u) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < bv.n_reads; i += gridDim.x * blockDim.x) {
    const uint32_t r = slot_of_read(bv.ref_read_off, bv.n_refs, i);
    const uint32_t ref_len = cv.ref_len[r];
    uint32_t ref_pos = bv.starts[i], read_pos = 0;
    const uint32_t wbase = bv.seq_woff[i], nwords = bv.seq_woff[i + 1] - wbase;
    for (uint32_t c = bv.cigar_off[i]; c < bv.cigar_off[i + 1]; c++) {
        const uint32_t op = bv.cigar[c] & 0xFu, len = bv.cigar[c] >> 4;
        if (op_is_match(op)) {
            if (len && (ref_pos >= ref_len || len > ref_len - ref_pos)) {
                const uint32_t first = ref_pos >= ref_len ? 0u : ref_len - ref_pos;
                for (uint32_t j = first; j < len; j++) {
                    const uint32_t rp = read_pos + j;
                    if (rp < read_pos || (rp >> 5) >= nwords) break;
                    bool counted;
                    const uint32_t ex = find_exception(bv, i, rp);
                    if (bv.okmask) {
                        counted = ((bv.okmask[wbase + (rp >> 5)] >> (rp & 31)) & 1u) || (ex != 0xFFFFFFFFu && (ex & 1u));
                    } else {
                        counted = (ex == 0xFFFFFFFFu) || (ex & 1u) || !(ex & 2u);
                    }
                    if (counted) { cv.status[kStatIndexError] = 1u; return; }
                }
            }
            ref_pos = sat_add(ref_pos, len);
            read_pos = sat_add(read_pos, len);
        } else if (op == 1u) {
            read_pos = sat_add(read_pos, len);
        } else if (op_is_refskip(op)) {
            ref_pos = sat_add(ref_pos, len);
        }
    }
    }
}

// uint32 per-batch counters -> int64 totals (only needed past 2^32 reads; main.py:132,155)
__global__ void k_fold_counts(uint32_t *__restrict__ c32, unsigned long long *__restrict__ c64, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    c64[i] += c32[i];
    c32[i] = 0u;
}

// planes -> refLen x 6 int64 row-major (the layout the reference hands to get_stats)
__global__ void k_export_counts(const uint32_t *__restrict__ c32, const unsigned long long *__restrict__ c64,
                                uint64_t stride, uint64_t col_base, uint32_t ref_len, long long *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)ref_len * kPlanes) return;
    const uint32_t pos = (uint32_t)(t / kPlanes), p = (uint32_t)(t % kPlanes);
    const uint64_t src = (uint64_t)p * stride + col_base + pos;
    out[t] = (long long)c32[src] + (c64 ? (long long)c64[src] : 0ll);
}

}  // namespace bc
