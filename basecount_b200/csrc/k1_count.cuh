// k1_count.cuh -- K1: CIGAR walk + per-position base counting on sm_100a.
//
// Replaces the loop nest of the reference operator (basecount/count.cpp:22-97).
//
// Design (B200-first, nothing here mirrors the reference's scalar loop):
//   * A warp owns a chunk of consecutive reads of one reference slot.  Its 32 lanes
//     form S = 32/G "read slots" of G lanes; lane (slot, w) covers the 32 reference
//     columns of window word w.  All slots share ONE window of 32*G columns, so window
//     moves and flushes are warp-uniform (no divergence between slots).
//   * Sequence data is 2-bit, bit-planar (lo/hi plane): one 64-bit load + a funnel
//     shift aligns 32 bases of a read onto a 32-column window word; three LOP3 build
//     the masked one-hot words for A/C/G/T.  No per-base work anywhere.
//   * Counts are accumulated VERTICALLY in registers as bit-sliced counters
//     (carry-save adders, Harley-Seal style): ~3 LOP3 per 32 bases per base letter.
//   * A flush converts the bit-sliced counters to integers (byte-packed extraction,
//     PRMT transposes, shared-memory staging) and adds them to the HBM count planes
//     with coalesced 128-byte RED.ADD -- one atomic per column per flush instead of
//     one per base.
//   * D / N ops add to the DS plane directly (rare); letters the 2-bit code cannot
//     express (N, IUPAC, lower case) are a sparse correction pass (k1_exceptions).
#pragma once
#include "bc_common.cuh"

namespace bc {

constexpr int kK1WarpsPerCta = 4;
constexpr int kK1MinCtas = 3;
constexpr int kK1Threads = kK1WarpsPerCta * 32;
constexpr int kNB = 8;                       // bit planes per vertical counter (counts to 255)
constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b)
{
    uint32_t t = a + b;
    return t < a ? 0xFFFFFFFFu : t;
}

// Vertical counters of one lane: for each of A,C,G,T, kNB bit planes over the lane's 32
// columns (Harley-Seal style carry-save adders).  Inputs arrive two at a time, so the
// weight-1 adder consumes both directly; pend[b][0] / pend[b][1] hold the not-yet-paired
// carries of weight 2 and 4.
struct VCounters {
    uint32_t pl[4][kNB];
    uint32_t pend[4][2];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int b = 0; b < 4; b++) {
#pragma unroll
            for (int k = 0; k < kNB; k++) pl[b][k] = 0;
            pend[b][0] = 0;
            pend[b][1] = 0;
        }
    }
    // Add two 32-column masks per letter.  `cnt` = inputs added so far (even, warp-uniform),
    // so every branch below is uniform.
    __device__ __forceinline__ void add2(const uint32_t xa[4], const uint32_t xb[4], uint32_t cnt)
    {
        uint32_t c1[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            c1[b] = maj3(pl[b][0], xa[b], xb[b]);
            pl[b][0] ^= xa[b] ^ xb[b];
        }
        if ((cnt & 2u) == 0u) {
#pragma unroll
            for (int b = 0; b < 4; b++) pend[b][0] = c1[b];
            return;
        }
        uint32_t c2[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            c2[b] = maj3(pl[b][1], pend[b][0], c1[b]);
            pl[b][1] ^= pend[b][0] ^ c1[b];
        }
        if ((cnt & 4u) == 0u) {
#pragma unroll
            for (int b = 0; b < 4; b++) pend[b][1] = c2[b];
            return;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t c = maj3(pl[b][2], pend[b][1], c2[b]);
            pl[b][2] ^= pend[b][1] ^ c2[b];
#pragma unroll
            for (int k = 3; k < kNB; k++) {          // ripple the weight-8 carry upwards
                uint32_t t = pl[b][k] & c;
                pl[b][k] ^= c;
                c = t;
            }
        }
    }
};

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

// ---- geometry -----------------------------------------------------------------------------
// A warp's 32 lanes form S = 32/G read slots of G lanes; every lane covers W consecutive
// 32-column window words, so the (warp-shared) window spans 32*W*G reference columns.
constexpr int kW = 2;
constexpr uint32_t kSeqCap = 512;     // staged 64-bit plane words per stage (4 KB): 32 reads x 400 bp = 416 words
constexpr uint32_t kSeqPad = 4;       // guard words so clamped out-of-piece loads stay inside the stage
constexpr uint32_t kCigCap = 256;     // staged CIGAR words per stage (1 KB)
constexpr int kStages = 2;
constexpr uint32_t kCntMax = 254u;    // per-slot count limit of the 8-plane counters (inputs come in pairs)
constexpr uint32_t kFlushStride = 34; // uint16 per staged window word (32 + 2 pad: conflict-free stores)
constexpr uint32_t kRing = 128;       // piece ring entries (uint4 each)
constexpr uint32_t kLaneSkipMax = 256; // longest D/N run a single lane adds itself

template <int G, bool HAS_OK>
__host__ __device__ constexpr uint32_t k1_warp_smem_bytes()
{
    return kStages * (kSeqCap + kSeqPad) * 8u + (HAS_OK ? kStages * (kSeqCap + kSeqPad) * 4u : 0u) +
           kStages * kCigCap * 4u + /* flush rows: 4 letters x (kW*G words x 34) u16 */ 4u * kW * G * kFlushStride * 2u +
           /* piece ring */ kRing * 16u + /* order list */ 64u + /* mbarriers */ 64u;
}

struct BlockMeta {            // lane l holds the metadata of read (block_first + l); raw loaded values only,
    uint32_t start, cbase, cend, wbase, wend;   // so nothing waits on the loads until the block is used
    __device__ __forceinline__ uint32_t ncig() const { return cend - cbase; }
    __device__ __forceinline__ uint32_t nwords() const { return wend - wbase; }
};
struct StagedRange {          // what one pipeline stage holds (warp-uniform)
    uint32_t s_lo, s_n;       // plane / okmask words [s_lo, s_lo + s_n)
    uint32_t c_lo, c_n;       // CIGAR words        [c_lo, c_lo + c_n)
};

// Convert the warp's vertical counters to integers and add them to the HBM planes.
// Byte-packed extraction per slot, widened to 16 bit before the S read slots are summed
// (so each slot may hold up to 255), staged in shared memory and written with coalesced
// RED.ADD (128 B per warp instruction).  The loop over the shift amount jj is deliberately
// NOT unrolled: registers stay statically indexed while the code stays a few hundred
// instructions (a fully unrolled flush inlined at every site made the kernel > 500 KB and
// instruction-fetch bound).
//   frow: 4 letters x (kW*G window words x kFlushStride) uint16
template <int G>
__device__ __forceinline__ void flush_counters(VCounters (&vc)[kW], uint32_t cnt, uint16_t *frow, uint64_t win_col,
                                               uint32_t *__restrict__ counts, uint64_t stride, int lane)
{
    constexpr int S = 32 / G;
    constexpr int kCols = (int)kFlushStride * kW * G;
    const int slot = lane / G, wl = lane % G;
    // pendings that hold no carry are zeroed so the loop below needs no cnt tests
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (!(cnt & 2u)) vc[w].pend[b][0] = 0u;
            if (!(cnt & 4u)) vc[w].pend[b][1] = 0u;
        }
    }
    const bool high = cnt >= 16u;                         // planes 4..7 can only be set after 16 inputs
#pragma unroll 1
    for (int jj = 0; jj < 8; jj++) {
        const bool mine = (jj % S) == slot;
#pragma unroll
        for (int b = 0; b < 4; b++) {
#pragma unroll
            for (int w = 0; w < kW; w++) {
                uint32_t acc = 0;                         // byte t = count of column jj + 8t (this slot only)
#pragma unroll
                for (int k = 0; k < 4; k++) acc += ((vc[w].pl[b][k] >> jj) & 0x01010101u) << k;
                if (high) {
#pragma unroll
                    for (int k = 4; k < kNB; k++) acc += ((vc[w].pl[b][k] >> jj) & 0x01010101u) << k;
                }
#pragma unroll
                for (int k = 0; k < 2; k++) acc += ((vc[w].pend[b][k] >> jj) & 0x01010101u) << (k + 1);
                uint32_t ev = acc & 0x00FF00FFu;          // columns jj, jj+16
                uint32_t od = (acc >> 8) & 0x00FF00FFu;   // columns jj+8, jj+24
#pragma unroll
                for (int d = G; d < 32; d <<= 1) {        // sum the read slots (<= 8 * 255 fits 16 bit)
                    ev += __shfl_xor_sync(kFull, ev, d);
                    od += __shfl_xor_sync(kFull, od, d);
                }
                if (mine) {
                    uint16_t *dst = frow + b * kCols + (kW * wl + w) * (int)kFlushStride + jj;
                    dst[0] = (uint16_t)ev;
                    dst[16] = (uint16_t)(ev >> 16);
                    dst[8] = (uint16_t)od;
                    dst[24] = (uint16_t)(od >> 16);
                }
            }
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
        uint32_t *plane = counts + (uint64_t)b * stride + win_col;
        const uint16_t *row = frow + b * kCols;
#pragma unroll 4
        for (int w = 0; w < kW * G; w++) {
            const uint32_t val = row[(int)kFlushStride * w + lane];
            if (val) atomicAdd(plane + 32 * w + lane, val);     // RED.ADD, 128 B per warp instruction
        }
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < kW; w++) vc[w].clear();
}

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n)
{
    uint32_t r;                                    // PTX shl clamps shift amounts above 31 (result 0)
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
// bits [a, e) of a 32-bit word; a, e may lie outside 0..32 (empty if e <= a)
__device__ __forceinline__ uint32_t bit_range(int a, int e)
{
    return shl_clamp(0xFFFFFFFFu, (uint32_t)max(a, 0)) & ~shl_clamp(0xFFFFFFFFu, (uint32_t)max(e, 0));
}

// Masked one-hot words of one piece for this lane's kW window words (branch-free).
//   pp / n1 : reference start and length of the (window-clipped) piece; n1 = 0 gives zeros
//   sbit    : bit index, in the stage buffer, of the piece's first base (read fully staged)
template <int G, bool HAS_OK>
__device__ __forceinline__ void piece_words(uint32_t (&x)[kW][4], uint32_t pp, uint32_t n1, uint32_t win_lo, int wl,
                                            const uint2 *__restrict__ sq, const uint32_t *__restrict__ okb, int sbit)
{
    const int rel = (int)(win_lo + 32u * kW * (uint32_t)wl) - (int)pp;   // piece offset of this lane's column 0
    const int bit = sbit + rel;
    const int sh = bit & 31;
    // Out-of-piece words are masked away below, so only memory safety matters for the index.
    const int base = min(max(bit >> 5, 0), (int)(kSeqCap + kSeqPad) - (kW + 1));
    uint2 r[kW + 1];
    uint32_t o[kW + 1];
#pragma unroll
    for (int i = 0; i <= kW; i++) {
        r[i] = sq[base + i];
        if (HAS_OK) o[i] = okb[base + i];
    }
#pragma unroll
    for (int w = 0; w < kW; w++) {
        // columns of word w cover piece offsets rel + 32w .. rel + 32w + 31; valid offsets are [0, n1)
        uint32_t m = bit_range(-rel - 32 * w, (int)n1 - rel - 32 * w);
        if (HAS_OK) m &= __funnelshift_r(o[w], o[w + 1], sh);
        const uint32_t lo = __funnelshift_r(r[w].x, r[w + 1].x, sh);
        const uint32_t hi = __funnelshift_r(r[w].y, r[w + 1].y, sh);
        x[w][0] = ~hi & ~lo & m;                                       // A
        x[w][1] = ~hi & lo & m;                                        // C
        x[w][2] = hi & ~lo & m;                                        // G
        x[w][3] = hi & lo & m;                                         // T
    }
}

// Same, for reads that are not (fully) staged: bounds-checked loads from HBM.
template <int G, bool HAS_OK>
__device__ __forceinline__ void piece_words_global(uint32_t (&x)[kW][4], uint32_t pp, uint32_t n1, uint32_t win_lo, int wl,
                                                   const BatchView &bv, uint32_t pq, uint32_t wbase, uint32_t nwords)
{
    const int rel = (int)(win_lo + 32u * kW * (uint32_t)wl) - (int)pp;
    const int bit = (int)pq + rel;                                     // read bit index of the lane's column 0
    const int k = bit >> 5, sh = bit & 31;
    uint2 r[kW + 1];
    uint32_t o[kW + 1];
#pragma unroll
    for (int i = 0; i <= kW; i++) {
        r[i] = make_uint2(0u, 0u);
        o[i] = 0u;
        if (n1 && k + i >= 0 && (uint32_t)(k + i) < nwords) {
            r[i] = __ldg(bv.planes + wbase + (uint32_t)(k + i));
            if (HAS_OK) o[i] = __ldg(bv.okmask + wbase + (uint32_t)(k + i));
        }
    }
#pragma unroll
    for (int w = 0; w < kW; w++) {
        uint32_t m = bit_range(-rel - 32 * w, (int)n1 - rel - 32 * w);
        if (HAS_OK) m &= __funnelshift_r(o[w], o[w + 1], sh);
        const uint32_t lo = __funnelshift_r(r[w].x, r[w + 1].x, sh);
        const uint32_t hi = __funnelshift_r(r[w].y, r[w + 1].y, sh);
        x[w][0] = ~hi & ~lo & m;
        x[w][1] = ~hi & lo & m;
        x[w][2] = hi & ~lo & m;
        x[w][3] = hi & lo & m;
    }
}

// The counting kernel.  One warp = one chunk of consecutive reads, processed in blocks of
// `rpb` <= 32 reads.  Per block: metadata sits in registers (one read per lane, handed to the
// read slots by shuffle); sequence / CIGAR words were staged in shared memory by TMA bulk
// copies two blocks ahead.  Reads that are a single M/=/X run (the common case) go through a
// branch-free fast loop, S reads per iteration; the rest (indels, clips, long or unstaged
// reads, reads deferred by a window move) go through the general CIGAR state machine.
template <int G, bool HAS_OK>
__global__ void __launch_bounds__(kK1Threads, kK1MinCtas)
k1_count_tiled(BatchView bv, CountView cv, const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t rpb)
{
    constexpr int S = 32 / G;
    constexpr uint32_t kWin = 32u * kW * G;         // window columns
    constexpr uint32_t kMaxFit = kWin - 31u;        // a piece this long fits a fresh window at any alignment
    constexpr uint32_t kStageWords = kSeqCap + kSeqPad;
    extern __shared__ __align__(128) unsigned char k1_smem[];

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const uint32_t warp_id = blockIdx.x * kK1WarpsPerCta + warp_in_cta;
    if (warp_id >= n_chunks) return;                // warps are independent: no CTA-wide barrier anywhere

    unsigned char *wsm = k1_smem + (size_t)warp_in_cta * k1_warp_smem_bytes<G, HAS_OK>();
    uint2 *seq_buf = reinterpret_cast<uint2 *>(wsm);
    uint32_t *ok_buf = reinterpret_cast<uint32_t *>(wsm + kStages * kStageWords * 8u);
    uint32_t *cig_buf = reinterpret_cast<uint32_t *>(wsm + kStages * kStageWords * 8u + (HAS_OK ? kStages * kStageWords * 4u : 0u));
    uint16_t *frow = reinterpret_cast<uint16_t *>(cig_buf + kStages * kCigCap);
    uint4 *ring = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(frow) + 4u * kW * G * kFlushStride * 2u);
    uint8_t *order2 = reinterpret_cast<uint8_t *>(ring + kRing);
    uint64_t *bars = reinterpret_cast<uint64_t *>(order2 + 64);

    const int slot = lane / G, wl = lane % G;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t slot_lead_below = (slot == 0) ? 0u : ((1u << (slot * G)) - 1u);

    const Chunk ch = chunks[warp_id];
    const uint32_t ref_len = ch.ref_len;
    const uint32_t rb = ch.read_begin, re = ch.read_end;
    const uint32_t nblk = (re - rb + rpb - 1) / rpb;
    const uint64_t col0 = ch.col_base;

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto load_meta = [&](uint32_t blk) {
        BlockMeta m = {0u, 0u, 0u, 0u, 0u};
        const uint64_t idx = (uint64_t)rb + (uint64_t)blk * rpb + lane;
        if ((uint32_t)lane < rpb && idx < re) {
            m.start = __ldg(bv.starts + idx);
            m.cbase = __ldg(bv.cigar_off + idx);
            m.cend = __ldg(bv.cigar_off + idx + 1);
            m.wbase = __ldg(bv.seq_woff + idx);
            m.wend = __ldg(bv.seq_woff + idx + 1);
        }
        return m;
    };
    // Stage the words of block `blk` (metadata m) into pipeline stage b.
    auto issue_block = [&](uint32_t blk, const BlockMeta &m, int b) {
        const uint32_t nvalid = min(rpb, re - (rb + blk * rpb));
        const uint32_t s0 = __shfl_sync(kFull, m.wbase, 0);
        const uint32_t s1 = __shfl_sync(kFull, m.wend, (int)nvalid - 1);
        const uint32_t c0 = __shfl_sync(kFull, m.cbase, 0);
        const uint32_t c1 = __shfl_sync(kFull, m.cend, (int)nvalid - 1);
        StagedRange r;
        r.s_lo = s0 & ~3u;                                           // 32 B / 16 B aligned sources
        r.s_n = min(((s1 + 3u) & ~3u) - r.s_lo, kSeqCap);
        r.c_lo = c0 & ~3u;
        r.c_n = min(((c1 + 3u) & ~3u) - r.c_lo, kCigCap);
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this stage
            const uint32_t bytes = r.s_n * 8u + (HAS_OK ? r.s_n * 4u : 0u) + r.c_n * 4u;
            mbar_expect_tx(&bars[b], bytes);
            if (r.s_n) {
                bulk_g2s(seq_buf + b * kStageWords + kSeqPad, bv.planes + r.s_lo, r.s_n * 8u, &bars[b]);
                if (HAS_OK) bulk_g2s(ok_buf + b * kStageWords + kSeqPad, bv.okmask + r.s_lo, r.s_n * 4u, &bars[b]);
            }
            if (r.c_n) bulk_g2s(cig_buf + b * kCigCap, bv.cigar + r.c_lo, r.c_n * 4u, &bars[b]);
        }
        return r;
    };

    BlockMeta m0 = load_meta(0), m1 = load_meta(1), m2 = load_meta(2);
    StagedRange rg = issue_block(0, m0, 0);
    StagedRange rg_nxt = {0u, 0u, 0u, 0u};
    if (nblk > 1) rg_nxt = issue_block(1, m1, 1);
    uint32_t parity = 0;                            // bit b: phase to wait for on stage b

    // warp-uniform window + counter + piece-ring state (persist across blocks)
    uint32_t ring_head = 0, ring_tail = 0;
    uint32_t win_lo = 0, cnt = 0;
    bool win_valid = false;
    VCounters vc[kW];
#pragma unroll
    for (int w = 0; w < kW; w++) vc[w].clear();

    auto do_flush = [&]() {
        flush_counters<G>(vc, cnt, frow, col0 + win_lo, cv.counts, cv.stride, lane);
        cnt = 0;
    };
    auto accumulate2 = [&](uint32_t (&xa)[kW][4], uint32_t (&xb)[kW][4]) {   // callers flush at kCntMax
#pragma unroll
        for (int w = 0; w < kW; w++) vc[w].add2(xa[w], xb[w], cnt);
        cnt += 2;
    };

    for (uint32_t j = 0; j < nblk; j++) {
        const int b = (int)(j & 1u);
        mbar_wait(&bars[b], (parity >> b) & 1u);
        parity ^= 1u << b;
        const uint2 *sq = seq_buf + b * kStageWords;
        const uint32_t *okb = ok_buf + b * kStageWords;
        const uint32_t *cg = cig_buf + b * kCigCap;
        const uint32_t nvalid = min(rpb, re - (rb + j * rpb));

        // ---- per-read classification, one read per lane: walk the read's own CIGAR and emit up
        //      to two M/=/X pieces (adjacent match ops merge) into the piece ring; anything
        //      bigger (more pieces, long pieces or skips, unstaged reads) goes to the general loop.
        const bool valid = (uint32_t)lane < nvalid;
        const bool staged = valid && (m0.wbase - rg.s_lo) <= rg.s_n && (m0.wbase - rg.s_lo) + m0.nwords() <= rg.s_n;
        const int my_sidx = staged ? (int)(m0.wbase - rg.s_lo + kSeqPad) : -1;
        uint32_t pc_pp[2] = {0u, 0u}, pc_pq[2] = {0u, 0u}, pc_pn[2] = {0u, 0u};
        uint32_t n_pc = 0, sk_pos = 0, sk_len = 0;
        bool complex_read = valid && m0.ncig() && !staged;
        if (valid && m0.ncig() && staged) {
            uint32_t rpos = m0.start, qpos = 0;
            bool open = false;                                           // last op was a match op (merge candidates)
            for (uint32_t c = m0.cbase; c < m0.cbase + m0.ncig(); c++) {
                const uint32_t ci = c - rg.c_lo;
                const uint32_t cw = ci < rg.c_n ? cg[ci] : __ldg(bv.cigar + c);
                const uint32_t op = cw & 0xFu, len = cw >> 4;
                if (op_is_match(op)) {
                    if (len) {
                        if (open) {
                            pc_pn[n_pc - 1] = sat_add(pc_pn[n_pc - 1], len);
                        } else if (n_pc < 2) {
                            pc_pp[n_pc] = rpos;
                            pc_pq[n_pc] = qpos;
                            pc_pn[n_pc] = len;
                            n_pc++;
                            open = true;
                        } else {
                            complex_read = true;
                            break;
                        }
                    }
                    rpos = sat_add(rpos, len);
                    qpos = sat_add(qpos, len);
                } else if (op == 1u) {
                    qpos = sat_add(qpos, len);
                    open = open && len == 0u;
                } else if (op_is_refskip(op)) {
                    if (len) {
                        if (sk_len || len > kLaneSkipMax) {
                            complex_read = true;
                            break;
                        }
                        sk_pos = rpos;
                        sk_len = len;
                        open = false;
                    }
                    rpos = sat_add(rpos, len);
                }
            }
            if (pc_pn[0] > kMaxFit || pc_pn[1] > kMaxFit) complex_read = true;
        }
        if (complex_read) n_pc = 0;
        if (valid && !complex_read) {
            // clip to the reference (count.cpp .at()): exactness decided by k1_check_overflow
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (pc_pn[i] && (pc_pp[i] >= ref_len || pc_pn[i] > ref_len - pc_pp[i])) {
                    cv.status[kStatMaybeOverflow] = 1u;
                    pc_pn[i] = pc_pp[i] < ref_len ? ref_len - pc_pp[i] : 0u;
                }
            }
            if (sk_len) {                                                // deletion / skip, count.cpp:80-87
                const uint32_t lim = sk_pos < ref_len ? min(sk_len, ref_len - sk_pos) : 0u;
                if (lim < sk_len) cv.status[kStatIndexError] = 1u;
                uint32_t *ds = cv.counts + (uint64_t)kPlaneDS * cv.stride + col0 + sk_pos;
                for (uint32_t t = 0; t < lim; t++) atomicAdd(ds + t, 1u);
            }
        }
        const uint32_t has1 = __ballot_sync(kFull, n_pc >= 1u && pc_pn[0]);
        const uint32_t has2 = __ballot_sync(kFull, n_pc >= 2u && pc_pn[1]);
        uint32_t rest_mask = __ballot_sync(kFull, complex_read);
        {
            uint32_t at = ring_tail + __popc(has1 & lt_mask) + __popc(has2 & lt_mask);
            if ((has1 >> lane) & 1u) {
                ring[at & (kRing - 1)] = make_uint4(pc_pp[0], pc_pn[0], (uint32_t)(my_sidx * 32) + pc_pq[0], 0u);
                at++;
            }
            if ((has2 >> lane) & 1u)
                ring[at & (kRing - 1)] = make_uint4(pc_pp[1], pc_pn[1], (uint32_t)(my_sidx * 32) + pc_pq[1], 0u);
            ring_tail += __popc(has1) + __popc(has2);
        }
        __syncwarp();

        // ---- fast loop: 2*S pieces per trip straight from the ring (two per read slot, so the
        //      weight-1 adder needs no pending register); straight-line body; rare events
        //      (flush, window move) are handled outside the tight inner loop
        for (;;) {
            uint32_t fit_any = 0u, low_pp = 0xFFFFFFFFu;
            while (ring_head != ring_tail) {
                const uint32_t avail = ring_tail - ring_head;
                const uint32_t adv = min((uint32_t)(2 * S), avail);     // entries consumed this trip
                const bool have_a = (uint32_t)slot < avail, have_b = (uint32_t)(slot + S) < avail;
                uint4 ea = ring[(ring_head + (uint32_t)slot) & (kRing - 1)];
                uint4 eb = ring[(ring_head + (uint32_t)(slot + S)) & (kRing - 1)];
                const bool fit_a = have_a && win_valid && ea.x >= win_lo && (ea.x - win_lo) <= kWin - ea.y;
                const bool fit_b = have_b && win_valid && eb.x >= win_lo && (eb.x - win_lo) <= kWin - eb.y;
                fit_any = __ballot_sync(kFull, fit_a || fit_b);
                if (fit_any == 0u || cnt >= kCntMax) {
                    low_pp = min(have_a ? ea.x : 0xFFFFFFFFu, have_b ? eb.x : 0xFFFFFFFFu);
                    break;
                }
                const uint32_t unfit_a = __ballot_sync(kFull, have_a && !fit_a && wl == 0);
                const uint32_t unfit_b = __ballot_sync(kFull, have_b && !fit_b && wl == 0);
                if (unfit_a | unfit_b) {                                 // re-queue pieces that wait for a window move
                    if (have_a && !fit_a && wl == 0) ring[(ring_tail + __popc(unfit_a & lt_mask)) & (kRing - 1)] = ea;
                    if (have_b && !fit_b && wl == 0)
                        ring[(ring_tail + __popc(unfit_a) + __popc(unfit_b & lt_mask)) & (kRing - 1)] = eb;
                    ring_tail += __popc(unfit_a) + __popc(unfit_b);
                    __syncwarp();
                }
                ring_head += adv;
                uint32_t xa[kW][4], xb[kW][4];
                piece_words<G, HAS_OK>(xa, ea.x, fit_a ? ea.y : 0u, win_lo, wl, sq, okb, (int)ea.z);
                piece_words<G, HAS_OK>(xb, eb.x, fit_b ? eb.y : 0u, win_lo, wl, sq, okb, (int)eb.z);
                accumulate2(xa, xb);
            }
            if (ring_head == ring_tail) break;
            if (cnt) do_flush();
            if (fit_any == 0u) {                                         // nobody fits: move the window
                win_lo = __reduce_min_sync(kFull, low_pp) & ~31u;
                win_valid = true;
            }
        }

        // ---- general loop: CIGAR state machine over the remaining reads of the block
        if (rest_mask) {
            if ((rest_mask >> lane) & 1u) order2[__popc(rest_mask & lt_mask)] = (uint8_t)lane;
            __syncwarp();
            const uint32_t n_rest = __popc(rest_mask);
            uint32_t cursor = 0;
            // per read-slot state (replicated over the slot's G lanes)
            uint32_t cur = 0, cend = 0, ref_pos = 0, read_pos = 0, wbase = 0, nwords = 0;
            int sidx = -1;
            bool exhausted = false;
            uint32_t pp = 0, pq = 0, pn = 0;        // pending M/=/X piece: ref pos, read pos, length
            for (;;) {
                // A: read slots that finished their read pull the next ones, in order
                const bool need = (pn == 0u) && (cur == cend) && !exhausted;
                const uint32_t need_mask = __ballot_sync(kFull, need && wl == 0);
                if (need_mask) {
                    const uint32_t idx = cursor + __popc(need_mask & slot_lead_below);
                    const bool take = need && idx < n_rest;
                    const int src = take ? (int)order2[idx] : 0;
                    const uint32_t t_start = __shfl_sync(kFull, m0.start, src);
                    const uint32_t t_cbase = __shfl_sync(kFull, m0.cbase, src);
                    const uint32_t t_ncig = __shfl_sync(kFull, m0.ncig(), src);
                    const uint32_t t_wbase = __shfl_sync(kFull, m0.wbase, src);
                    const uint32_t t_nwords = __shfl_sync(kFull, m0.nwords(), src);
                    const int t_sidx = __shfl_sync(kFull, my_sidx, src);
                    if (take) {
                        ref_pos = t_start;
                        cur = t_cbase;
                        cend = t_cbase + t_ncig;
                        wbase = t_wbase;
                        nwords = t_nwords;
                        sidx = t_sidx;
                        read_pos = 0;
                    } else if (need) {
                        exhausted = true;
                    }
                    cursor += __popc(need_mask);
                }
                // B: no piece pending -> consume one CIGAR op (count.cpp:40-96)
                if (pn == 0u && cur < cend) {
                    const uint32_t ci = cur - rg.c_lo;
                    const uint32_t cw = ci < rg.c_n ? cg[ci] : __ldg(bv.cigar + cur);
                    cur++;
                    const uint32_t op = cw & 0xFu, len = cw >> 4;
                    if (op_is_match(op)) {                                   // count.cpp:51
                        pp = ref_pos;
                        pq = read_pos;
                        pn = len;
                        ref_pos = sat_add(ref_pos, len);
                        read_pos = sat_add(read_pos, len);
                        if (pn && (pp >= ref_len || pn > ref_len - pp)) {    // would index past the matrix
                            if (wl == 0) cv.status[kStatMaybeOverflow] = 1u;
                            pn = pp < ref_len ? ref_len - pp : 0u;
                        }
                    } else if (op == 1u) {                                   // insertion, count.cpp:74
                        read_pos = sat_add(read_pos, len);
                    } else if (op_is_refskip(op)) {                          // deletion / skip, count.cpp:80-87
                        uint32_t lim = ref_pos < ref_len ? min(len, ref_len - ref_pos) : 0u;
                        if (lim < len && wl == 0) cv.status[kStatIndexError] = 1u;
                        uint32_t *ds = cv.counts + (uint64_t)kPlaneDS * cv.stride + col0 + ref_pos;
                        for (uint32_t t = wl; t < lim; t += G) atomicAdd(ds + t, 1u);
                        ref_pos = sat_add(ref_pos, len);
                    }                                                        // S,H,P,B: ignored, count.cpp:92-95
                }
                // C: done with the block?
                const bool active = pn > 0u;
                if (!__any_sync(kFull, active || cur < cend || !exhausted)) break;
                // D: which pieces can go into the current window?
                const bool fits = active && win_valid && pp >= win_lo && (pp - win_lo) < kWin &&
                                  (pn <= kWin - (pp - win_lo) || pn > kMaxFit);
                const uint32_t fit_any = __ballot_sync(kFull, fits);
                if (fit_any == 0u || cnt >= kCntMax) {                       // the one flush site of this loop
                    if (fit_any == 0u && __ballot_sync(kFull, active) == 0u) continue;   // still walking ops / fetching
                    if (cnt) do_flush();
                    if (fit_any == 0u) {
                        win_lo = __reduce_min_sync(kFull, active ? pp : 0xFFFFFFFFu) & ~31u;
                        win_valid = true;
                    }
                    continue;
                }
                // E: masked words of the fitting pieces, added to the vertical counters
                uint32_t x[kW][4], zero[kW][4] = {};
                const uint32_t n1 = fits ? min(pn, kWin - (pp - win_lo)) : 0u;
                if (sidx >= 0) piece_words<G, HAS_OK>(x, pp, n1, win_lo, wl, sq, okb, sidx * 32 + (int)pq);
                else piece_words_global<G, HAS_OK>(x, pp, n1, win_lo, wl, bv, pq, wbase, nwords);
                accumulate2(x, zero);
                pp += n1;
                pq += n1;
                pn -= n1;
            }
        }

        // ---- stage b is free again: refill it with block j+2, rotate the metadata pipeline
        __syncwarp();
        m0 = m1;
        m1 = m2;
        rg = rg_nxt;
        if (j + 2 < nblk) rg_nxt = issue_block(j + 2, m1, b);
        m2 = load_meta(j + 3);
    }
    if (cnt) do_flush();
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

// Sparse corrections for letters outside ACGT (see bc_batch.exc_* in the header).
__global__ void k1_exceptions(BatchView bv, CountView cv)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= bv.n_exc) return;
    const uint32_t i = bv.exc_read[e];
    const uint32_t pos = bv.exc_pos[e] >> 2, flags = bv.exc_pos[e] & 3u;
    if (i >= bv.n_reads) return;
    uint32_t col;
    if (!read_pos_to_col(bv, i, pos, &col)) return;
    const uint32_t r = slot_of_read(bv.ref_read_off, bv.n_refs, i);
    const uint64_t base = cv.col_base[r];
    if (col >= cv.ref_len[r]) {
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
