// k1_split.cuh -- K1, count variant 2 (experimental): the CIGAR walk and the counting run in DIFFERENT warps.
//
// Why (profiles/r1_t_k1_summary.md): in k1_count_tiled one warp alternates between
// two kinds of code.  The per-block machinery (metadata, mbarrier wait, CIGAR decode, votes, ring push,
// TMA issue: 44 % of the instructions, 53 % of the stall samples) is a chain of dependent shuffles,
// votes and shared-memory loads; the trips (41 %) are straight-line LOP3 work bound by the ALU pipe.
// At 164 registers only three warps per scheduler are resident, so the latency of the first kind is
// not hidden and the kernel issues on 60 % of the cycles.  Here a chunk of reads is served by a PAIR
// of warps:
//   * the walking warp (warps 0..3 of the CTA, one warpgroup) does everything up to the piece ring
//     and needs few registers (setmaxnreg.dec);
//   * the counting warp (warps 4..7) owns the bit-sliced counters, consumes the ring one trip at a
//     time and flushes (setmaxnreg.inc takes the registers the walkers gave back).
// The two meet in the same shared-memory ring k1_count_tiled uses, cut into NS = ring / Q trip slots
// with a full / empty mbarrier pair per slot.  Window moves, the end of the chunk and the stage refills
// travel in band: a 16-byte descriptor per trip slot ("flush, then this window" / "flush and leave" /
// "after this trip refill stage s"), written by the walker before it publishes the trip.  A TMA stage is
// refilled by the COUNTING warp the moment the last trip that reads it is in registers, from the ranges the
// walker left beside the stage; the walker only waits for ring space and for its own stage's data.
// Block metadata reaches the walker through cp.async into a small shared-memory ring, and every
// shared-memory access is "pair base + index" with the array offset as an instruction immediate.
//
// Semantics are those of k1_count_tiled (same decode, same pieces, same flush); only who runs them changed.
// Measured: bit-exact, 103-107 us against 90 us on the bench workload (profiles/r3_b_k1_split_summary.md).
#pragma once
#include "k1_count.cuh"

namespace bc {

#ifndef BC_K1S_MINCTAS
#define BC_K1S_MINCTAS 3          // CTAs per SM the register split below is sized for: 12 warp pairs per SM
#define BC_K1S_REG_WALK 64        // registers per thread of a walking warp
#define BC_K1S_REG_COUNT 96       // ... of a counting warp: 4 x 32 x (64 + 96) = 256 x 80 = a CTA's launch allocation
#endif
#ifndef BC_K1S_WAIT_HINT
#define BC_K1S_WAIT_HINT 0        // > 0: suspend-time hint (ns) of the barrier waits instead of a plain spin
#endif
#ifndef BC_K1S_TRIPS
#define BC_K1S_TRIPS 4            // trip slots in the ring (how far the walker may run ahead)
#endif
constexpr int kSplitPairs = 4;                    // setmaxnreg works on warpgroups: 4 walkers + 4 counters
constexpr int kSplitThreads = 64 * kSplitPairs;
constexpr int kSplitMinCtas = BC_K1S_MINCTAS;

template <int G, bool HAS_OK>
struct K1SplitCfg {
    static constexpr uint32_t NS = BC_K1S_TRIPS > (32 / (4 * (32 / G))) ? BC_K1S_TRIPS : 32 / (4 * (32 / G)) * 2;  // trip slots
    using C = K1Cfg<G, HAS_OK, NS * 4u * (32u / G)>;
    static constexpr uint32_t full_off = C::warp_bytes;            // NS mbarriers: trip published
    static constexpr uint32_t empty_off = full_off + NS * 8u;      // NS mbarriers: trip consumed
    static constexpr uint32_t desc_off = empty_off + NS * 8u;      // NS x uint4 {command, new window, stage to refill + 1, -}
    static constexpr uint32_t meta_off = desc_off + NS * 16u;      // kStages x {cigar_off, seq_woff, start} x 32 lanes
    static constexpr uint32_t pair_bytes = meta_off + kStages * 384u;
    static constexpr uint32_t cta_bytes = C::lut_bytes + kSplitPairs * pair_bytes;
    static_assert((NS & (NS - 1u)) == 0u && NS >= 4u, "trip slots: a power of two, at least four");
    static_assert(pair_bytes % 16u == 0u, "TMA destinations are 16-byte aligned");
};
template <int G, bool HAS_OK>
__host__ __device__ constexpr uint32_t k1_split_cta_smem_bytes()
{
    return K1SplitCfg<G, HAS_OK>::cta_bytes;
}

enum SplitCommand : uint32_t { kCmdNone = 0u, kCmdMove = 1u, kCmdEnd = 2u };

// Shared-memory accessors with the region offset as an IMMEDIATE of the instruction: every address below is
// "the pair's base + a small index", so no register holds a per-array base (the walker has 64 registers).
template <uint32_t O> __device__ __forceinline__ uint32_t lds32o(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(O));
    return v;
}
template <uint32_t O> __device__ __forceinline__ uint2 lds64o(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(O));
    return v;
}
template <uint32_t O> __device__ __forceinline__ uint4 lds128o(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(O));
    return v;
}
template <uint32_t O> __device__ __forceinline__ void sts32o(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(O) : "memory");
}
template <uint32_t O> __device__ __forceinline__ void sts64o(uint32_t a, uint2 v)
{
    asm volatile("st.shared.v2.u32 [%0+%3], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y), "n"(O) : "memory");
}
template <uint32_t O> __device__ __forceinline__ void sts128o(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0+%5], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "n"(O) : "memory");
}
template <uint32_t O> __device__ __forceinline__ void mbar_arrive_o(uint32_t a)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(a), "n"(O) : "memory");
}
template <uint32_t O> __device__ __forceinline__ void mbar_wait_o(uint32_t a, uint32_t parity)
{
#if BC_K1S_WAIT_HINT > 0
    asm volatile(                                     // the waiting warp may sleep up to the hint (ns) per attempt
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0+%2], %1, %3;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity), "n"(O), "n"(BC_K1S_WAIT_HINT)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0+%2], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity), "n"(O)
        : "memory");
#endif
}
template <uint32_t O> __device__ __forceinline__ void cp_async4o(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1], 4;" ::"r"(dst), "l"(src), "n"(O) : "memory");
}

// Masked words of one piece for a lane's two window words (the `piece` of k1_count_tiled).
//   e.w: the pair's base + byte offset INSIDE the sequence stages of the plane word that holds window column 0
template <bool HAS_OK, uint32_t SEQ_OFF, uint32_t OK_OFF>
__device__ __forceinline__ void split_piece(const uint4 e, uint32_t (&x)[kW][kNC], int L0, uint32_t lutb,
                                            uint32_t lane_seq_off, uint32_t wb)
{
    const int a_c = __viaddmin_s32_relu((int)e.x, -L0, 64);
    const int e_c = __viaddmin_s32_relu((int)e.y, -L0, 64);
    const uint2 ga = lds64(lutb + 8u * (uint32_t)a_c), ge = lds64(lutb + 8u * (uint32_t)e_c);
    uint32_t m[kW] = {ga.x & ~ge.x, ga.y & ~ge.y};
    const uint32_t wa = e.w + lane_seq_off;
    const uint2 r0 = lds64o<SEQ_OFF>(wa), r1 = lds64o<SEQ_OFF + 8u>(wa), r2 = lds64o<SEQ_OFF + 16u>(wa);
    const uint32_t lo[kW] = {__funnelshift_r(r0.x, r1.x, e.z), __funnelshift_r(r1.x, r2.x, e.z)};
    const uint32_t hi[kW] = {__funnelshift_r(r0.y, r1.y, e.z), __funnelshift_r(r1.y, r2.y, e.z)};
    if (HAS_OK) {
        const uint32_t oa = wb + (uint32_t)((int)(wa - wb) >> 1);
        const uint32_t o0 = lds32o<OK_OFF>(oa), o1 = lds32o<OK_OFF + 4u>(oa), o2 = lds32o<OK_OFF + 8u>(oa);
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

// Four inputs per counter into the carry-save planes (the trip body of k1_count_tiled).
__device__ __forceinline__ void split_trip_add(uint32_t (&pl)[kW][kNC][kNR], uint32_t (&pa)[kW][kNC], uint32_t (&pb)[kW][kNC],
                                               const uint32_t (&x)[4][kW][kNC], uint32_t cnt, uint32_t spb)
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
        return;
    }
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
        return;
    }
    uint32_t c4[kW][kNC];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
            c4[w][k] = maj3(pl[w][k][3], pb[w][k], c3[w][k]);
            pl[w][k][3] ^= pb[w][k] ^ c3[w][k];
        }
    }
    if (!(cnt & 16u)) {                               // every 8th trip: park the weight-16 carry
#pragma unroll
        for (int w = 0; w < kW; w++)
#pragma unroll
            for (int k = 0; k < kNC; k++) sts32(spb + 128u * (uint32_t)((w * kNC + k) * 4), c4[w][k]);
        return;
    }
    const bool have = cnt >= 32u;                     // planes 5..7 were never written before trip 8
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kNC; k++) {
            const uint32_t qa = spb + 128u * (uint32_t)((w * kNC + k) * 4);
            const uint32_t pcv = lds32(qa);
            uint32_t c = maj3(pl[w][k][4], pcv, c4[w][k]);
            pl[w][k][4] ^= pcv ^ c4[w][k];
#pragma unroll
            for (int p = 1; p < 4; p++) {             // ripple the weight-32 carry upwards
                const uint32_t v = have ? lds32(qa + 128u * p) : 0u;
                sts32(qa + 128u * p, v ^ c);
                c &= v;
            }
        }
    }
}

template <int G, bool HAS_OK>
__global__ void __launch_bounds__(kSplitThreads, kSplitMinCtas)
k1_count_split(BatchView bv, CountView cv, const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t rpb)
{
    using P = K1SplitCfg<G, HAS_OK>;
    using C = typename P::C;
    constexpr int S = C::S;
    constexpr uint32_t Q = (uint32_t)C::Q, NS = P::NS;
    constexpr uint32_t kWin = C::kWin, kRing = C::kRing;
    extern __shared__ __align__(128) unsigned char k1_smem[];

    uint32_t lane_u;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
    const int lane = (int)lane_u;
    const int warp_in_cta = threadIdx.x >> 5;
    const bool counting = warp_in_cta >= kSplitPairs;
    const int pair = warp_in_cta & (kSplitPairs - 1);

    uint2 *lut = reinterpret_cast<uint2 *>(k1_smem);
    for (int v = threadIdx.x; v <= 64; v += kSplitThreads)
        lut[v] = make_uint2(v < 32 ? 0xFFFFFFFFu << v : 0u, v <= 32 ? 0xFFFFFFFFu : (v < 64 ? 0xFFFFFFFFu << (v - 32) : 0u));

    unsigned char *wsm = k1_smem + C::lut_bytes + (size_t)pair * P::pair_bytes;
    uint16_t *frow = reinterpret_cast<uint16_t *>(wsm + C::frow_off);
    const uint32_t lutb = opaque(smem_u32(k1_smem));
    const uint32_t wb = opaque(smem_u32(wsm));                                 // the pair's region
    static_assert(C::ring_off == 0u, "ring entries are addressed from the pair's base");
    if (!counting && lane == 0) {                                              // the walker sets up its pair's barriers
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init_s(wb + C::bar_off + 8u * s, 1);
        for (uint32_t s = 0; s < NS; s++) {
            mbar_init_s(wb + P::full_off + 8u * s, 1);
            mbar_init_s(wb + P::empty_off + 8u * s, 1);
            sts128(wb + P::desc_off + 16u * s, make_uint4(kCmdNone, 0u, 0u, 0u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                // the only CTA-wide barrier

    const uint32_t chunk_id = blockIdx.x * kSplitPairs + pair;
    Chunk ch = {0u, 0u, 0u, 0u};
    if (chunk_id < n_chunks) ch = chunks[chunk_id];
    const uint32_t ref_len = ch.ref_len;
    const uint32_t rb = ch.read_begin, re = ch.read_end;
    const bool have_chunk = re > rb;                // the same for both warps of the pair
    uint32_t *const plane0 = cv.counts + ch.col_base;                       // plane A, column 0 of this slot

    // The register split.  Every warp of a warpgroup executes its setmaxnreg (also the pairs without a
    // chunk), and the two roles never join again: ptxas allocates each branch under its own limit.
    // =========================================================================== counting warp
    if (counting) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BC_K1S_REG_COUNT));
        if (!have_chunk) return;
        const int slot = lane / G, wl = lane % G;
        const int L0 = 32 * kW * wl;
        const uint32_t spb = opaque(wb + C::frow_off + 4u * (uint32_t)lane);
        const uint32_t trip_ringb = opaque(wb + 16u * (uint32_t)slot);
        const uint32_t lane_seq_off = 8u * kW * (uint32_t)wl;
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
        uint32_t cnt = 0, win_lo = 0;
        for (uint32_t t = 0;; t++) {
            const uint32_t s = t & (NS - 1u);
            mbar_wait_o<P::full_off>(wb + 8u * s, (t / NS) & 1u);
            const uint4 d = lds128o<P::desc_off>(wb + 16u * s);
            if (d.x != kCmdNone || cnt == kCntMax) {                           // rare: everything but a plain trip
                if (cnt != 0u) {
                    flush_counters<G>(pl, pa, pb, cnt, frow, plane0 + win_lo, cv.stride, lane);
                    cnt = 0u;
                }
                if (d.x == kCmdEnd) break;
                if (d.x == kCmdMove) win_lo = d.y;
            }
            const uint32_t ea = trip_ringb + 16u * Q * s;
            uint4 e[4];
#pragma unroll
            for (int q = 0; q < 4; q++) e[q] = lds128(ea + 16u * (uint32_t)(q * S));
            uint32_t x[4][kW][kNC];
#pragma unroll
            for (int q = 0; q < 4; q++) split_piece<HAS_OK, C::seq_off, C::ok_off>(e[q], x[q], L0, lutb, lane_seq_off, wb);
            split_trip_add(pl, pa, pb, x, cnt, spb);
            __syncwarp();                                                      // every lane is through with the trip's entries and data
            if (lane == 0) {
                if ((d.x | d.z) != 0u) sts128o<P::desc_off>(wb + 16u * s, make_uint4(kCmdNone, 0u, 0u, 0u));
                if (d.z != 0u) {
                    // this was the last trip that reads stage d.z - 1: refill it with the block whose ranges the
                    // walker left beside the stage (the walker waits on the stage's own barrier as before)
                    const uint32_t stg = d.z - 1u;
                    const uint4 rg = lds128o<C::rng_off>(wb + 16u * stg);      // s_lo, s_n, c_lo, c_n
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const uint32_t bar = wb + C::bar_off + 8u * stg;
                    const uint32_t seqb = wb + C::seq_off, okb = wb + C::ok_off, cigb = wb + C::cig_off;
                    mbar_expect_tx_s(bar, rg.y * 8u + (HAS_OK ? rg.y * 4u : 0u) + rg.w * 4u);
                    if (rg.y) {
                        bulk_g2s_s(seqb + stg * (kSeqCap * 8u), bv.planes + rg.x, rg.y * 8u, bar);
                        if (HAS_OK) bulk_g2s_s(okb + stg * (kSeqCap * 4u), bv.okmask + rg.x, rg.y * 4u, bar);
                    }
                    if (rg.w) bulk_g2s_s(cigb + stg * (kCigCap * 4u), bv.cigar + rg.z, rg.w * 4u, bar);
                }
                mbar_arrive_o<P::empty_off>(wb + 8u * s);
            }
            cnt += 4u;
        }
        return;
    }

    // =========================================================================== walking warp
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BC_K1S_REG_WALK));
    if (!have_chunk) return;
    const uint32_t lt_mask = opaque((1u << lane) - 1u);
    const uint32_t nblk = (re - rb + rpb - 1) / rpb;
    uint32_t *const ds_plane = plane0 + (uint64_t)kPlaneDS * cv.stride;

    // Block metadata, one read per lane: {cigar_off, seq_woff, start} of block b go to slot b % kStages of a
    // small shared-memory ring with cp.async (no register is held while the loads are in flight), three
    // blocks ahead; word l + 1 of an offset array is read l's end.
    auto fetch_meta = [&](uint32_t blk, uint32_t slot) {
        const uint32_t a = wb + slot * 384u + 4u * (uint32_t)lane;
        const uint32_t idx = rb + blk * rpb + (uint32_t)lane;
        if (blk < nblk && idx <= re) {
            cp_async4o<P::meta_off>(a, bv.cigar_off + idx);
            cp_async4o<P::meta_off + 128u>(a, bv.seq_woff + idx);
            if (idx < re)
                cp_async4o<P::meta_off + 256u>(a, bv.starts + idx);
            else
                sts32o<P::meta_off + 256u>(a, 0u);
        } else {
            sts32o<P::meta_off>(a, 0u);
            sts32o<P::meta_off + 128u>(a, 0u);
            sts32o<P::meta_off + 256u>(a, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto meta_landed = [&]() {                      // every lane's copies are done and visible to the warp
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    };
    // What stage `stg` is to hold for block `blk` {first plane word, plane words, first CIGAR word, CIGAR words}
    // goes beside the stage; with `now` the walker also starts the copies itself.
    auto issue_block = [&](uint32_t blk, uint32_t stg, bool now) {
        if (lane == 0) {
            const uint32_t nvalid = min(rpb, re - (rb + blk * rpb));
            const uint32_t ma = wb + stg * 384u;                     // (a block's metadata slot is its stage number)
            const uint32_t c0 = lds32o<P::meta_off>(ma), c1 = lds32o<P::meta_off>(ma + 4u * nvalid);
            const uint32_t s0 = lds32o<P::meta_off + 128u>(ma), s1 = lds32o<P::meta_off + 128u>(ma + 4u * nvalid);
            const uint32_t s_lo = s0 & ~3u, s_n = min(((s1 + 3u) & ~3u) - s_lo, kSeqCap);
            const uint32_t c_lo = c0 & ~3u, c_n = min(((c1 + 3u) & ~3u) - c_lo, kCigCap);
            sts128o<C::rng_off>(wb + 16u * stg, make_uint4(s_lo, s_n, c_lo, c_n));
            if (now) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t bar = wb + C::bar_off + 8u * stg;
                const uint32_t seqb = wb + C::seq_off, okb = wb + C::ok_off, cigb = wb + C::cig_off;
                mbar_expect_tx_s(bar, s_n * 8u + (HAS_OK ? s_n * 4u : 0u) + c_n * 4u);
                if (s_n) {
                    bulk_g2s_s(seqb + stg * (kSeqCap * 8u), bv.planes + s_lo, s_n * 8u, bar);
                    if (HAS_OK) bulk_g2s_s(okb + stg * (kSeqCap * 4u), bv.okmask + s_lo, s_n * 4u, bar);
                }
                if (c_n) bulk_g2s_s(cigb + stg * (kCigCap * 4u), bv.cigar + c_lo, c_n * 4u, bar);
            }
        }
    };

    fetch_meta(0, 0);
    fetch_meta(1, 1);
    fetch_meta(2, 2);
    meta_landed();
    issue_block(0, 0, true);
    if (nblk > 1) issue_block(1, 1, true);
    if (nblk > 2) issue_block(2, 2, true);
    uint32_t phases = 0, st = 0;

    // ---- the ring as the walker sees it
    uint32_t tail = 0;                              // entries written
    uint32_t pub = 0;                               // trips published (complete trips below tail)
    uint32_t acq = NS;                              // trips below acq may be written (their slots are free)
    uint32_t win_lo = 0, pend_lo = 0;
    uint32_t refill_trip = 0xFFFFFFFFu;             // the open trip a stage refill rides on, if any
    bool win_valid = false, pend_move = false;

    auto acquire_through = [&](uint32_t trip) {     // wait until trips <= trip may be written, i.e. trip - NS is consumed
        while (acq <= trip) {
            mbar_wait_o<P::empty_off>(wb + 8u * (acq & (NS - 1u)), ((acq / NS) - 1u) & 1u);
            acq++;
        }
    };
    auto publish = [&]() {                          // hand every complete trip to the counting warp
        const uint32_t np = tail / Q;
        if (np != pub) {
            __syncwarp();                           // all lanes' entries (and lane 0's command) come before the arrive
            if (lane == 0)
                for (uint32_t t = pub; t < np; t++) mbar_arrive_o<P::full_off>(wb + 8u * (t & (NS - 1u)));
            pub = np;
        }
    };
    auto open_for = [&](uint32_t n) {               // room for n more entries; a pending window move rides on their first trip
        acquire_through((tail + n - 1u) / Q);
        if (pend_move) {                            // (a move pads, so tail is a trip boundary here)
            if (lane == 0) sts64o<P::desc_off>(wb + 16u * ((tail / Q) & (NS - 1u)), make_uint2(kCmdMove, pend_lo));
            pend_move = false;
        }
    };
    auto pad = [&]() {                              // fill the open trip with empty pieces and publish it
        const uint32_t r = tail & (Q - 1u);
        if (r != 0u) {
            if ((uint32_t)lane < Q - r) sts128(wb + 16u * ((tail + lane) & (kRing - 1u)), make_uint4(0u, 0u, 0u, wb));
            tail += Q - r;
            publish();
        }
    };
    auto move_window = [&](uint32_t new_lo) {
        pad();
        pend_move = true;
        pend_lo = new_lo;
        win_lo = new_lo;
        win_valid = true;
    };
    auto quiesce = [&]() {                          // everything pushed so far is counted (its stage data may go)
        pad();
        if (pub != 0u) acquire_through(pub - 1u + NS);
    };
    auto make_entry = [&](uint32_t rel, uint32_t n, int qbit) {
        const int z = qbit - (int)rel;
        return make_uint4(rel, rel + n, (uint32_t)z, wb + (uint32_t)((z >> 5) * 8));
    };

    for (uint32_t j = 0; j < nblk; j++) {
        // this block's metadata out of its slot, then the slot goes to block j+3
        const uint32_t ma = wb + st * 384u + 4u * (uint32_t)lane, ma1 = wb + st * 384u + 4u * (uint32_t)((lane + 1) & 31);
        const uint32_t cbase = lds32o<P::meta_off>(ma), cend_all = lds32o<P::meta_off>(ma1);
        const uint32_t wbase = lds32o<P::meta_off + 128u>(ma), wend = lds32o<P::meta_off + 128u>(ma1);
        const uint32_t start0 = lds32o<P::meta_off + 256u>(ma);
        __syncwarp();
        fetch_meta(j + 3u, st);
        mbar_wait_o<C::bar_off>(wb + 8u * st, (phases >> st) & 1u);
        phases ^= 1u << st;
        const uint32_t nvalid = min(rpb, re - (rb + j * rpb));
        __syncwarp();
        const uint4 rg = lds128o<C::rng_off>(wb + 16u * st);
        const uint32_t cg = wb + st * (kCigCap * 4u) - rg.z * 4u;       // (+ C::cig_off in the loads) CIGAR word 0
        const int seg_bit0 = (int)(st * kSeqCap * 32u);

        // ---- per-lane read state (count.cpp:35-38)
        const bool valid = (uint32_t)lane < nvalid;
        const bool staged = valid && (wend - rg.x) <= rg.y;
        const bool cig_staged = (cend_all - rg.z) <= rg.w;
        uint32_t unst = __ballot_sync(kFull, valid && !staged && cend_all > cbase);
        uint32_t cur = cbase, cend = staged ? cend_all : cbase;
        uint32_t rpos = min(start0, ref_len), rem = 0u, ds_pos = 0u, ds_n = 0u;
        int qb = seg_bit0 + (int)((wbase - rg.x) * 32u);
        int qend = qb + (int)((wend - wbase) * 32u);
        int slow_lane = -1;
        uint32_t seg_w = 0u, seg_end = 0u;

        // ---- straight-line decode of reads with at most three CIGAR ops (see k1_count_tiled)
        uint32_t nA = 0u, nB = 0u, ppB = 0u;
        int pqB = 0;
        bool fast = false;
        {
            const uint32_t ncig = cend_all - cbase;
            bool bad = unst != 0u || (valid && (!cig_staged || ncig > 3u));
            uint32_t sk_pos = 0u, sk_n = 0u;
            if (staged) {
                const uint32_t ca = cg + cbase * 4u;
                uint32_t cw[3];
                cw[0] = ncig > 0u ? lds32o<C::cig_off>(ca) : 0u;
                cw[1] = ncig > 1u ? lds32o<C::cig_off + 4u>(ca) : 0u;
                cw[2] = ncig > 2u ? lds32o<C::cig_off + 8u>(ca) : 0u;
                uint32_t r[4], q[4], mlen[3], dlen[3];
                bool brk[3];
                r[0] = rpos;
                q[0] = (uint32_t)qb;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const uint32_t len = cw[k] >> 4;
                    const uint32_t cls = (kOpClass >> ((cw[k] << 1) & 30u)) & 3u;
                    r[k + 1] = r[k] + ((cls & 1u) ? len : 0u);
                    q[k + 1] = q[k] + (((cls + 1u) & 2u) ? len : 0u);
                    mlen[k] = cls == 1u ? len : 0u;
                    dlen[k] = cls == 3u ? len : 0u;
                    brk[k] = cls >= 2u && len != 0u;
                }
                nA = mlen[0] + (brk[0] ? 0u : mlen[1]) + ((brk[0] || brk[1]) ? 0u : mlen[2]);
                nB = (brk[0] ? mlen[1] : 0u) + ((brk[0] != brk[1]) ? mlen[2] : 0u);
                ppB = brk[0] ? r[1] : r[2];
                pqB = (int)(brk[0] ? q[1] : q[2]);
                sk_n = dlen[0] + dlen[1] + dlen[2];
                sk_pos = dlen[0] ? r[0] : (dlen[1] ? r[1] : r[2]);
                const uint32_t nsk = (dlen[0] ? 1u : 0u) + (dlen[1] ? 1u : 0u) + (dlen[2] ? 1u : 0u);
                bad = bad || r[3] > ref_len || q[3] > (uint32_t)qend || (brk[0] && brk[1] && mlen[2] != 0u) || nsk > 1u ||
                      sk_n > kLaneSkipMax || nA > C::kMaxFit || nB > C::kMaxFit;
            }
            fast = !__any_sync(kFull, bad);
            if (fast) {
                if (sk_n) {                                          // count.cpp:80-87
                    uint32_t *const dp = ds_plane + sk_pos;
                    red_add(dp, 1u);
                    if (sk_n > 1u) red_add(dp + 1, 1u);
                    if (sk_n > 2u) red_add(dp + 2, 1u);
#pragma unroll 1
                    for (uint32_t t = 3u; t < sk_n; t++) red_add(dp + t, 1u);
                }
            } else {
                nA = 0u;
                nB = 0u;
            }
        }

        if (fast) {
            // ---- push the pieces that fit the window (run A first, then run B); move the window to the
            //      lowest piece that is left and go round again
            for (;;) {
                const uint32_t relA = rpos - win_lo, relB = ppB - win_lo;
                const bool fitA = nA != 0u && win_valid && relA <= kWin - nA;
                const bool fitB = nB != 0u && win_valid && relB <= kWin - nB;
                const uint32_t mA = __ballot_sync(kFull, fitA), mB = __ballot_sync(kFull, fitB);
                if (mA) {
                    open_for(__popc(mA));
                    if (fitA) {
                        sts128(wb + 16u * ((tail + __popc(mA & lt_mask)) & (kRing - 1u)), make_entry(relA, nA, qb));
                        nA = 0u;
                    }
                    tail += __popc(mA);
                    publish();
                }
                if (mB) {
                    open_for(__popc(mB));
                    if (fitB) {
                        sts128(wb + 16u * ((tail + __popc(mB & lt_mask)) & (kRing - 1u)), make_entry(relB, nB, pqB));
                        nB = 0u;
                    }
                    tail += __popc(mB);
                    publish();
                }
                if (!__any_sync(kFull, (nA | nB) != 0u)) break;
                move_window(__reduce_min_sync(kFull, min(nA ? rpos : 0xFFFFFFFFu, nB ? ppB : 0xFFFFFFFFu)) & ~31u);
            }
        } else {
            for (;;) {
                // ---- F: fetch CIGAR ops until an M/=/X run is open (count.cpp:40-96)
                bool moved = false;
                while (rem == 0u && ds_n == 0u && cur < cend) {
                    const uint32_t cw = (cur - rg.z) < rg.w ? lds32o<C::cig_off>(cg + cur * 4u) : __ldg(bv.cigar + cur);
                    cur++;
                    moved = true;
                    const uint32_t len = cw >> 4;
                    const uint32_t cls = (kOpClass >> ((cw << 1) & 30u)) & 3u;
                    if (cls == 1u) {                                     // M / = / X, count.cpp:51
                        const uint32_t lim = ref_len - rpos;
                        if (len > lim) cv.status[kStatMaybeOverflow] = 1u;
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
                    }
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
                    open_for(__popc(pm));
                    if (n1) {
                        sts128(wb + 16u * ((tail + __popc(pm & lt_mask)) & (kRing - 1u)), make_entry(relp, n1, qb));
                        rpos += n1;
                        qb += (int)n1;
                        rem -= n1;
                        moved = true;
                    }
                    tail += __popc(pm);
                    publish();
                }
                if (__any_sync(kFull, cur < cend || rem != 0u)) {
                    if (__any_sync(kFull, moved)) continue;
                    const bool wst = rem != 0u && qb < qend;         // waits for the window (not for data)
                    if (__any_sync(kFull, wst)) {
                        move_window(__reduce_min_sync(kFull, wst ? rpos : 0xFFFFFFFFu) & ~31u);
                        continue;
                    }
                }
                // ---- pass over: reads that were not staged are copied into this stage segment by segment,
                //      once everything that reads the stage has been counted
                if (unst == 0u && slow_lane < 0) break;
                quiesce();
                if (slow_lane >= 0) {
                    const int done = __shfl_sync(kFull, (int)(cur >= cend && rem == 0u), slow_lane);
                    const int qb_u = __shfl_sync(kFull, qb, slow_lane);
                    const uint32_t adv = (uint32_t)(qb_u - seg_bit0) >> 5;
                    if (done || adv == 0u || adv >= seg_end - seg_w) {
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
                        sts64o<C::seq_off>(wb + (st * kSeqCap + i) * 8u, __ldg(bv.planes + seg_w + i));
                        if (HAS_OK) sts32o<C::ok_off>(wb + (st * kSeqCap + i) * 4u, __ldg(bv.okmask + seg_w + i));
                    }
                    if (lane == slow_lane) qend = seg_bit0 + (int)(nw * 32u);
                    __syncwarp();
                }
            }
        }

        // ---- this block's stage is refilled with block j+3 as soon as the last trip that reads it has been
        //      counted.  That trip is normally still open (the next block's pieces complete it): the refill
        //      rides on it and the counting warp starts the copies.  If it is already published the walker
        //      waits for it and starts them itself.
        if (j + 3u < nblk) {
            bool open = (tail & (Q - 1u)) != 0u;
            if (open && refill_trip == pub) {                        // (a block that did not even complete the open trip)
                pad();
                open = false;
            }
            if (!open && pub != 0u) acquire_through(pub - 1u + NS);
            meta_landed();
            issue_block(j + 3u, st, !open);
            if (open) {
                if (lane == 0) sts32o<P::desc_off + 8u>(wb + 16u * (pub & (NS - 1u)), st + 1u);
                refill_trip = pub;
            }
        }
        else meta_landed();                                          // (keeps "the next slot has landed" unconditional)
        st = (st == (uint32_t)kStages - 1u) ? 0u : st + 1u;
    }

    // ---- the end of the chunk: close the open trip, then "flush and leave" on a trip of its own
    pad();
    acquire_through(pub);
    if (lane == 0) {
        sts64o<P::desc_off>(wb + 16u * (pub & (NS - 1u)), make_uint2(kCmdEnd, 0u));
        mbar_arrive_o<P::full_off>(wb + 8u * (pub & (NS - 1u)));
    }
}

}  // namespace bc
