// k1_fast.cuh -- K1, the lean counting kernel: CIGAR walk + per-position base counting on sm_100a for the
// reads a straight-line decode can take (at most three CIGAR ops: M, M-I-M, M-D-M, clipped and =/X
// spellings -- practically every short read).  Replaces the loop nest of the reference operator
// (basecount/count.cpp:22-97) together with k1_count_tiled (k1_count.cuh), which stays as the general walker:
// a block of reads this kernel does not take (more ops, a run longer than a window, clipping at the
// reference end, reads longer than a pipeline stage) is appended, untouched, to a list of deferred chunks
// that k1_count_tiled processes right afterwards.  The decode has no side effects before a block commits,
// so there is exactly one kernel that defines the semantics of the rare cases.
//
// Same counting as k1_count_tiled (bit-sliced vertical counters, pieces in a shared-memory ring, trips of
// four pieces per read slot, bit-sliced slot combine in the flush, coalesced RED.ADD); what is different is
// everything per block of reads, which was 44 % of the instructions of k1_count_tiled:
//   * 32 reads per block, one per lane; {start, CIGAR range, sequence range} come from five coalesced loads
//     two blocks ahead (indices clamped to the block's end, so a lane without a read holds an empty read and
//     the last lane always holds the block's end offsets: no shuffles, no validity predicates);
//   * the (at most three) CIGAR words of a read are loaded straight from HBM one block ahead -- no CIGAR
//     staging, no staged-range bookkeeping;
//   * sequence words are staged by one 1-D TMA bulk copy per block (two with a quality mask) into a ring of
//     stages; the range a stage holds is recomputed from the metadata instead of being parked in shared memory;
//   * the decode is multiply-add arithmetic on a 2-bit "consumes reference / consumes query" code per op;
//   * the ring holds a whole block's pieces plus a partial trip, so pushes never wait for ring space, and an
//     invalid window is a window position no read can fit (no validity flag).
#pragma once
#include "k1_count.cuh"

namespace bc {

#ifndef BC_K1F_STAGES
#define BC_K1F_STAGES 3
#endif
#ifndef BC_K1F_MINCTAS
#define BC_K1F_MINCTAS 3
#endif
constexpr int kFastStages = BC_K1F_STAGES;
constexpr uint32_t kFastRing = 128;            // ring entries: <= 64 pieces of a block + a partial trip (< 32)
constexpr uint32_t kFastRpbMax = 32;
constexpr uint32_t kAdvCode = 0x3C05Bu;        // 2 bits per CIGAR op: bit 0 = consumes reference, bit 1 = consumes query
                                               // (M,=,X: 3; I: 2; D,N: 1; S,H,P,B: 0 -- soft clips are trimmed already)
constexpr uint32_t kNoWindow = 0x80000000u;    // a window position no piece fits (reference positions are < 2^31)

template <int G, bool HAS_OK>
struct K1FastCfg {
    static constexpr int S = 32 / G;
    static constexpr int Q = 4 * S;
    static constexpr uint32_t kWin = 32u * kW * G;
    static constexpr uint32_t kMaxFit = kWin - 31u;
    static constexpr uint32_t kCols = kFlushStride * kW * G;
    static constexpr uint32_t lut_bytes = 528;
    // per warp: ring | sequence stages | quality-mask stages | flush rows (= spill words) | mbarriers.  A piece's
    // unclamped word index reaches up to 64*G columns before or after its data: the ring in front and the flush
    // rows behind keep those (masked-away) reads inside the CTA's shared memory.
    static constexpr uint32_t ring_off = 0;
    static constexpr uint32_t seq_off = ring_off + kFastRing * 16u;
    static constexpr uint32_t ok_off = seq_off + kFastStages * kSeqCap * 8u;
    static constexpr uint32_t frow_off = ok_off + (HAS_OK ? kFastStages * kSeqCap * 4u : 0u);
    static constexpr uint32_t frow_bytes = kNC * kCols * 2u > kSpillWords * 128u ? kNC * kCols * 2u : kSpillWords * 128u;
    static constexpr uint32_t bar_off = frow_off + frow_bytes;
    static constexpr uint32_t warp_bytes = bar_off + 32u;
    static constexpr uint32_t cta_bytes = lut_bytes + kK1WarpsPerCta * warp_bytes;
    static_assert(seq_off % 16 == 0 && ok_off % 16 == 0 && frow_off % 16 == 0 && bar_off % 16 == 0 && warp_bytes % 16 == 0,
                  "TMA destinations are 16-byte aligned");
    static_assert(seq_off >= (kWin / 32 + 4) * 8 && frow_bytes >= (kWin / 32 + 4) * 8, "guard bands around the stages");
    static_assert(kFastRing % Q == 0 && kFastRing >= 64 + Q, "a block's pieces and a partial trip fit the ring");
    static_assert(kFastStages >= 2 && kFastStages <= 3, "stages");
};
template <int G, bool HAS_OK>
__host__ __device__ constexpr uint32_t k1_fast_cta_smem_bytes()
{
    return K1FastCfg<G, HAS_OK>::cta_bytes;
}

// Predicated read-only load without a branch: 0 when the predicate is off.
__device__ __forceinline__ uint32_t ldg_if(const uint32_t *p, bool on)
{
    uint32_t v = 0u;
    asm volatile("{\n.reg .pred P1;\nsetp.ne.u32 P1, %2, 0;\n@P1 ld.global.nc.u32 %0, [%1];\n}" : "+r"(v) : "l"(p), "r"((uint32_t)on));
    return v;
}

template <int G, bool HAS_OK>
__global__ void __launch_bounds__(kK1Threads, BC_K1F_MINCTAS)
k1_count_fast(BatchView bv, CountView cv, const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t rpb,
              Chunk *__restrict__ deferred, uint32_t *__restrict__ n_deferred)
{
    using C = K1FastCfg<G, HAS_OK>;
    constexpr int S = C::S, Q = C::Q;
    constexpr uint32_t kWin = C::kWin;
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

    const uint32_t warp_id = blockIdx.x * kK1WarpsPerCta + warp_in_cta;
    if (warp_id >= n_chunks) return;

    unsigned char *wsm = k1_smem + C::lut_bytes + (size_t)warp_in_cta * C::warp_bytes;
    uint16_t *frow = reinterpret_cast<uint16_t *>(wsm + C::frow_off);          // flush only
    const uint32_t lutb = opaque(smem_u32(k1_smem));
    const uint32_t wb = opaque(smem_u32(wsm));                                 // the warp's region
    const uint32_t ringb = wb + C::ring_off, seqb = wb + C::seq_off, okb = wb + C::ok_off, barb = wb + C::bar_off;

    const int slot = lane / G, wl = lane % G;
    const int L0 = 32 * kW * wl;                    // window column of this lane's bit 0
    const uint32_t lt_mask = opaque((1u << lane) - 1u);
    const uint32_t spb = opaque(wb + C::frow_off + 4u * (uint32_t)lane);       // this lane's spill words, 128 B apart
    const uint32_t trip_ringb = opaque(ringb + 16u * (uint32_t)slot);          // ring entry of this slot in a trip
    const uint32_t lane_seq_off = 8u * kW * (uint32_t)wl;                      // byte offset of this lane's window words

    const Chunk ch = chunks[warp_id];
    const uint32_t ref_len = ch.ref_len;
    const uint32_t rb = ch.read_begin, re = ch.read_end;
    if (re <= rb) return;
    const uint32_t nblk = (re - rb + rpb - 1) / rpb;
    uint32_t *const plane0 = cv.counts + ch.col_base;                       // plane A, column 0 of this slot
    uint32_t *const ds_plane = plane0 + (uint64_t)kPlaneDS * cv.stride;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kFastStages; s++) mbar_init_s(barb + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---- block metadata.  Lane l holds read (block begin + l); indices are clamped to the block's end, so a lane
    //      without a read holds an empty read (no CIGAR words, no sequence words) and lane 31's ends are the block's.
    struct Meta { uint32_t start, cb, ce, wb, we; };
    auto load_meta = [&](uint32_t blk) {
        const uint32_t b0 = rb + blk * rpb;                                  // (the host keeps n_reads < 2^32 - 128)
        const uint32_t bend = min(b0 + rpb, re);
        const uint32_t i0 = min(b0 + (uint32_t)lane, bend);
        const uint32_t i1 = min(i0 + 1u, bend);
        Meta m;
        m.cb = __ldg(bv.cigar_off + i0);
        m.ce = __ldg(bv.cigar_off + i1);
        m.wb = __ldg(bv.seq_woff + i0);
        m.we = __ldg(bv.seq_woff + i1);
        m.start = __ldg(bv.starts + min(i0, re - 1u));
        return m;
    };
    struct Cig { uint32_t c0, c1, c2; };
    auto load_cig = [&](const Meta &m) {                                     // a missing op reads as a zero-length M
        const uint32_t n = m.ce - m.cb;
        const uint32_t *p = bv.cigar + m.cb;
        Cig c;
        c.c0 = ldg_if(p, n > 0u);
        c.c1 = ldg_if(p + 1, n > 1u);
        c.c2 = ldg_if(p + 2, n > 2u);
        return c;
    };
    // Stage the sequence words of a block: [first word & ~3, last word rounded up) clipped to a stage.
    auto issue_block = [&](const Meta &m, uint32_t stg) {
        const uint32_t s0 = __shfl_sync(kFull, m.wb, 0), s1 = __shfl_sync(kFull, m.we, 31);
        if (lane == 0) {
            const uint32_t s_lo = s0 & ~3u, s_n = min(((s1 + 3u) & ~3u) - s_lo, kSeqCap);     // 32 B / 16 B aligned sources
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of this stage
            const uint32_t bar = barb + 8u * stg;
            mbar_expect_tx_s(bar, s_n * 8u + (HAS_OK ? s_n * 4u : 0u));
            if (s_n) {
                bulk_g2s_s(seqb + stg * (kSeqCap * 8u), bv.planes + s_lo, s_n * 8u, bar);
                if (HAS_OK) bulk_g2s_s(okb + stg * (kSeqCap * 4u), bv.okmask + s_lo, s_n * 4u, bar);
            }
        }
    };

    Meta M0 = load_meta(0), M1 = load_meta(1);
    issue_block(M0, 0);
    if (nblk > 1 && kFastStages > 2) issue_block(M1, 1);
    Cig C0 = load_cig(M0);

    uint32_t phases = 0;                            // bit s: parity to wait for on stage s
    uint32_t st = 0;                                // stage of the current block
    uint32_t ring_head = 0, ring_tail = 0, mark = 0;      // mark: entries before it come from earlier blocks
    uint32_t win_lo = kNoWindow, cnt = 0;
    uint32_t def_begin = 0xFFFFFFFFu, def_end = 0xFFFFFFFFu;   // open run of deferred blocks (reads [begin, end))
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
    auto emit_deferred = [&]() {
        if (lane == 0 && def_end > def_begin && def_begin != 0xFFFFFFFFu) {
            const uint32_t at = atomicAdd(n_deferred, 1u);
            Chunk c;
            c.read_begin = def_begin;
            c.read_end = def_end;
            c.col_base = ch.col_base;
            c.ref_len = ref_len;
            deferred[at] = c;
            atomicAdd(cv.status + kStatDeferredReads, def_end - def_begin);
        }
    };
    // A ring entry: x / y = first / end column of the piece relative to the window, z = bit index of window
    // column 0 in the staged data, w = shared address of the plane word that holds window column 0.
    auto push_entry = [&](uint32_t at, uint32_t rel, uint32_t n, int qbit) {
        const int z = qbit - (int)rel;
        sts128(ringb + 16u * (at & (kFastRing - 1u)), make_uint4(rel, rel + n, (uint32_t)z, seqb + (uint32_t)((z >> 5) * 8)));
    };

    // j == nblk is a virtual empty block: it drains the ring and does the final flush in the one trip / flush site.
    // Stage schedule: block j lives in stage j % kFastStages; block j + kFastStages - 1 is staged during block j, as
    // soon as the pieces of block j - 1 (the previous tenant of that stage) have all been counted.
    for (uint32_t j = 0; j <= nblk; j++) {
        const bool last = (j == nblk);
        const Meta M2 = load_meta(j + 2u);                                   // in flight during this block
        const Cig C1 = load_cig(M1);
        if (!last) {
            mbar_wait_s(barb + 8u * st, (phases >> st) & 1u);
            phases ^= 1u << st;
        }

        // ---- straight-line decode of at most three ops (count.cpp:35-96): runs of M/=/X merge into pieces; run A
        //      starts at the read start, run B right after the first non-empty I/D/N.
        const uint32_t s_lo = __shfl_sync(kFull, M0.wb, 0) & ~3u;
        const int qb = (int)(st * (kSeqCap * 32u) + (M0.wb - s_lo) * 32u);   // bit index of the read's first base
        uint32_t nA, nB, ppB, sk_n, sk_pos;
        int pqB;
        bool bad;
        {
            const uint32_t cw[3] = {C0.c0, C0.c1, C0.c2};
            uint32_t r[4], q[4], ml[3], dl[3];
            bool bk[3];
            r[0] = M0.start;
            q[0] = (uint32_t)qb;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint32_t len = cw[k] >> 4;
                const uint32_t code = kAdvCode >> ((cw[k] & 15u) * 2u);
                const uint32_t fr = code & 1u, fq = (code >> 1) & 1u;
                const uint32_t ra = len * fr, qa = len * fq;                 // count.cpp:67-68, 75, 87
                r[k + 1] = r[k] + ra;
                q[k + 1] = q[k] + qa;
                ml[k] = ra & (0u - fq);                                      // M/=/X
                dl[k] = ra - ml[k];                                          // D/N
                bk[k] = ra != qa;                                            // a non-empty I/D/N ends the match run
            }
            nA = ml[0] + (bk[0] ? 0u : ml[1]) + ((bk[0] || bk[1]) ? 0u : ml[2]);
            nB = (bk[0] ? ml[1] : 0u) + ((bk[0] != bk[1]) ? ml[2] : 0u);
            ppB = bk[0] ? r[1] : r[2];
            pqB = (int)(bk[0] ? q[1] : q[2]);
            sk_n = dl[0] + dl[1] + dl[2];
            sk_pos = dl[0] ? r[0] : (dl[1] ? r[1] : r[2]);
            const uint32_t qend = (uint32_t)qb + (M0.we - M0.wb) * 32u;      // end of the staged data of this read
            bad = (M0.ce - M0.cb) > 3u || (M0.we - s_lo) > kSeqCap              // more ops; not (fully) staged
                  || r[3] > ref_len || q[3] > qend                            // reference end; CIGAR overruns the read
                  || (bk[0] && bk[1] && ml[2] != 0u)                          // a third match run
                  || sk_n != max(dl[0], max(dl[1], dl[2])) || sk_n > kLaneSkipMax   // two D/N runs; a long one
                  || nA > C::kMaxFit || nB > C::kMaxFit;
        }
        if (__any_sync(kFull, bad)) {                                        // the whole block goes to the general walker
            nA = 0u;
            nB = 0u;
            sk_n = 0u;
            const uint32_t b0 = rb + j * rpb;
            if (def_end != b0) {
                emit_deferred();
                def_begin = b0;
            }
            def_end = min(b0 + rpb, re);
        }
        if (sk_n) {                                                          // count.cpp:80-87; short deletions are the rule
            uint32_t *const dp = ds_plane + sk_pos;
            red_add(dp, 1u);
            if (sk_n > 1u) red_add(dp + 1, 1u);
            if (sk_n > 2u) red_add(dp + 2, 1u);
#pragma unroll 1
            for (uint32_t t = 3u; t < sk_n; t++) red_add(dp + t, 1u);
        }
        const uint32_t rpA = M0.start;
        const bool want_issue = j + (uint32_t)(kFastStages - 1) < nblk;

        for (;;) {
            // ---- push the pending pieces that fit the window: rel + n <= kWin with rel = pos - win_lo as unsigned
            const uint32_t relA = rpA - win_lo, relB = ppB - win_lo;
            const bool fitA = nA != 0u && relA <= kWin - nA;
            const bool fitB = nB != 0u && relB <= kWin - nB;
            const uint32_t mA = __ballot_sync(kFull, fitA), mB = __ballot_sync(kFull, fitB);
            if (mA | mB) {
                __syncwarp();                                                // earlier ring reads are done
                const uint32_t at = ring_tail + __popc(mA & lt_mask) + __popc(mB & lt_mask);
                if (fitA) {
                    push_entry(at, relA, nA, qb);
                    nA = 0u;
                }
                if (fitB) {
                    push_entry(at + (fitA ? 1u : 0u), relB, nB, pqB);
                    nB = 0u;
                }
                ring_tail += __popc(mA) + __popc(mB);
                __syncwarp();
            }
            const bool left = __any_sync(kFull, (nA | nB) != 0u);            // someone waits for a window move
            // ---- the one trip site and the one flush site
            for (;;) {
                const uint32_t avail = ring_tail - ring_head;
                if (avail < (uint32_t)Q || cnt == kCntMax) {                 // rare: everything but a plain trip
                    if (avail < (uint32_t)Q) {
                        const bool drain = left || last || (want_issue && (int)(ring_head - mark) < 0);
                        if (avail != 0u && drain) {                          // pad the ring with empty pieces to a full trip
                            if ((uint32_t)lane < (uint32_t)Q - avail)
                                sts128(ringb + 16u * ((ring_tail + lane) & (kFastRing - 1u)), make_uint4(0u, 0u, 0u, seqb));
                            ring_tail += (uint32_t)Q - avail;
                            __syncwarp();
                            continue;
                        }
                        if (!(cnt != 0u && (left || last))) break;
                    }
                    flush_counters<G>(pl, pa, pb, cnt, frow, plane0 + win_lo, cv.stride, lane);
                    cnt = 0u;
                    continue;
                }
                // -- trip: four pieces per read slot, straight-line.  ring_head is a multiple of Q and Q divides
                //    the ring, so the Q entries of a trip never wrap.
                const uint32_t ea = trip_ringb + 16u * (ring_head & (kFastRing - 1u));
                uint4 e[4];
#pragma unroll
                for (int q = 0; q < 4; q++) e[q] = lds128(ea + 16u * (uint32_t)(q * S));
                ring_head += (uint32_t)Q;
                uint32_t x[4][kW][kNC];
#pragma unroll
                for (int q = 0; q < 4; q++) piece_words<HAS_OK>(e[q], x[q], L0, lutb, lane_seq_off, seqb, okb);
                csa_trip(x, pl, pa, pb, cnt, spb);
                cnt += 4u;
            }
            if (!left) break;
            // move the window to the lowest pending piece (the counters were flushed above)
            win_lo = __reduce_min_sync(kFull, min(nA ? rpA : 0xFFFFFFFFu, nB ? ppB : 0xFFFFFFFFu)) & ~31u;
        }
        if (want_issue) issue_block(kFastStages > 2 ? M2 : M1, st == 0u ? (uint32_t)(kFastStages - 1) : st - 1u);

        // ---- rotate the pipelines
        mark = ring_tail;
        st = (st == (uint32_t)kFastStages - 1u) ? 0u : st + 1u;
        // (the values loaded at the top of this block are first touched HERE, by instructions the compiler cannot
        //  hoist: left to itself it copies them right behind the loads and every block waits out the HBM latency)
        M0 = M1;
        M1.start = opaque(M2.start);
        M1.cb = opaque(M2.cb);
        M1.ce = opaque(M2.ce);
        M1.wb = opaque(M2.wb);
        M1.we = opaque(M2.we);
        C0.c0 = opaque(C1.c0);
        C0.c1 = opaque(C1.c1);
        C0.c2 = opaque(C1.c2);
    }
    emit_deferred();
}

}  // namespace bc
