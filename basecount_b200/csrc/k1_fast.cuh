// k1_fast.cuh -- K1, the lean counting kernel: CIGAR walk + per-position base counting on sm_100a for the
// reads a straight-line decode can take (M, M-I-M, M-D-M once clips are dropped and =/X are spelled M, which the
// packers do -- practically every short read).  Replaces the loop nest of the reference operator
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
//     staging, no staged-range bookkeeping.  (Loading them by POSITION at the same time as the offsets and
//     handing each lane its own through shared memory removes the dependent load and the register rotation --
//     60 fewer instructions per block -- and was measured SLOWER: the exchange sits on the block's critical
//     path, the rotation's moves do not.  profiles/r4_e_k1_times.txt);
//   * sequence words are staged by one 1-D TMA bulk copy per block (two with a quality mask) into a ring of
//     stages; the range a stage holds is recomputed from the metadata instead of being parked in shared memory;
//   * the decode knows two shapes only, "M" and "M, I or D, M", in the packers' CIGAR normal form (cigar_canon.h);
//   * the ring holds a whole block's pieces plus a partial trip, so pushes never wait for ring space, and an
//     invalid window is a window position no read can fit (no validity flag).
#pragma once
#include "k1_count.cuh"

namespace bc {

#ifndef BC_K1F_MINCTAS
#define BC_K1F_MINCTAS 3
#endif
// Register cap of the kernel (0: what three CTAs per SM allow, 168).  160 leaves 4,096 registers of an SM free beside
// its three CTAs; a summarise kernel of 64-thread CTAs that fits there was measured (profiles/r4_g_step_ab.txt) and
// lost: two warps per SM of dependent FP64 chains take 250 us.
#ifndef BC_K1F_REGS
#define BC_K1F_REGS 0
#endif
#if BC_K1F_REGS
#define BC_K1F_BOUNDS __maxnreg__(BC_K1F_REGS)                      // (cannot be combined with __launch_bounds__)
#else
#define BC_K1F_BOUNDS __launch_bounds__(kK1Threads, BC_K1F_MINCTAS)
#endif
// A/B switches of tools/ab_variants.py (the defaults are the measured winners)
#ifndef BC_K1F_MASK3
#define BC_K1F_MASK3 1          // 1: lo&m, hi&m, (lo&m)&(hi&m) with m folded in (3 LOP3 per word); 0: m first, then 3 parallel ANDs
#endif
#ifndef BC_K1F_CSA_BRANCH
#define BC_K1F_CSA_BRANCH 1     // 1: a parked carry is computed in the branch that parks it; 0: ahead of the branch
#endif
#ifndef BC_K1F_NL0_SMEM
#define BC_K1F_NL0_SMEM 1       // 1: the lane's window offset comes back from shared memory (one VIADDMNMX per bound)
#endif
constexpr uint32_t kFastSeqBuf = 3u * kSeqCap;   // staged 64-bit plane words per warp: 3 stages of kSeqCap words, or 4 of 3/4 kSeqCap
constexpr uint32_t kFastStageShort = kFastSeqBuf / 4u;
constexpr int kFastMaxStages = 4;
constexpr uint32_t kFastRing = 128;            // ring entries: <= 64 pieces of a block + a partial trip (< 32)
constexpr uint32_t kFastRpbMax = 32;
constexpr uint32_t kNoWindow = 0x80000000u;    // a window position no piece fits (reference positions are < 2^31)

template <int G, bool HAS_OK>
struct K1FastCfg {
    static constexpr int S = 32 / G;
    static constexpr int Q = 4 * S;                         // ring entries per half trip (four pieces per read slot)
    static constexpr int Q2 = 2 * Q;                        // per trip
    // Counted quantities per window word.  Without a quality mask the number of pieces covering a column is an
    // interval count: it comes from a difference array in shared memory (+1 at a piece's first column, -1 behind
    // its last, two shared-memory atomics per piece when it is pushed, one prefix sum per flush) instead of a
    // fourth bit-sliced counter.  With a mask, "valid" is per base and stays bit-sliced.
    static constexpr int NC = HAS_OK ? 4 : 3;
    static constexpr int NB = HAS_OK ? 7 : 8;               // bit planes per counter, all in registers
    static constexpr uint32_t kCntMax = (1u << NB) - 8u;    // pieces per slot between flushes (a multiple of 8)
    static constexpr uint32_t kWin = 32u * kW * G;
    static constexpr uint32_t kMaxFit = kWin - 31u;
    static constexpr uint32_t kCols = kFlushStride * kW * G;
    // per warp: ring | sequence stages | quality-mask stages | flush rows | difference array | mbarriers.  A piece's
    // unclamped word index reaches up to 64*G columns before or after its data: the ring in front and the flush
    // rows behind keep those (masked-away) reads inside the CTA's shared memory.
    static constexpr uint32_t ring_off = 0;
    static constexpr uint32_t seq_off = ring_off + kFastRing * 16u;
    static constexpr uint32_t ok_off = seq_off + kFastSeqBuf * 8u;
    static constexpr uint32_t frow_off = ok_off + (HAS_OK ? kFastSeqBuf * 4u : 0u);
    static constexpr uint32_t frow_bytes = (uint32_t)NC * kCols * 2u;
    static constexpr uint32_t cov_off = frow_off + frow_bytes;
    static constexpr uint32_t cov_bytes = HAS_OK ? 0u : (kWin + 4u) * 4u;
    static constexpr uint32_t bar_off = cov_off + cov_bytes;
    static constexpr uint32_t warp_bytes = bar_off + 32u;
    static constexpr uint32_t cta_bytes = kK1WarpsPerCta * warp_bytes;      // + the statically allocated mask table
    static_assert(seq_off % 16 == 0 && ok_off % 16 == 0 && frow_off % 16 == 0 && cov_off % 16 == 0 && bar_off % 16 == 0 &&
                      warp_bytes % 16 == 0,
                  "TMA destinations are 16-byte aligned");
    static_assert(seq_off >= (kWin / 32 + 4) * 8 && frow_bytes >= (kWin / 32 + 4) * 8, "guard bands around the stages");
    static_assert(kFastRing % Q2 == 0 && kFastRing >= 64 + Q2, "a block's pieces and a partial trip fit the ring");
    static_assert(kFastStageShort % 4 == 0, "stages start on 32-byte boundaries");
};
template <int G, bool HAS_OK>
__host__ __device__ constexpr uint32_t k1_fast_cta_smem_bytes()
{
    return K1FastCfg<G, HAS_OK>::cta_bytes;
}

__device__ __forceinline__ uint32_t ldg_if(const uint32_t *p, bool on)
{
    uint32_t v = 0u;
    asm volatile("{\n.reg .pred P1;\nsetp.ne.u32 P1, %2, 0;\n@P1 ld.global.nc.u32 %0, [%1];\n}" : "+r"(v) : "l"(p), "r"((uint32_t)on));
    return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t a, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// Masked words of one ring entry (a piece) for a lane's two window words: lo, hi, lo & hi (and the mask itself when
// "valid" is counted bit-sliced).  Entry: x / y = 8 * first / end column of the piece relative to the window (the
// byte offset of its mask in the table), z = bit index of window column 0 in the staged data, w = shared address
// of the plane word that holds window column 0.  nL0_8 = -8 * (the lane's first window column); lut = the mask
// table's shared address, a link-time constant that ends up as the immediate of the LDS.
template <bool HAS_OK, int NC>
__device__ __forceinline__ void fast_piece(const uint4 e, uint32_t (&x)[kW][NC], int nL0_8, uint32_t lut, uint32_t lane_seq_off,
                                           uint32_t seqb, uint32_t okb)
{
    const uint32_t a_c = (uint32_t)__viaddmin_s32_relu((int)e.x, nL0_8, 512);   // 8 * clamp(first - L0, 0, 64)
    const uint32_t e_c = (uint32_t)__viaddmin_s32_relu((int)e.y, nL0_8, 512);
    const uint2 ga = lds64(lut + a_c), ge = lds64(lut + e_c);
    const uint32_t g1[kW] = {ga.x, ga.y}, g0[kW] = {ge.x, ge.y};                 // columns at or above first / end
    const uint32_t wa = e.w + lane_seq_off;
    const uint2 r0 = lds64(wa), r1 = lds64(wa + 8u), r2 = lds64(wa + 16u);
    const uint32_t lo[kW] = {__funnelshift_r(r0.x, r1.x, e.z), __funnelshift_r(r1.x, r2.x, e.z)};
    const uint32_t hi[kW] = {__funnelshift_r(r0.y, r1.y, e.z), __funnelshift_r(r1.y, r2.y, e.z)};
    if (HAS_OK) {
        const uint32_t oa = okb + (uint32_t)((int)(wa - seqb) >> 1);
        const uint32_t o0 = lds32(oa), o1 = lds32(oa + 4u), o2 = lds32(oa + 8u);
        const uint32_t ok[kW] = {__funnelshift_r(o0, o1, e.z), __funnelshift_r(o1, o2, e.z)};
#pragma unroll
        for (int w = 0; w < kW; w++) {
            const uint32_t m = ok[w] & g1[w] & ~g0[w];
            x[w][0] = lo[w] & m;
            x[w][1] = hi[w] & m;
            x[w][2] = x[w][0] & x[w][1];
            if (NC > 3) x[w][3] = m;
        }
    } else {
#pragma unroll
        for (int w = 0; w < kW; w++) {
#if BC_K1F_MASK3
            x[w][0] = lo[w] & g1[w] & ~g0[w];                                     // three LOP3 per window word
            x[w][1] = hi[w] & g1[w] & ~g0[w];
            x[w][2] = x[w][0] & x[w][1];
#else
            const uint32_t m = g1[w] & ~g0[w];
            x[w][0] = lo[w] & m;
            x[w][1] = hi[w] & m;
            x[w][2] = lo[w] & hi[w] & m;
#endif
        }
    }
}

// Four pieces per read slot into carry-save levels 0 and 1; the weight-4 carry comes back in c2.
template <int NC, int NB>
__device__ __forceinline__ void csa_half(const uint32_t (&x)[4][kW][NC], uint32_t (&pl)[kW][NC][NB], uint32_t (&c2)[kW][NC])
{
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const uint32_t c1a = maj3(pl[w][k][0], x[0][w][k], x[1][w][k]);
            const uint32_t t = pl[w][k][0] ^ x[0][w][k] ^ x[1][w][k];
            const uint32_t c1b = maj3(t, x[2][w][k], x[3][w][k]);
            pl[w][k][0] = t ^ x[2][w][k] ^ x[3][w][k];
            c2[w][k] = maj3(pl[w][k][1], c1a, c1b);
            pl[w][k][1] ^= c1a ^ c1b;
        }
    }
}

// majority / parity of three words as ONE instruction each that the compiler neither merges with an identical
// expression in another branch nor moves out of its branch (see csa_upper)
__device__ __forceinline__ uint32_t maj3_here(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// The upper levels of one trip (eight pieces per read slot): the two weight-4 carries of its halves meet in plane 2
// with no pending register; the weight-8 carry is parked every other trip (pb), the weight-16 carry every fourth
// (pc), and the weight-32 carry ripples through planes 5.. every eighth.  cnt = pieces per slot counted before
// this trip (a multiple of 8).  A carry that is parked is COMPUTED in the branch that parks it, straight into the
// pending register: computed ahead of the branch it cost a register move per counter in every trip.
template <int NC, int NB>
__device__ __forceinline__ void csa_upper(const uint32_t (&c2a)[kW][NC], const uint32_t (&c2b)[kW][NC], uint32_t (&pl)[kW][NC][NB],
                                          uint32_t (&pb)[kW][NC], uint32_t (&pc)[kW][NC], uint32_t cnt)
{
#if !BC_K1F_CSA_BRANCH
#define maj3_here maj3
#endif
    if (!(cnt & 8u)) {
#pragma unroll
        for (int w = 0; w < kW; w++)
#pragma unroll
            for (int k = 0; k < NC; k++) {
                pb[w][k] = maj3_here(pl[w][k][2], c2a[w][k], c2b[w][k]);
                pl[w][k][2] ^= c2a[w][k] ^ c2b[w][k];
            }
        return;
    }
    uint32_t c3[kW][NC];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
            c3[w][k] = maj3(pl[w][k][2], c2a[w][k], c2b[w][k]);
            pl[w][k][2] ^= c2a[w][k] ^ c2b[w][k];
        }
    }
    if (!(cnt & 16u)) {
#pragma unroll
        for (int w = 0; w < kW; w++)
#pragma unroll
            for (int k = 0; k < NC; k++) {
                pc[w][k] = maj3_here(pl[w][k][3], pb[w][k], c3[w][k]);
                pl[w][k][3] ^= pb[w][k] ^ c3[w][k];
            }
        return;
    }
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const uint32_t c4 = maj3(pl[w][k][3], pb[w][k], c3[w][k]);
            pl[w][k][3] ^= pb[w][k] ^ c3[w][k];
            uint32_t c = maj3(pl[w][k][4], pc[w][k], c4);
            pl[w][k][4] ^= pc[w][k] ^ c4;
#pragma unroll
            for (int p = 5; p < NB; p++) {        // ripple the weight-32 carry upwards
                const uint32_t t = pl[w][k][p] & c;
                pl[w][k][p] ^= c;
                c = t;
            }
        }
    }
#if !BC_K1F_CSA_BRANCH
#undef maj3_here
#endif
}

// Convert the warp's vertical counters to integers and add them to the HBM planes (the flush of k1_count.cuh with
// every plane in registers, the pending carries of csa_upper, and -- without a quality mask -- the coverage taken
// from the difference array at shared address covb: an in-place prefix sum, read, then zeroed for the next window).
//   frow: NC rows x (kW*G window words x kFlushStride) uint16
// PIN = planes that can be non-zero in a lane's counters (cnt < 2^PIN pieces went in): the slot combine and the
// extraction only touch those (and the R carry planes the combine adds), which is most of a flush after a short
// window -- whole-genome coverage flushes every ~70 reads.
template <int G, int NC, int NB, int PIN>
__device__ __forceinline__ void flush_fast_body(uint32_t (&pl)[kW][NC][NB], uint32_t (&pb)[kW][NC], uint32_t (&pc)[kW][NC],
                                                uint32_t cnt, uint16_t *frow, uint32_t covb, uint32_t *__restrict__ counts,
                                                uint64_t stride, int lane)
{
    constexpr int S = 32 / G;
    constexpr int R = S == 8 ? 3 : (S == 4 ? 2 : (S == 2 ? 1 : 0));      // combine rounds
    constexpr int kCols = (int)kFlushStride * kW * G;
    constexpr int kQ = 4;                                                 // counter slots per window word (NC real ones)
    constexpr int kN = kW * kQ;                                           // counter slots per lane (8)
    constexpr int P = PIN + R;                                            // planes after the combine
    const int slot = lane / G, wl = lane % G;
    uint32_t A[kN][P];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < kQ; k++) {
            const int i = w * kQ + k;
            if (k < NC) {
#pragma unroll
                for (int p = 0; p < PIN; p++) A[i][p] = pl[w][k][p];
                // pending carries of weight 8 / 16 are live iff that bit of cnt is set
                if (PIN > 3) {
                    uint32_t c = (cnt & 8u) ? pb[w][k] : 0u;
                    {
                        const uint32_t t = A[i][3] & c;
                        A[i][3] ^= c;
                        c = t;
                    }
                    if (PIN > 4) {
                        const uint32_t d = (cnt & 16u) ? pc[w][k] : 0u;  // two carries into plane 4: full adder
                        const uint32_t t = maj3(A[i][4], c, d);
                        A[i][4] ^= c ^ d;
                        c = t;
                    }
#pragma unroll
                    for (int p = 5; p < PIN; p++) {
                        const uint32_t t = A[i][p] & c;
                        A[i][p] ^= c;
                        c = t;
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < PIN; p++) A[i][p] = 0u;
            }
#pragma unroll
            for (int p = PIN; p < P; p++) A[i][p] = 0u;
        }
    }
    // ---- coverage from the difference array: lane l owns entries [l * 2G, (l + 1) * 2G)
    if (NC == 3) {
        constexpr int kPer = 2 * G;                   // entries per lane (a multiple of 4)
        const uint32_t mine = covb + 4u * (uint32_t)(lane * kPer);
        uint32_t tot = 0u;
#pragma unroll 1
        for (int i = 0; i < kPer; i += 4) {
            const uint4 v = lds128(mine + 4u * (uint32_t)i);
            tot += v.x + v.y + v.z + v.w;
        }
        uint32_t run = tot;                           // inclusive scan of the lane totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(kFull, run, d);
            if (lane >= d) run += o;
        }
        run -= tot;                                   // exclusive: what lies before this lane's entries
#pragma unroll 1
        for (int i = 0; i < kPer; i += 4) {
            uint4 v = lds128(mine + 4u * (uint32_t)i);
            v.x += run;
            v.y += v.x;
            v.z += v.y;
            v.w += v.z;
            run = v.w;
            sts128(mine + 4u * (uint32_t)i, v);
        }
    }
    // ---- bit-sliced sum over the read slots: each round a lane keeps half of its counter slots and adds the
    //      partner's copies of them (ripple-carry over the planes, one more plane per round)
    int first = 0;                                    // counter index (w * kQ + k) of A[0] after the rounds
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int d = G << r;
        const int h = kN >> (r + 1);
        const bool up = (slot >> r) & 1;
        if (up) first += h;
#pragma unroll
        for (int i = 0; i < h; i++) {
            if (NC == 3 && r == 0 && i == kQ - 1) continue;           // slot 3 of either window word is empty
            uint32_t carry = 0u;
#pragma unroll
            for (int p = 0; p < PIN + r; p++) {
                const uint32_t mine = up ? A[h + i][p] : A[i][p];
                const uint32_t give = up ? A[i][p] : A[h + i][p];
                const uint32_t got = __shfl_xor_sync(kFull, give, d);
                A[i][p] = mine ^ got ^ carry;
                carry = maj3(mine, got, carry);
            }
            A[i][PIN + r] = carry;
        }
    }
    // ---- extraction: byte t of acc = count of column jj + 8t (its low 8 bits; acc2 holds bits 8..), one shift and
    //      one (x & mask) | acc per plane with every shift amount an immediate
    constexpr int kLeft = kN >> R;                    // counter slots this lane extracts
#pragma unroll
    for (int i = 0; i < kLeft; i++) {
        const int ci = first + i, w = ci / kQ, k = ci % kQ;
        uint16_t *const dst = frow + k * kCols + (kW * wl + w) * (int)kFlushStride;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            uint32_t acc = 0u, acc2 = 0u;
#pragma unroll
            for (int p = 0; p < P && p < 8; p++) {
                const uint32_t t = p <= jj ? A[i][p] >> (jj - p) : A[i][p] << (p - jj);
                acc |= t & (0x01010101u << p);
            }
#pragma unroll
            for (int p = 8; p < P; p++) {
                const uint32_t t = (p - 8) <= jj ? A[i][p] >> (jj - (p - 8)) : A[i][p] << ((p - 8) - jj);
                acc2 |= t & (0x01010101u << (p - 8));
            }
            const uint32_t ev = P > 8 ? __byte_perm(acc, acc2, 0x6240) : (acc & 0x00FF00FFu);             // columns jj, jj+16
            const uint32_t od = P > 8 ? __byte_perm(acc, acc2, 0x7351) : ((acc >> 8) & 0x00FF00FFu);      // columns jj+8, jj+24
            if (k < NC) {                             // (slot 3 of a window word is empty without a quality mask)
                dst[jj] = (uint16_t)ev;
                dst[jj + 16] = (uint16_t)(ev >> 16);
                dst[jj + 8] = (uint16_t)od;
                dst[jj + 24] = (uint16_t)(od >> 16);
            }
        }
    }
    __syncwarp();
    uint32_t *const pA = counts + lane, *const pC = pA + stride, *const pG = pC + stride, *const pT = pG + stride;
#pragma unroll
    for (int w = 0; w < kW * G; w++) {
        const int at = (int)kFlushStride * w + lane;
        const uint32_t nlo = frow[at], nhi = frow[kCols + at], nb = frow[2 * kCols + at];
        const uint32_t nv = NC == 3 ? lds32(covb + 4u * (uint32_t)(32 * w + lane)) : (uint32_t)frow[3 * kCols + at];
        red_add(pA + 32 * w, nv + nb - nlo - nhi);           // RED.ADD, 128 B per warp instruction
        red_add(pC + 32 * w, nlo - nb);
        red_add(pG + 32 * w, nhi - nb);
        red_add(pT + 32 * w, nb);
    }
    __syncwarp();
    if (NC == 3) {                                    // a fresh difference array for the next window / next pieces
#pragma unroll
        for (int w = 0; w < kW * G; w++) sts32(covb + 4u * (uint32_t)(32 * w + lane), 0u);
        if (lane < 4) sts32(covb + 4u * (uint32_t)(32 * kW * G + lane), 0u);
    }
    __syncwarp();
}

template <int G, int NC, int NB>
__device__ __forceinline__ void flush_fast(uint32_t (&pl)[kW][NC][NB], uint32_t (&pb)[kW][NC], uint32_t (&pc)[kW][NC],
                                           uint32_t cnt, uint16_t *frow, uint32_t covb, uint32_t *__restrict__ counts,
                                           uint64_t stride, int lane)
{
    if (cnt < 32u) flush_fast_body<G, NC, NB, 5>(pl, pb, pc, cnt, frow, covb, counts, stride, lane);
    else if (NB > 7 && cnt < 128u) flush_fast_body<G, NC, NB, 7>(pl, pb, pc, cnt, frow, covb, counts, stride, lane);
    else flush_fast_body<G, NC, NB, NB>(pl, pb, pc, cnt, frow, covb, counts, stride, lane);
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
#pragma unroll
            for (int p = 0; p < NB; p++) pl[w][k][p] = 0u;
            pb[w][k] = 0u;
            pc[w][k] = 0u;
        }
    }
}

template <int G, bool HAS_OK>
__global__ void BC_K1F_BOUNDS
k1_count_fast(BatchView bv, CountView cv, const Chunk *__restrict__ chunks, uint32_t n_chunks, uint32_t rpb, uint32_t nst,
              Chunk *__restrict__ deferred, uint32_t *__restrict__ n_deferred)
{
    using C = K1FastCfg<G, HAS_OK>;
    constexpr int S = C::S, Q = C::Q, Q2 = C::Q2, NC = C::NC, NB = C::NB;
    constexpr uint32_t kWin = C::kWin;
    extern __shared__ __align__(128) unsigned char k1_smem[];
    // lut[v] = the 64 window columns of a lane at or above column v (v in [0, 64]).  Statically allocated: its shared
    // address is a constant, so a mask lookup is one LDS [8 * v + constant].
    __shared__ uint2 k1f_lut[66];

    uint32_t lane_u;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
    const int lane = (int)lane_u;
    const int warp_in_cta = threadIdx.x >> 5;

    for (int v = threadIdx.x; v <= 64; v += kK1Threads)
        k1f_lut[v] = make_uint2(v < 32 ? 0xFFFFFFFFu << v : 0u, v <= 32 ? 0xFFFFFFFFu : (v < 64 ? 0xFFFFFFFFu << (v - 32) : 0u));
    __syncthreads();                                // the only CTA-wide barrier: warps are independent from here on

    const uint32_t warp_id = blockIdx.x * kK1WarpsPerCta + warp_in_cta;
    if (warp_id >= n_chunks) return;

    unsigned char *wsm = k1_smem + (size_t)warp_in_cta * C::warp_bytes;
    uint16_t *frow = reinterpret_cast<uint16_t *>(wsm + C::frow_off);          // flush only
    const uint32_t lut = smem_u32(k1f_lut);
    const uint32_t wb = opaque(smem_u32(wsm));                                 // the warp's region
    const uint32_t ringb = wb + C::ring_off, seqb = wb + C::seq_off, okb = wb + C::ok_off, barb = wb + C::bar_off,
                   covb = wb + C::cov_off, xchb = wb + C::frow_off;

    const int slot = lane / G, wl = lane % G;
    // -8 * (window column of this lane's bit 0), through shared memory: a value ptxas cannot re-derive from the lane
    // number, so the clamp in fast_piece is ONE add-min-relu instead of a multiply-add and a min-relu per bound
#if BC_K1F_NL0_SMEM
    sts32(xchb + 4u * (uint32_t)lane, (uint32_t)(-8 * 32 * kW * wl));
    int nL0_8;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(nL0_8) : "r"(xchb + 4u * (uint32_t)lane) : "memory");
#else
    const int nL0_8 = -8 * 32 * kW * wl;
#endif
    const uint32_t lt_mask = opaque((1u << lane) - 1u);
    const uint32_t trip_ringb = opaque(ringb + 16u * (uint32_t)slot);          // ring entry of this slot in a trip
    const uint32_t lane_seq_off = 8u * kW * (uint32_t)wl;                      // byte offset of this lane's window words

    const Chunk ch = chunks[warp_id];
    const uint32_t ref_len = ch.ref_len;
    const uint32_t rb = ch.read_begin, re = ch.read_end;
    if (re <= rb) return;
    const uint32_t nblk = (re - rb + rpb - 1) / rpb;
    uint32_t *const plane0 = cv.counts + ch.col_base;                       // plane A, column 0 of this slot
    uint32_t *const ds_plane = plane0 + (uint64_t)kPlaneDS * cv.stride;

    // nst = 3 stages of kSeqCap words, or 4 of kFastStageShort when a block of reads is that short: block j + 2 is
    // staged during block j into the stage block j + 2 - nst has left, and with four stages that tenant's pieces are
    // always counted by then (with three, a trip that takes more entries than a block pushes -- 64 at G = 4 -- had
    // to be padded every other block to free the stage).
    const uint32_t stage_words = nst == 3u ? kSeqCap : kFastStageShort;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kFastMaxStages; s++) mbar_init_s(barb + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NC == 3) {
        for (uint32_t i = (uint32_t)lane; i < kWin + 4u; i += 32u) sts32(covb + 4u * i, 0u);
    }
    __syncwarp();

    // ---- block metadata.  Lane l holds read (block begin + l); indices are clamped to the block's end, so a lane
    //      without a read holds an empty read (no CIGAR words, no sequence words) and lane 31's ends are the block's.
    struct Meta { uint32_t start, cb, ce, wb, we; };
    auto load_meta = [&](uint32_t blk) {
        const uint32_t b0 = rb + blk * rpb;                                  // (the host keeps n_reads < 2^32 - 256)
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
    // Stage the sequence words [t_lo, t_hi) of a block: [first word & ~3, last word rounded up) clipped to a stage.
    auto issue_block = [&](uint32_t t_lo, uint32_t t_hi, uint32_t stg) {
        if (lane == 0) {
            const uint32_t s_lo = t_lo & ~3u, s_n = min(((t_hi + 3u) & ~3u) - s_lo, stage_words);  // 32 B / 16 B aligned sources
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of this stage
            const uint32_t bar = barb + 8u * stg;
            mbar_expect_tx_s(bar, s_n * 8u + (HAS_OK ? s_n * 4u : 0u));
            if (s_n) {
                bulk_g2s_s(seqb + stg * (stage_words * 8u), bv.planes + s_lo, s_n * 8u, bar);
                if (HAS_OK) bulk_g2s_s(okb + stg * (stage_words * 4u), bv.okmask + s_lo, s_n * 4u, bar);
            }
        }
    };
    auto block_end_woff = [&](uint32_t blk) {                                // seq_woff at the end of block blk (one address per warp)
        return __ldg(bv.seq_woff + min(rb + (blk + 1u) * rpb, re));
    };

    Meta M0 = load_meta(0), M1 = load_meta(1);
    uint32_t t_next;                                // seq_woff at the end of the last block staged so far
    {
        const uint32_t t0 = __ldg(bv.seq_woff + rb), t1 = block_end_woff(0);
        issue_block(t0, t1, 0);
        t_next = t1;
        if (nblk > 1) {
            t_next = block_end_woff(1);
            issue_block(t1, t_next, 1);
        }
    }
    Cig C0 = load_cig(M0);

    uint32_t phases = 0;                            // bit s: parity to wait for on stage s
    uint32_t st = 0;                                // stage of the current block
    uint32_t ring_head = 0, ring_tail = 0;
    uint32_t mark1 = 0, mark2 = 0;                  // ring entries before mark1 / mark2 come from blocks before j / j - 1
    uint32_t win_lo = kNoWindow, cnt = 0;
    uint32_t def_begin = 0xFFFFFFFFu, def_end = 0xFFFFFFFFu;   // open run of deferred blocks (reads [begin, end))
    uint32_t pl[kW][NC][NB], pb[kW][NC], pc[kW][NC];
#pragma unroll
    for (int w = 0; w < kW; w++) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
#pragma unroll
            for (int p = 0; p < NB; p++) pl[w][k][p] = 0u;
            pb[w][k] = 0u;
            pc[w][k] = 0u;
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
    // column 0 in the staged data, w = shared address of the plane word that holds window column 0.  Without a
    // quality mask the piece also enters the coverage difference array here.
    auto push_entry = [&](uint32_t at, uint32_t rel, uint32_t n, int qbit) {
        const int z = qbit - (int)rel;
        sts128(ringb + 16u * (at & (kFastRing - 1u)), make_uint4(8u * rel, 8u * (rel + n), (uint32_t)z, seqb + (uint32_t)((z >> 5) * 8)));
        if (NC == 3) {
            red_shared_add(covb + 4u * rel, 1u);
            red_shared_add(covb + 4u * (rel + n), 0xFFFFFFFFu);
        }
    };

    // j == nblk is a virtual empty block: it drains the ring and does the final flush in the one trip / flush site.
    // Stage schedule: block j lives in stage j % nst; block j + 2 is staged during block j, as soon as the pieces of
    // block j + 2 - nst (the previous tenant of that stage) have all been counted.
    for (uint32_t j = 0; j <= nblk; j++) {
        const bool last = (j == nblk);
        const Meta M2 = load_meta(j + 2u);                                   // in flight during this block
        const Cig C1 = load_cig(M1);
        const uint32_t t_lo = t_next;
        t_next = block_end_woff(j + 2u);                                     // (clamped to the chunk's end: always loadable)
        if (!last) {
            mbar_wait_s(barb + 8u * st, (phases >> st) & 1u);
            phases ^= 1u << st;
        }

        // ---- straight-line decode (count.cpp:35-96) of the two shapes a short read has in the packers' normal
        //      form (csrc/cigar_canon.h): "M" and "M, I or D, M".  Piece A is the first match run, piece B the
        //      second; a missing op reads as a zero word.  Any other CIGAR sends the block to the general walker.
        const uint32_t s_lo = __shfl_sync(kFull, M0.wb, 0) & ~3u;
        const int qb = (int)(st * (stage_words * 32u) + (M0.wb - s_lo) * 32u);   // bit index of the read's first base
        uint32_t nA, nB, ppB, sk_n, sk_pos;
        int pqB;
        bool bad;
        {
            const uint32_t ncig = M0.ce - M0.cb;
            const uint32_t len1 = C0.c1 >> 4, op1 = C0.c1 & 15u;
            nA = C0.c0 >> 4;
            nB = C0.c2 >> 4;
            const bool ins = op1 == 1u;                                       // count.cpp:74; otherwise D (count.cpp:80-87)
            sk_n = ins ? 0u : len1;
            sk_pos = M0.start + nA;
            ppB = sk_pos + sk_n;
            pqB = qb + (int)(nA + (ins ? len1 : 0u));
            const uint32_t qend = (uint32_t)qb + (M0.we - M0.wb) * 32u;      // end of the staged data of this read
            bad = ncig == 2u || ncig > 3u || ((C0.c0 | C0.c2) & 15u) != 0u || (ncig == 3u && (op1 - 1u) > 1u)   // shape
                  || (M0.we - s_lo) > stage_words                             // not (fully) staged
                  || ppB + nB > ref_len || (uint32_t)pqB + nB > qend           // reference end; CIGAR overruns the read
                  || sk_n > kLaneSkipMax || nA > C::kMaxFit || nB > C::kMaxFit;
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
        const bool want_issue = j + 2u < nblk;
        const uint32_t mark = nst == 3u ? mark1 : mark2;                     // the stage's last tenant ends here

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
                if (avail < (uint32_t)Q2 || cnt == C::kCntMax) {             // rare: everything but a plain trip
                    if (avail < (uint32_t)Q2) {
                        const bool drain = left || last || (want_issue && (int)(ring_head - mark) < 0);
                        if (avail != 0u && drain) {                          // pad the ring with empty pieces to a full trip
                            for (uint32_t i = (uint32_t)lane; i < (uint32_t)Q2 - avail; i += 32u)
                                sts128(ringb + 16u * ((ring_tail + i) & (kFastRing - 1u)), make_uint4(0u, 0u, 0u, seqb));
                            ring_tail += (uint32_t)Q2 - avail;
                            __syncwarp();
                            continue;
                        }
                        if (!(cnt != 0u && (left || last))) break;
                    }
                    flush_fast<G, NC, NB>(pl, pb, pc, cnt, frow, covb, plane0 + opaque(win_lo), cv.stride, lane);
                    cnt = 0u;
                    continue;
                }
                // -- trip: eight pieces per read slot in two halves, straight-line.  ring_head is a multiple of Q2 and
                //    Q2 divides the ring, so the entries of a trip never wrap.
                const uint32_t ea = trip_ringb + 16u * (ring_head & (kFastRing - 1u));
                ring_head += (uint32_t)Q2;
                uint32_t c2a[kW][NC], c2b[kW][NC];
                {
                    uint4 e[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) e[q] = lds128(ea + 16u * (uint32_t)(q * S));
                    uint32_t x[4][kW][NC];
#pragma unroll
                    for (int q = 0; q < 4; q++) fast_piece<HAS_OK, NC>(e[q], x[q], nL0_8, lut, lane_seq_off, seqb, okb);
                    csa_half<NC, NB>(x, pl, c2a);
                }
                {
                    uint4 e[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) e[q] = lds128(ea + 16u * (uint32_t)(Q + q * S));
                    uint32_t x[4][kW][NC];
#pragma unroll
                    for (int q = 0; q < 4; q++) fast_piece<HAS_OK, NC>(e[q], x[q], nL0_8, lut, lane_seq_off, seqb, okb);
                    csa_half<NC, NB>(x, pl, c2b);
                }
                csa_upper<NC, NB>(c2a, c2b, pl, pb, pc, cnt);
                cnt += 8u;
            }
            if (!left) break;
            // move the window to the lowest pending piece (the counters were flushed above)
            win_lo = __reduce_min_sync(kFull, min(nA ? rpA : 0xFFFFFFFFu, nB ? ppB : 0xFFFFFFFFu)) & ~31u;
        }
        if (want_issue) issue_block(t_lo, t_next, st + 2u >= nst ? st + 2u - nst : st + 2u);

        // ---- rotate the pipelines
        // (the values loaded at the top of this block are first touched HERE, by instructions the compiler cannot
        //  hoist: left to itself it copies them right behind the loads and every block waits out the HBM latency)
        mark2 = mark1;
        mark1 = ring_tail;
        st = (st + 1u == nst) ? 0u : st + 1u;
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
