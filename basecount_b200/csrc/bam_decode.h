// bam_decode.h -- native BGZF / BAM decode straight into the flat arrays the packer consumes.
//
// Replaces, on the host side of the path, what the reference does per read through pysam
// (basecount/main.py:97-99 open, :127 fetch(until_eof=True), :165 the read filter, :166-173 the
// four per-read attributes query_alignment_sequence / query_alignment_qualities /
// reference_start / cigartuples) -- SURVEY.md section 8(f) rank 1.  No per-read objects: the file
// is inflated block-parallel (BGZF blocks are independent deflate streams, SAM spec 4.1), the
// records are indexed once, and a selection (record range, reference id, minimum MAPQ) is copied
// into caller-owned arrays by a pool of threads: start, BAM-native CIGAR words, soft-clip-trimmed
// ASCII bases and phred bytes -- exactly a ReadBatch (basecount_b200/records.py).
//
// Host-only code (zlib + std::thread); compiled into the same C-ABI library as the kernels.
#pragma once
#include "cigar_canon.h"
#include <stdint.h>
#include <zlib.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "inflate_fast.h"
#endif
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace bcbam {

// CRC-32 of a BGZF block's payload (RFC 1952, the gzip polynomial, reflected).  zlib's table-driven crc32 runs at
// ~1.3 GB/s per thread, a quarter of what a block costs to open next to zlib's inflate; folding with carry-less
// multiplies (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ", the 4 x 128-bit
// fold zlib-ng / Chromium use) runs an order of magnitude faster.  Used when the CPU has PCLMULQDQ; zlib's
// crc32 finishes the tail and is the fallback (and the reference tests/test_bamio.py holds this to).
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_fold_clmul(const uint8_t *buf, size_t len, uint32_t crc)
{
    // len >= 64 and a multiple of 16; crc comes in and goes out in the register convention (bit-inverted)
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};     // x^(512+-32) mod P, fold by 4
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};     // x^(128+-32) mod P, fold by 1
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};     // x^64 mod P
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};     // P and the Barrett constant
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
    x3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
    x4 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k1k2));
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
        y6 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
        y7 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
        y8 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k3k4));       // four lanes -> one
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf));
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);                            // 128 -> 64 bits
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(k5k0));
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(poly));       // Barrett reduction to 32 bits
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

inline uint32_t crc32_fast(const uint8_t *p, size_t n)
{
    uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) {
        const size_t m = n & ~(size_t)15;
        crc = ~crc32_fold_clmul(p, m, ~crc);
        p += m;
        n -= m;
    }
#endif
    while (n) {                                           // (zlib takes a 32-bit length)
        const uInt step = (uInt)std::min<size_t>(n, 1u << 30);
        crc = (uint32_t)crc32(crc, p, step);
        p += step;
        n -= step;
    }
    return crc;
}

// One BGZF block: inflate `in` into dst[0, isize) and check the block's CRC-32.  The table-driven decoder of
// inflate_fast.h goes first (1.4-1.6 x zlib on BAM payloads); zlib decides whenever it declines or the CRC differs,
// so what counts as a valid block is exactly what zlib accepts.  `in` is followed by the block's 8-byte trailer.
// BASECOUNT_B200_INFLATE=zlib switches the first decoder off (A/B runs).
inline bool fast_inflate_enabled()
{
    static const bool on = [] {
        const char *e = std::getenv("BASECOUNT_B200_INFLATE");
        return !(e && std::strcmp(e, "zlib") == 0);
    }();
    return on;
}
inline bool inflate_block_checked(FastInflater *fi, z_stream *zs, const uint8_t *in, size_t in_len, uint8_t *dst,
                                  uint32_t isize, uint32_t crc)
{
    if (fi && fi->inflate(in, in_len, dst, isize) && crc32_fast(dst, isize) == crc) return true;
    inflateReset(zs);
    zs->next_in = const_cast<Bytef *>(in);
    zs->avail_in = (uInt)in_len;
    zs->next_out = dst;
    zs->avail_out = isize;
    const int rc = inflate(zs, Z_FINISH);
    return rc == Z_STREAM_END && zs->avail_out == 0 && crc32_fast(dst, isize) == crc;
}
// Byte buffers that are NOT zero-filled on resize: a 150 MB memset (and its page faults, on one thread) cost
// as much as the parallel inflate that overwrites every byte; the inflating threads touch the pages instead.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    // big buffers: 2 MB aligned and advised for transparent huge pages (512 x fewer first-touch faults)
    T *allocate(std::size_t n)
    {
        const std::size_t bytes = n * sizeof(T);
        if (bytes >= (8u << 20)) {
            void *p = nullptr;
            const std::size_t round = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
            if (::posix_memalign(&p, 2u << 20, round) != 0 || !p) throw std::bad_alloc();
            ::madvise(p, round, MADV_HUGEPAGE);
            return static_cast<T *>(p);
        }
        void *p = std::malloc(bytes ? bytes : 1);
        if (!p) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) noexcept { std::free(p); }
    template <class U, class... A>
    void construct(U *p, A &&...a)
    {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
using ByteVec = std::vector<uint8_t, NoInitAlloc<uint8_t>>;
}  // namespace bcbam

struct bc_bam {
    bcbam::ByteVec raw;                     // the inflated BAM stream
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lens;
    std::vector<uint64_t> rec_off;          // offset of every record's block_size field, plus the end
    int threads = 1;
    std::string err;
};

namespace bcbam {

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct Block { uint64_t c0, c1, u0; uint32_t isize, crc; };

template <class F>
inline void parallel_for(int threads, uint64_t n, uint64_t grain, F f)
{
    if (n == 0) return;
    const uint64_t chunks = (n + grain - 1) / grain;
    const int t = (int)std::min<uint64_t>((uint64_t)std::max(threads, 1), chunks);
    if (t <= 1) {
        f(0, n);
        return;
    }
    std::atomic<uint64_t> next(0);
    std::vector<std::thread> pool;
    for (int i = 0; i < t; i++)
        pool.emplace_back([&]() {
            for (;;) {
                const uint64_t c = next.fetch_add(1);
                if (c >= chunks) break;
                f(c * grain, std::min(n, (c + 1) * grain));
            }
        });
    for (auto &th : pool) th.join();
}

// Every BGZF block of the file (SAM spec 4.1: gzip member with a 'BC' extra subfield holding BSIZE-1).
inline bool scan_blocks(const uint8_t *d, uint64_t n, std::vector<Block> &out, uint64_t &total, std::string &err)
{
    uint64_t off = 0;
    total = 0;
    while (off < n) {
        if (n - off < 18 || d[off] != 31 || d[off + 1] != 139 || d[off + 2] != 8 || !(d[off + 3] & 4)) {
            err = "not a BGZF file (bad gzip member header)";
            return false;
        }
        const uint32_t xlen = rd16(&d[off + 10]);
        uint64_t p = off + 12;
        const uint64_t end = off + 12 + xlen;
        uint32_t bsize = 0;
        if (end > n) {
            err = "truncated BGZF block";
            return false;
        }
        while (p + 4 <= end) {
            const uint32_t slen = rd16(&d[p + 2]);
            if (d[p] == 66 && d[p + 1] == 67 && slen == 2 && p + 6 <= end) bsize = (uint32_t)rd16(&d[p + 4]) + 1u;
            p += 4 + slen;
        }
        if (bsize == 0 || off + bsize > n || bsize < 12 + xlen + 8) {
            err = "BGZF block without BC subfield or truncated";
            return false;
        }
        Block b;
        b.c0 = off + 12 + xlen;
        b.c1 = off + bsize - 8;
        b.crc = rd32(&d[off + bsize - 8]);
        b.isize = rd32(&d[off + bsize - 4]);
        b.u0 = total;
        total += b.isize;
        out.push_back(b);
        off += bsize;
    }
    return true;
}

// Soft-clipped bases at the left / right end of a record's CIGAR: what pysam's
// query_alignment_start / query_alignment_end trim (hard clips hold no bases; an optional H may sit
// outside the S).  Same rule as records._leading_trailing_clips.
inline void clips(const uint8_t *cig, uint32_t n_cigar, uint32_t &lead, uint32_t &trail)
{
    lead = trail = 0;
    if (n_cigar == 0) return;
    const uint32_t l = n_cigar - 1;
    uint32_t f2 = 0, l2 = l;
    if ((rd32(cig) & 15u) == 5u && 1u <= l) f2 = 1;                        // optional hard clip outside the soft clip
    const uint32_t wf = rd32(cig + 4 * f2);
    if ((wf & 15u) == 4u) lead = wf >> 4;
    if ((rd32(cig + 4 * l) & 15u) == 5u && l >= 1u) l2 = l - 1;
    const uint32_t wl = rd32(cig + 4 * l2);
    if ((wl & 15u) == 4u && !(l2 == f2 && (wf & 15u) == 4u)) trail = wl >> 4;   // a single S op is not counted twice
}

struct RecView {
    int32_t ref_id, pos;
    uint32_t l_read_name, mapq, n_cigar, flag, l_seq;
    const uint8_t *cig, *seq, *qual;
};

inline bool view(const bc_bam *b, uint64_t i, RecView &v)
{
    const uint64_t o = b->rec_off[i], e = b->rec_off[i + 1];
    if (e - o < 36) return false;
    const uint8_t *p = b->raw.data() + o;
    v.ref_id = (int32_t)rd32(p + 4);
    v.pos = (int32_t)rd32(p + 8);
    v.l_read_name = p[12];
    v.mapq = p[13];
    v.n_cigar = rd16(p + 16);
    v.flag = rd16(p + 18);
    v.l_seq = rd32(p + 20);
    const uint64_t need = 36ull + v.l_read_name + 4ull * v.n_cigar + (v.l_seq + 1ull) / 2 + v.l_seq;
    if (need > e - o) return false;
    v.cig = p + 36 + v.l_read_name;
    v.seq = v.cig + 4ull * v.n_cigar;
    v.qual = v.seq + (v.l_seq + 1ull) / 2;
    // A CIGAR of more than 65535 operations does not fit n_cigar_op: BAM stores the placeholder
    // "<l_seq>S<reference span>N" and the real CIGAR in the CG:B,I tag (SAM spec 4.2.2); htslib -- and so pysam's
    // cigartuples, basecount/main.py:173 -- hands out the real one.  A placeholder without the tag is malformed.
    if (v.n_cigar == 2 && (rd32(v.cig) & 15u) == 4u && (rd32(v.cig) >> 4) == v.l_seq && (rd32(v.cig + 4) & 15u) == 3u) {
        const uint8_t *a = v.qual + v.l_seq, *end = p + (e - o);
        bool found = false;
        while (a + 3 <= end) {
            const uint8_t t0 = a[0], t1 = a[1], ty = a[2];
            a += 3;
            uint64_t sz = 0;
            if (ty == 'A' || ty == 'c' || ty == 'C') sz = 1;
            else if (ty == 's' || ty == 'S') sz = 2;
            else if (ty == 'i' || ty == 'I' || ty == 'f') sz = 4;
            else if (ty == 'Z' || ty == 'H') {
                const uint8_t *z = a;
                while (z < end && *z) z++;
                if (z >= end) return false;
                sz = (uint64_t)(z - a) + 1;
            } else if (ty == 'B') {
                if (a + 5 > end) return false;
                const uint8_t sub = a[0];
                const uint64_t cnt = rd32(a + 1);
                const uint64_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
                if (es == 0) return false;
                if (t0 == 'C' && t1 == 'G' && sub == 'I') {
                    if (a + 5 + 4 * cnt > end || cnt > 0xFFFFFFFFull) return false;
                    v.cig = a + 5;
                    v.n_cigar = (uint32_t)cnt;
                    found = true;
                    break;
                }
                sz = 5 + es * cnt;
            } else {
                return false;
            }
            if (sz > (uint64_t)(end - a)) return false;
            a += sz;
        }
        (void)found;                       // (without the tag the record keeps its two operations, as in htslib)
    }
    return true;
}

inline bool keep(const RecView &v, int32_t ref_id, uint32_t min_mapq)
{
    return !(v.flag & 4u) && v.mapq >= min_mapq && v.ref_id == ref_id;      // basecount/main.py:165-166
}

// QUAL '*' is stored as 0xFF bytes: pysam then returns None for query_alignment_qualities and the reference's
// bcount raises TypeError for the whole chunk (the pybind11 cast at count.cpp:11), whatever min_base_quality is.
inline bool missing_qual(const RecView &v) { return v.l_seq > 0 && v.qual[0] == 0xFF; }

}  // namespace bcbam

// ---- API (exported with C linkage from bc_api.cu) ---------------------------------------------
inline int bc_bam_open_impl(const char *path, int threads, bc_bam **out, std::string &err)
{
    using namespace bcbam;
    // the compressed file is mapped, not copied: the inflating threads read it straight from the page cache
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) {
        err = std::string("cannot open ") + path;
        return 1;
    }
    struct stat st;
    if (::fstat(fd, &st) != 0 || st.st_size < 0) {
        ::close(fd);
        err = "cannot size the file";
        return 1;
    }
    const uint64_t fsize = (uint64_t)st.st_size;
    void *map = fsize ? ::mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    ::close(fd);
    if (fsize && map == MAP_FAILED) {
        err = "cannot map the file";
        return 1;
    }
    struct Unmap {
        void *p;
        uint64_t n;
        ~Unmap() { if (p && n) ::munmap(p, n); }
    } unmap{map, fsize};
    const uint8_t *file = static_cast<const uint8_t *>(map);

    bc_bam *b = new bc_bam();
    b->threads = threads > 0 ? threads : (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::vector<Block> blocks;
    uint64_t total = 0;
    if (!scan_blocks(file, fsize, blocks, total, err)) {
        delete b;
        return 2;
    }
    b->raw.resize(total);
    std::atomic<int> bad(0);
    parallel_for(b->threads, blocks.size(), 16, [&](uint64_t a, uint64_t e) {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = 1;
            return;
        }
        std::unique_ptr<FastInflater> fi(fast_inflate_enabled() ? new FastInflater() : nullptr);
        for (uint64_t i = a; i < e; i++) {
            const Block &k = blocks[i];
            if (k.isize == 0) continue;
            if (!inflate_block_checked(fi.get(), &zs, file + k.c0, k.c1 - k.c0, b->raw.data() + k.u0, k.isize, k.crc)) bad = 1;
        }
        inflateEnd(&zs);
    });
    if (bad) {
        err = "corrupt BGZF block (inflate or CRC failed)";
        delete b;
        return 2;
    }
    // header (SAM spec 4.2)
    const ByteVec &r = b->raw;
    if (r.size() < 12 || std::memcmp(r.data(), "BAM\1", 4) != 0) {
        err = "not a BAM file (bad magic)";
        delete b;
        return 2;
    }
    uint64_t p = 8ull + rd32(&r[4]);
    if (p + 4 > r.size()) {
        err = "truncated BAM header";
        delete b;
        return 2;
    }
    const uint32_t n_ref = rd32(&r[p]);
    p += 4;
    for (uint32_t i = 0; i < n_ref; i++) {
        if (p + 4 > r.size()) {
            err = "truncated BAM header";
            delete b;
            return 2;
        }
        const uint32_t l_name = rd32(&r[p]);
        if (p + 8ull + l_name > r.size() || l_name == 0) {
            err = "truncated BAM header";
            delete b;
            return 2;
        }
        b->ref_names.emplace_back((const char *)&r[p + 4], l_name - 1);
        b->ref_lens.push_back(rd32(&r[p + 4 + l_name]));
        p += 8ull + l_name;
    }
    // record index.  The walk is a chain of dependent loads (a record's size gives the next record's offset),
    // one cache miss each; records of one file are about the same size, so the lines a few records ahead are
    // prefetched on that guess.  The second pass (validation) is independent per record and runs in parallel.
    b->rec_off.reserve((size_t)(r.size() / 256 + 16));
    while (p + 4 <= r.size()) {
        b->rec_off.push_back(p);
        const uint64_t step = 4ull + rd32(&r[p]);
        if (p + 8 * step < r.size()) {
            __builtin_prefetch(&r[p + 4 * step]);
            __builtin_prefetch(&r[p + 8 * step]);
        }
        p += step;
    }
    if (p != r.size()) {
        err = "truncated BAM record";
        delete b;
        return 2;
    }
    b->rec_off.push_back(p);
    std::atomic<int> malformed(0);
    parallel_for(b->threads, b->rec_off.size() - 1, 1 << 13, [&](uint64_t a, uint64_t e) {
        RecView v;
        for (uint64_t i = a; i < e; i++)
            if (!view(b, i, v)) malformed = 1;
    });
    if (malformed) {
        err = "malformed BAM record";
        delete b;
        return 2;
    }
    *out = b;
    return 0;
}

// ---- a BAM file read span by span (bounded host memory) ----------------------------------------------------------
// bc_bam_open_impl holds the whole inflated file, which a whole-genome BAM (hundreds of GB inflated) does not
// allow.  A stream maps the file, scans the BGZF block headers once (no inflation), parses the header from the
// first blocks, and then hands out consecutive SPANS: an ordinary bc_bam holding the records that start in the next
// ~max_bytes of the inflated stream (whole records only; the record that straddles the end opens the next span).
// Everything that works on a bc_bam -- core, select, one-pass pack -- works on a span; the host counts span after
// span into the same accumulators (the counts do not depend on where the cuts fall, main.py:142).
struct bc_bam_stream {
    void *map = nullptr;
    uint64_t fsize = 0;
    std::vector<bcbam::Block> blocks;
    uint64_t total = 0;                     // inflated bytes of the whole file
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lens;
    uint64_t cur = 0;                       // inflated offset of the next record
    bool first = true;                      // the first span is handed out even when it holds no record
    int threads = 1;
    ~bc_bam_stream()
    {
        if (map && fsize) ::munmap(map, fsize);
    }
};

namespace bcbam {

// Inflate blocks [i0, i1) of the mapped file into dst (dst[0] = inflated offset blocks[i0].u0).
inline bool inflate_blocks(const uint8_t *file, const std::vector<Block> &blocks, uint64_t i0, uint64_t i1, uint8_t *dst, int threads)
{
    if (i1 <= i0) return true;
    const uint64_t base = blocks[i0].u0;
    std::atomic<int> bad(0);
    parallel_for(threads, i1 - i0, 16, [&](uint64_t a, uint64_t e) {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = 1;
            return;
        }
        std::unique_ptr<FastInflater> fi(fast_inflate_enabled() ? new FastInflater() : nullptr);
        for (uint64_t i = i0 + a; i < i0 + e; i++) {
            const Block &k = blocks[i];
            if (k.isize == 0) continue;
            if (!inflate_block_checked(fi.get(), &zs, file + k.c0, k.c1 - k.c0, dst + (k.u0 - base), k.isize, k.crc)) bad = 1;
        }
        inflateEnd(&zs);
    });
    return !bad;
}

// First block index whose inflated range ends beyond offset u (blocks.size() if none).
inline uint64_t block_of(const std::vector<Block> &blocks, uint64_t u)
{
    uint64_t lo = 0, hi = blocks.size();
    while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        if (blocks[mid].u0 + blocks[mid].isize <= u) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

}  // namespace bcbam

inline int bc_bam_stream_open_impl(const char *path, int threads, bc_bam_stream **out, std::string &err)
{
    using namespace bcbam;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) {
        err = std::string("cannot open ") + path;
        return 1;
    }
    struct stat st;
    if (::fstat(fd, &st) != 0 || st.st_size < 0) {
        ::close(fd);
        err = "cannot size the file";
        return 1;
    }
    std::unique_ptr<bc_bam_stream> s(new bc_bam_stream());
    s->fsize = (uint64_t)st.st_size;
    s->map = s->fsize ? ::mmap(nullptr, s->fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    ::close(fd);
    if (s->fsize && s->map == MAP_FAILED) {
        s->map = nullptr;
        err = "cannot map the file";
        return 1;
    }
    s->threads = threads > 0 ? threads : (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    const uint8_t *file = static_cast<const uint8_t *>(s->map);
    if (!scan_blocks(file, s->fsize, s->blocks, s->total, err)) return 2;
    // the header (SAM spec 4.2) from a prefix of the inflated stream, doubled until it holds all of it
    uint64_t want = 1u << 20;
    for (;;) {
        const uint64_t i1 = std::min<uint64_t>(block_of(s->blocks, std::min(want, s->total)) + 1, s->blocks.size());
        const uint64_t have = i1 ? s->blocks[i1 - 1].u0 + s->blocks[i1 - 1].isize : 0;
        ByteVec r;
        r.resize(have);
        if (!inflate_blocks(file, s->blocks, 0, i1, r.data(), s->threads)) {
            err = "corrupt BGZF block (inflate or CRC failed)";
            return 2;
        }
        const bool all = have >= s->total;
        auto retry = [&]() { return !all; };                 // truncated inside this prefix: look at a longer one
        if (r.size() < 12 || std::memcmp(r.data(), "BAM\1", 4) != 0) {
            if (r.size() < 12 && retry()) { want *= 2; continue; }
            err = "not a BAM file (bad magic)";
            return 2;
        }
        uint64_t p = 8ull + rd32(&r[4]);
        bool cut = p + 4 > r.size();
        uint32_t n_ref = cut ? 0u : rd32(&r[p]);
        p += 4;
        s->ref_names.clear();
        s->ref_lens.clear();
        for (uint32_t i = 0; i < n_ref && !cut; i++) {
            if (p + 4 > r.size()) { cut = true; break; }
            const uint32_t l_name = rd32(&r[p]);
            if (l_name == 0) {
                err = "truncated BAM header";
                return 2;
            }
            if (p + 8ull + l_name > r.size()) { cut = true; break; }
            s->ref_names.emplace_back((const char *)&r[p + 4], l_name - 1);
            s->ref_lens.push_back(rd32(&r[p + 4 + l_name]));
            p += 8ull + l_name;
        }
        if (cut) {
            if (retry()) { want *= 2; continue; }
            err = "truncated BAM header";
            return 2;
        }
        s->cur = p;
        break;
    }
    *out = s.release();
    return 0;
}

// The next span (see bc_bam_stream): *out = nullptr at the end of the file.
inline int bc_bam_stream_next_impl(bc_bam_stream *s, uint64_t max_bytes, bc_bam **out, std::string &err)
{
    using namespace bcbam;
    *out = nullptr;
    if (s->cur >= s->total && !s->first) return 0;
    const uint8_t *file = static_cast<const uint8_t *>(s->map);
    if (max_bytes < (1u << 16)) max_bytes = 1u << 16;
    for (;;) {
        std::unique_ptr<bc_bam> b(new bc_bam());
        b->threads = s->threads;
        b->ref_names = s->ref_names;
        b->ref_lens = s->ref_lens;
        const uint64_t i0 = block_of(s->blocks, s->cur);
        const uint64_t end_want = s->total - s->cur > max_bytes ? s->cur + max_bytes : s->total;
        const uint64_t i1 = end_want >= s->total ? s->blocks.size() : std::min<uint64_t>(block_of(s->blocks, end_want) + 1, s->blocks.size());
        const uint64_t base = i0 < s->blocks.size() ? s->blocks[i0].u0 : s->total;
        const uint64_t have = i1 > i0 ? s->blocks[i1 - 1].u0 + s->blocks[i1 - 1].isize - base : 0;
        b->raw.resize(have);
        if (!inflate_blocks(file, s->blocks, i0, i1, b->raw.data(), s->threads)) {
            err = "corrupt BGZF block (inflate or CRC failed)";
            return 2;
        }
        const ByteVec &r = b->raw;
        uint64_t p = s->cur - base;
        b->rec_off.reserve((size_t)(r.size() / 256 + 16));
        while (p + 4 <= r.size()) {
            const uint64_t step = 4ull + rd32(&r[p]);
            if (p + step > r.size()) break;                    // straddles the end of the span: opens the next one
            b->rec_off.push_back(p);
            if (p + 8 * step < r.size()) {
                __builtin_prefetch(&r[p + 4 * step]);
                __builtin_prefetch(&r[p + 8 * step]);
            }
            p += step;
        }
        const bool at_end = i1 >= s->blocks.size();
        if (b->rec_off.empty() && !at_end) {                  // one record longer than the span: look at a longer one
            max_bytes *= 2;
            continue;
        }
        if (at_end && p != r.size()) {
            err = "truncated BAM record";
            return 2;
        }
        b->rec_off.push_back(p);
        std::atomic<int> malformed(0);
        bc_bam *bp = b.get();
        parallel_for(b->threads, b->rec_off.size() - 1, 1 << 13, [&](uint64_t a, uint64_t e) {
            RecView v;
            for (uint64_t i = a; i < e; i++)
                if (!view(bp, i, v)) malformed = 1;
        });
        if (malformed) {
            err = "malformed BAM record";
            return 2;
        }
        s->cur = base + p;
        s->first = false;
        *out = b.release();
        return 0;
    }
}

// ref_id / pos / mapq / flag of every record (what count_alignments needs to place its chunk cuts).
inline void bc_bam_core_impl(const bc_bam *b, int32_t *ref_id, int32_t *pos, uint8_t *mapq, uint16_t *flag)
{
    using namespace bcbam;
    const uint64_t n = b->rec_off.size() - 1;
    parallel_for(b->threads, n, 1 << 14, [&](uint64_t a, uint64_t e) {
        RecView v;
        for (uint64_t i = a; i < e; i++) {
            view(b, i, v);
            if (ref_id) ref_id[i] = v.ref_id;
            if (pos) pos[i] = v.pos;
            if (mapq) mapq[i] = (uint8_t)v.mapq;
            if (flag) flag[i] = (uint16_t)v.flag;
        }
    });
}

// Sizes of the selection: kept reads, their CIGAR ops and their soft-clip-trimmed bases.
// Returns false if a kept read has no QUAL (the reference raises TypeError there).
inline bool bc_bam_select_sizes_impl(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                                     uint64_t *n_reads, uint64_t *n_cigar, uint64_t *n_bases)
{
    using namespace bcbam;
    std::atomic<uint64_t> nr(0), nc(0), nb(0);
    std::atomic<int> noqual(0);
    parallel_for(b->threads, rec_b - rec_a, 1 << 13, [&](uint64_t a, uint64_t e) {
        uint64_t r = 0, c = 0, s = 0;
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            if (missing_qual(v)) noqual = 1;
            uint32_t lead, trail;
            clips(v.cig, v.n_cigar, lead, trail);
            const uint64_t s0 = std::min<uint64_t>(lead, v.l_seq);
            const uint64_t s1 = std::max<uint64_t>(s0, (uint64_t)v.l_seq > trail ? v.l_seq - trail : 0);
            r++;
            c += v.n_cigar;
            s += s1 - s0;
        }
        nr += r;
        nc += c;
        nb += s;
    });
    *n_reads = nr;
    *n_cigar = nc;
    *n_bases = nb;
    return noqual == 0;
}

// Fill a ReadBatch for the selection.  Arrays sized by bc_bam_select_sizes (offset arrays n+1).
inline void bc_bam_select_fill_impl(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                                    uint32_t *starts, uint32_t *cigar, uint64_t *cigar_off, uint8_t *seq, uint8_t *qual,
                                    uint64_t *seq_off)
{
    using namespace bcbam;
    static const char kNib[] = "=ACMGRSVTWYHKDBN";
    uint16_t pair[256];                                     // packed byte -> its two letters, in memory order
    for (int x = 0; x < 256; x++) {
        const uint8_t two[2] = {(uint8_t)kNib[x >> 4], (uint8_t)kNib[x & 15]};
        std::memcpy(&pair[x], two, 2);
    }
    const uint64_t n = rec_b - rec_a;
    const uint64_t grain = 1 << 13;
    const uint64_t chunks = (n + grain - 1) / grain;
    // pass 1: per-chunk totals -> exclusive offsets, so every chunk writes its own disjoint range
    std::vector<uint64_t> cr(chunks + 1, 0), cc(chunks + 1, 0), cs(chunks + 1, 0);
    parallel_for(b->threads, n, grain, [&](uint64_t a, uint64_t e) {
        uint64_t r = 0, c = 0, s = 0;
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            uint32_t lead, trail;
            clips(v.cig, v.n_cigar, lead, trail);
            const uint64_t s0 = std::min<uint64_t>(lead, v.l_seq);
            const uint64_t s1 = std::max<uint64_t>(s0, (uint64_t)v.l_seq > trail ? v.l_seq - trail : 0);
            r++;
            c += v.n_cigar;
            s += s1 - s0;
        }
        const uint64_t k = a / grain;
        cr[k + 1] = r;
        cc[k + 1] = c;
        cs[k + 1] = s;
    });
    for (uint64_t k = 0; k < chunks; k++) {
        cr[k + 1] += cr[k];
        cc[k + 1] += cc[k];
        cs[k + 1] += cs[k];
    }
    cigar_off[0] = 0;
    seq_off[0] = 0;
    parallel_for(b->threads, n, grain, [&](uint64_t a, uint64_t e) {
        const uint64_t k = a / grain;
        uint64_t r = cr[k], c = cc[k], s = cs[k];
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            uint32_t lead, trail;
            clips(v.cig, v.n_cigar, lead, trail);
            const uint64_t s0 = std::min<uint64_t>(lead, v.l_seq);
            const uint64_t s1 = std::max<uint64_t>(s0, (uint64_t)v.l_seq > trail ? v.l_seq - trail : 0);
            starts[r] = (uint32_t)v.pos;
            for (uint32_t t = 0; t < v.n_cigar; t++) cigar[c + t] = rd32(v.cig + 4 * t);
            {
                uint64_t q = s0;
                uint8_t *dst = seq + s;
                if ((q & 1) && q < s1) {                                // odd start (an odd-length soft clip): one nibble
                    *dst++ = (uint8_t)kNib[v.seq[q >> 1] & 15];
                    q++;
                }
                for (; q + 2 <= s1; q += 2, dst += 2) std::memcpy(dst, &pair[v.seq[q >> 1]], 2);   // two bases per byte
                if (q < s1) *dst = (uint8_t)kNib[v.seq[q >> 1] >> 4];
            }
            if (qual) std::memcpy(qual + s, v.qual + s0, s1 - s0);   // NULL: the caller filters nothing by base quality
            r++;
            c += v.n_cigar;
            s += s1 - s0;
            cigar_off[r] = c;
            seq_off[r] = s;
        }
    });
}

// ---- selection straight into the device's packed batch (one reference slot) -------------------
// bc_bam_select_fill + bc_pack_reads in one pass over the records: the 4-bit bases of a kept read go
// directly to the 2-bit bit-planar words (struct bc_batch), without the ASCII copy in between.  Same
// arrays as the two-step path (tests/test_bamio.py compares them).
struct bc_pack_sizes {
    uint64_t n_reads, n_cigar, n_words, n_bases, aligned_bases;
    uint32_t sorted;                         // starts are non-decreasing
    uint32_t missing_qual;                   // a kept read has no QUAL (the reference raises TypeError there)
};

namespace bcbam {

// Nibble -> class: 0..3 = A,C,G,T; 4 = N; 5 = anything else (never counted, count.cpp:58-65).
inline const uint8_t *nibble_class()
{
    static const uint8_t k[16] = {5, 0, 1, 5, 2, 5, 5, 5, 3, 5, 5, 5, 5, 5, 5, 4};   // "=ACMGRSVTWYHKDBN"
    return k;
}

struct Trim { uint64_t s0, s1; };
inline Trim trimmed(const RecView &v)
{
    uint32_t lead, trail;
    clips(v.cig, v.n_cigar, lead, trail);
    Trim t;
    t.s0 = std::min<uint64_t>(lead, v.l_seq);
    t.s1 = std::max<uint64_t>(t.s0, (uint64_t)v.l_seq > trail ? v.l_seq - trail : 0);
    return t;
}

}  // namespace bcbam

inline void bc_bam_pack_sizes_impl(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                                   bc_pack_sizes *out)
{
    using namespace bcbam;
    const uint64_t n = rec_b - rec_a, grain = 1 << 13, chunks = (n + grain - 1) / grain;
    struct Part { uint64_t r, c, w, s, al; int32_t first, last; bool sorted, any, noqual; };
    std::vector<Part> parts(chunks ? chunks : 1);
    parallel_for(b->threads, n, grain, [&](uint64_t a, uint64_t e) {
        Part p = {0, 0, 0, 0, 0, 0, 0, true, false, false};
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            if (missing_qual(v)) p.noqual = true;
            const Trim t = trimmed(v);
            p.r++;
            p.c += bccanon::canon_cigar(v.n_cigar, [&](uint32_t k) { return rd32(v.cig + 4 * k); }, nullptr);
            p.s += t.s1 - t.s0;
            p.w += (t.s1 - t.s0 + 31) / 32;
            for (uint32_t k = 0; k < v.n_cigar; k++) {
                const uint32_t w = rd32(v.cig + 4 * k), op = w & 15u;
                if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) p.al += w >> 4;
            }
            if (!p.any) p.first = v.pos;
            else if (v.pos < p.last) p.sorted = false;
            p.last = v.pos;
            p.any = true;
        }
        parts[a / grain] = p;
    });
    bc_pack_sizes z = {0, 0, 0, 0, 0, 1, 0};
    bool any = false;
    int32_t last = 0;
    for (uint64_t k = 0; k < chunks; k++) {
        const Part &p = parts[k];
        z.n_reads += p.r;
        z.n_cigar += p.c;
        z.n_words += p.w;
        z.n_bases += p.s;
        z.aligned_bases += p.al;
        if (p.noqual) z.missing_qual = 1;
        if (!p.any) continue;
        if (!p.sorted || (any && p.first < last)) z.sorted = 0;
        last = p.last;
        any = true;
    }
    *out = z;
}

// Arrays sized from bc_bam_pack_sizes: starts[n], cigar[n_cigar], cigar_off[n+1], seq_woff[n+1], planes[n_words],
// okmask[n_words] (NULL unless min_base_quality > 0).  Exceptions as in bc_pack_reads: (read, pos << 2 | flags),
// sorted by (read, pos); *n_exc may exceed exc_cap (then only the first exc_cap were written: call again).
// Returns 0, or 5 (BC_ERR_READ_OVERRUN) when a CIGAR consumes more bases than its read holds.
inline int bc_bam_pack_fill_impl(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                                 uint32_t min_base_quality, uint32_t *starts, uint32_t *cigar, uint32_t *cigar_off,
                                 uint32_t *seq_woff, uint64_t *planes, uint32_t *okmask, uint32_t *exc_read,
                                 uint32_t *exc_pos, uint64_t exc_cap, uint64_t *n_exc)
{
    using namespace bcbam;
    const uint8_t *cls = nibble_class();
    const uint64_t n = rec_b - rec_a, grain = 1 << 12, chunks = (n + grain - 1) / grain;
    std::vector<uint64_t> cr(chunks + 1, 0), cc(chunks + 1, 0), cw(chunks + 1, 0);
    parallel_for(b->threads, n, grain, [&](uint64_t a, uint64_t e) {
        uint64_t r = 0, c = 0, w = 0;
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            const Trim t = trimmed(v);
            r++;
            c += bccanon::canon_cigar(v.n_cigar, [&](uint32_t k) { return rd32(v.cig + 4 * k); }, nullptr);
            w += (t.s1 - t.s0 + 31) / 32;
        }
        const uint64_t k = a / grain;
        cr[k + 1] = r;
        cc[k + 1] = c;
        cw[k + 1] = w;
    });
    for (uint64_t k = 0; k < chunks; k++) {
        cr[k + 1] += cr[k];
        cc[k + 1] += cc[k];
        cw[k + 1] += cw[k];
    }
    if (cr[chunks] > 0xFFFFFFFFull || cc[chunks] > 0xFFFFFFFFull || cw[chunks] > 0xFFFFFFFFull) return 1;
    cigar_off[0] = 0;
    seq_woff[0] = 0;
    std::vector<std::vector<uint32_t>> exc(chunks ? chunks : 1);
    std::atomic<int> overrun(0);
    parallel_for(b->threads, n, grain, [&](uint64_t a, uint64_t e) {
        const uint64_t k = a / grain;
        uint64_t r = cr[k], c = cc[k], w = cw[k];
        std::vector<uint32_t> &ex = exc[k];
        RecView v;
        for (uint64_t i = rec_a + a; i < rec_a + e; i++) {
            view(b, i, v);
            if (!keep(v, ref_id, min_mapq)) continue;
            const Trim t = trimmed(v);
            const uint64_t len = t.s1 - t.s0;
            starts[r] = (uint32_t)v.pos;
            uint64_t rp = 0;                                             // count.cpp:56,58 index the read unchecked
            const uint32_t n_canon = bccanon::canon_cigar(v.n_cigar, [&](uint32_t k) { return rd32(v.cig + 4 * k); }, cigar + c);
            for (uint32_t q = 0; q < v.n_cigar; q++) {
                const uint32_t word = rd32(v.cig + 4 * q), op = word & 15u, l = word >> 4;
                if (op == 0u || op == 7u || op == 8u) {
                    if (l && rp + l > len) overrun = 1;
                    rp += l;
                } else if (op == 1u) {
                    rp += l;
                }
            }
            for (uint64_t j0 = 0; j0 < len; j0 += 32, w++) {
                const uint32_t m = (uint32_t)std::min<uint64_t>(32, len - j0);
                const uint64_t q0 = t.s0 + j0;                           // first base of this word in the record
                uint32_t lo = 0, hi = 0, odd = 0, ok = 0;
#if defined(__SSE2__)
                if (m == 32 && !(q0 & 1)) {                              // 16 packed bytes -> 32 nibbles in order
                    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(v.seq + (q0 >> 1)));
                    const __m128i f = _mm_set1_epi8(0x0F);
                    const __m128i hn = _mm_and_si128(_mm_srli_epi16(x, 4), f), ln = _mm_and_si128(x, f);
                    const __m128i h[2] = {_mm_unpacklo_epi8(hn, ln), _mm_unpackhi_epi8(hn, ln)};
                    uint32_t ma = 0, mc = 0, mg = 0, mt = 0;
                    for (int z = 0; z < 2; z++) {
                        ma |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(h[z], _mm_set1_epi8(1))) << (16 * z);
                        mc |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(h[z], _mm_set1_epi8(2))) << (16 * z);
                        mg |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(h[z], _mm_set1_epi8(4))) << (16 * z);
                        mt |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(h[z], _mm_set1_epi8(8))) << (16 * z);
                    }
                    lo = mc | mt;
                    hi = mg | mt;
                    odd = ~(ma | mc | mg | mt);
                } else
#endif
                {
                    for (uint32_t j = 0; j < m; j++) {
                        const uint64_t q = q0 + j;
                        const uint8_t byte = v.seq[q >> 1];
                        const uint32_t cl = cls[(q & 1) ? (byte & 15) : (byte >> 4)];
                        lo |= (cl & 1u) << j;
                        hi |= ((cl >> 1) & 1u) << j;
                        odd |= (cl >> 2) << j;
                    }
                    lo &= ~odd;
                    hi &= ~odd;
                }
                if (min_base_quality > 0) {
                    const uint8_t *qp = v.qual + q0;
                    for (uint32_t j = 0; j < m; j++) ok |= (uint32_t)(qp[j] >= min_base_quality) << j;
                } else {
                    ok = m == 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
                }
                for (uint32_t rest = odd & (m == 32 ? 0xFFFFFFFFu : ((1u << m) - 1u)); rest; rest &= rest - 1u) {
                    const uint32_t j = (uint32_t)__builtin_ctz(rest);
                    const uint64_t q = q0 + j;
                    const uint8_t byte = v.seq[q >> 1];
                    const bool is_n = cls[(q & 1) ? (byte & 15) : (byte >> 4)] == 4;
                    uint32_t flags = 0;
                    if (min_base_quality == 0) flags = is_n ? 3u : 2u;             // undo the 'A', maybe count N
                    else if (is_n && ((ok >> j) & 1u)) flags = 1u;                  // masked out already; count N
                    if (flags) {
                        ex.push_back((uint32_t)r);
                        ex.push_back((uint32_t)((j0 + j) << 2) | flags);
                    }
                }
                planes[w] = (uint64_t)lo | ((uint64_t)hi << 32);
                if (okmask) okmask[w] = ok & ~odd;
            }
            r++;
            c += n_canon;
            cigar_off[r] = (uint32_t)c;
            seq_woff[r] = (uint32_t)w;
        }
    });
    uint64_t ne = 0;
    for (const auto &ex : exc)
        for (size_t k = 0; k + 1 < ex.size(); k += 2, ne++)
            if (ne < exc_cap && exc_read && exc_pos) {
                exc_read[ne] = ex[k];
                exc_pos[ne] = ex[k + 1];
            }
    *n_exc = ne;
    return overrun ? 5 : 0;
}
