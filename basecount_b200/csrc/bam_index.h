// bam_index.h -- BAI index (SAM spec section 5.2): build one for a coordinate-sorted BAM, and open only the
// part of a BAM that holds the records STARTING in [beg, end) of one reference.
//
// the region-sharded path (config 5) gives every GPU's host process the
// reads whose start lies in its region (the split of tests/test_basecount.py:146-150).  With the whole-file
// decoder (bam_decode.h) every rank inflates the whole BAM; with an index a rank reads and inflates only
// the BGZF blocks between the linear-index offset of its first 16 kbp window and the offset of the window
// behind its last one, so decode time scales with 1 / ranks.  The reference itself never reads an index
// (`fetch(until_eof=True)`, basecount/main.py:127); its tests index BAMs for their pysam pileup oracle
// (tests/test_basecount.py:343-344), which is the decomposition reused here.
//
// Only the linear index is needed to find where a region starts (a record that starts at or after `beg`
// overlaps window beg >> 14 or a later one, and the file is coordinate-sorted); the binning index is
// written for other tools and read only for a reference's first / last offset.  Host-only code.
#pragma once
#include "bam_decode.h"

#include <map>

namespace bcbam {

inline uint32_t reg2bin(int64_t beg, int64_t end)          // SAM spec 5.3
{
    --end;
    if (beg >> 14 == end >> 14) return (uint32_t)(4681 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint32_t)(585 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint32_t)(73 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint32_t)(9 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint32_t)(1 + (beg >> 26));
    return 0;
}

constexpr uint32_t kMetaBin = 37450;                       // samtools' pseudo-bin: offsets and read counts
constexpr uint64_t kNoOffset = ~0ull;

inline void wr32(ByteVec &o, uint32_t v) { for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i))); }
inline void wr64(ByteVec &o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

inline bool load_file(const char *path, ByteVec &out, std::string &err)
{
    FILE *fh = std::fopen(path, "rb");
    if (!fh) {
        err = std::string("cannot open ") + path;
        return false;
    }
    std::fseek(fh, 0, SEEK_END);
    const long sz = std::ftell(fh);
    std::fseek(fh, 0, SEEK_SET);
    if (sz < 0) {
        std::fclose(fh);
        err = "cannot size the file";
        return false;
    }
    out.resize((size_t)sz);
    const bool ok = sz == 0 || std::fread(out.data(), 1, (size_t)sz, fh) == (size_t)sz;
    std::fclose(fh);
    if (!ok) err = "short read";
    return ok;
}

// Complete BGZF blocks inside buf (which starts at file offset `base`); a truncated block at the end is
// left alone.  Block::c0 / c1 index buf, Block::u0 continues from `total`; fpos receives file offsets.
inline bool scan_blocks_partial(const ByteVec &d, uint64_t from, uint64_t base, std::vector<Block> &out,
                                std::vector<uint64_t> &fpos, uint64_t &total, uint64_t &consumed, std::string &err)
{
    uint64_t off = from;
    const uint64_t n = d.size();
    while (off < n) {
        if (n - off < 18) break;
        if (d[off] != 31 || d[off + 1] != 139 || d[off + 2] != 8 || !(d[off + 3] & 4)) {
            err = "not a BGZF block (bad virtual offset or corrupt file)";
            return false;
        }
        const uint32_t xlen = rd16(&d[off + 10]);
        if (off + 12 + xlen > n) break;
        uint64_t p = off + 12;
        const uint64_t end = off + 12 + xlen;
        uint32_t bsize = 0;
        while (p + 4 <= end) {
            const uint32_t slen = rd16(&d[p + 2]);
            if (d[p] == 66 && d[p + 1] == 67 && slen == 2 && p + 6 <= end) bsize = (uint32_t)rd16(&d[p + 4]) + 1u;
            p += 4 + slen;
        }
        if (bsize == 0 || bsize < 12 + xlen + 8) {
            err = "BGZF block without BC subfield";
            return false;
        }
        if (off + bsize > n) break;
        Block b;
        b.c0 = off + 12 + xlen;
        b.c1 = off + bsize - 8;
        b.crc = rd32(&d[off + bsize - 8]);
        b.isize = rd32(&d[off + bsize - 4]);
        b.u0 = total;
        total += b.isize;
        out.push_back(b);
        fpos.push_back(base + off);
        off += bsize;
    }
    consumed = off;
    return true;
}

// Inflate blocks [a, e) of `src` into dst (dst already sized); false on a bad block.
inline bool inflate_blocks(int threads, const ByteVec &src, const std::vector<Block> &blocks, uint64_t a,
                           uint64_t e, uint8_t *dst, uint64_t dst_u0)
{
    std::atomic<int> bad(0);
    parallel_for(threads, e - a, 16, [&](uint64_t x, uint64_t y) {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = 1;
            return;
        }
        std::unique_ptr<FastInflater> fi(fast_inflate_enabled() ? new FastInflater() : nullptr);
        for (uint64_t i = a + x; i < a + y; i++) {
            const Block &k = blocks[i];
            if (k.isize == 0) continue;
            if (!inflate_block_checked(fi.get(), &zs, src.data() + k.c0, k.c1 - k.c0, dst + (k.u0 - dst_u0), k.isize, k.crc))
                bad = 1;
        }
        inflateEnd(&zs);
    });
    return bad == 0;
}

// BAM header (SAM spec 4.2) at the start of `r`; `p` = offset of the first record.  Returns 0 ok,
// 1 need more data, 2 not a BAM.
inline int parse_header(const ByteVec &r, std::vector<std::string> &names, std::vector<uint32_t> &lens,
                        uint64_t &p)
{
    names.clear();
    lens.clear();
    if (r.size() < 12) return 1;
    if (std::memcmp(r.data(), "BAM\1", 4) != 0) return 2;
    p = 8ull + rd32(&r[4]);
    if (p + 4 > r.size()) return 1;
    const uint32_t n_ref = rd32(&r[p]);
    p += 4;
    for (uint32_t i = 0; i < n_ref; i++) {
        if (p + 4 > r.size()) return 1;
        const uint32_t l_name = rd32(&r[p]);
        if (l_name == 0) return 2;
        if (p + 8ull + l_name > r.size()) return 1;
        names.emplace_back((const char *)&r[p + 4], l_name - 1);
        lens.push_back(rd32(&r[p + 4 + l_name]));
        p += 8ull + l_name;
    }
    return 0;
}

// Reference span of a record's CIGAR (M, D, N, =, X consume the reference).
inline uint64_t ref_span(const uint8_t *cig, uint32_t n_cigar)
{
    uint64_t s = 0;
    for (uint32_t t = 0; t < n_cigar; t++) {
        const uint32_t w = rd32(cig + 4 * t), op = w & 15u;
        if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) s += w >> 4;
    }
    return s;
}

struct RefIndex {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;
    uint64_t off_beg = kNoOffset, off_end = 0, n_mapped = 0, n_unmapped = 0;
};

struct Bai {
    std::vector<RefIndex> refs;
    uint64_t n_no_coor = 0;
};

inline bool parse_bai(const ByteVec &d, Bai &out, std::string &err)
{
    auto bad = [&]() {
        err = "truncated or malformed BAI index";
        return false;
    };
    if (d.size() < 8 || std::memcmp(d.data(), "BAI\1", 4) != 0) {
        err = "not a BAI index (bad magic)";
        return false;
    }
    uint64_t p = 4;
    const uint32_t n_ref = rd32(&d[p]);
    p += 4;
    out.refs.assign(n_ref, RefIndex());
    for (uint32_t r = 0; r < n_ref; r++) {
        RefIndex &ri = out.refs[r];
        if (p + 4 > d.size()) return bad();
        const uint32_t n_bin = rd32(&d[p]);
        p += 4;
        for (uint32_t k = 0; k < n_bin; k++) {
            if (p + 8 > d.size()) return bad();
            const uint32_t bin = rd32(&d[p]), n_chunk = rd32(&d[p + 4]);
            p += 8;
            if (p + 16ull * n_chunk > d.size()) return bad();
            if (bin == kMetaBin) {
                if (n_chunk >= 2) {
                    ri.off_beg = rd64(&d[p]);
                    ri.off_end = rd64(&d[p + 8]);
                    ri.n_mapped = rd64(&d[p + 16]);
                    ri.n_unmapped = rd64(&d[p + 24]);
                }
            } else {
                auto &v = ri.bins[bin];
                for (uint32_t c = 0; c < n_chunk; c++) v.emplace_back(rd64(&d[p + 16ull * c]), rd64(&d[p + 16ull * c + 8]));
            }
            p += 16ull * n_chunk;
        }
        if (p + 4 > d.size()) return bad();
        const uint32_t n_intv = rd32(&d[p]);
        p += 4;
        if (p + 8ull * n_intv > d.size()) return bad();
        ri.linear.resize(n_intv);
        for (uint32_t i = 0; i < n_intv; i++) ri.linear[i] = rd64(&d[p + 8ull * i]);
        p += 8ull * n_intv;
        if (ri.off_beg == kNoOffset)                                 // no pseudo-bin: first / last chunk offsets
            for (auto &kv : ri.bins)
                for (auto &c : kv.second) {
                    ri.off_beg = std::min(ri.off_beg, c.first);
                    ri.off_end = std::max(ri.off_end, c.second);
                }
    }
    if (p + 8 <= d.size()) out.n_no_coor = rd64(&d[p]);
    return true;
}

}  // namespace bcbam

// Build the BAI of a coordinate-sorted BAM.  Conventions follow samtools: a record's chunk runs from its
// virtual offset to the next record's; consecutive records of one bin share a chunk; windows without a
// record take the offset of the next window that has one; pseudo-bin 37450 holds the reference's first /
// last offset and its mapped / unmapped record counts; n_no_coor closes the file.
inline int bc_bam_index_build_impl(const char *bam_path, const char *bai_path, int threads, std::string &err)
{
    using namespace bcbam;
    ByteVec file;
    if (!load_file(bam_path, file, err)) return 1;
    std::vector<Block> blocks;
    std::vector<uint64_t> fpos;
    uint64_t total = 0, consumed = 0;
    if (!scan_blocks_partial(file, 0, 0, blocks, fpos, total, consumed, err)) return 2;
    if (consumed != file.size()) {
        err = "truncated BGZF block";
        return 2;
    }
    const int nt = threads > 0 ? threads : (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    ByteVec raw(total);
    if (!inflate_blocks(nt, file, blocks, 0, blocks.size(), raw.data(), 0)) {
        err = "corrupt BGZF block (inflate or CRC failed)";
        return 2;
    }
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    uint64_t p = 0;
    if (parse_header(raw, names, lens, p) != 0) {
        err = "not a BAM file (bad or truncated header)";
        return 2;
    }
    // virtual offset of an uncompressed offset (blocks in order; empty blocks never hold a byte)
    size_t bk = 0;
    auto voff = [&](uint64_t u) -> uint64_t {
        while (bk < blocks.size() && u >= blocks[bk].u0 + blocks[bk].isize) bk++;
        if (bk == blocks.size()) return (uint64_t)file.size() << 16;
        return (fpos[bk] << 16) | (u - blocks[bk].u0);
    };
    Bai bai;
    bai.refs.assign(names.size(), RefIndex());
    int32_t last_ref = -1, last_pos = -1;
    uint32_t last_bin = ~0u;
    while (p + 4 <= raw.size()) {
        const uint64_t o = p, sz = rd32(&raw[p]);
        if (sz < 32 || o + 4 + sz > raw.size()) {
            err = "truncated BAM record";
            return 2;
        }
        const uint8_t *q = raw.data() + o;
        const int32_t ref_id = (int32_t)rd32(q + 4), pos = (int32_t)rd32(q + 8);
        const uint32_t l_read_name = q[12], n_cigar = rd16(q + 16), flag = rd16(q + 18);
        if (36ull + l_read_name + 4ull * n_cigar > 4 + sz) {
            err = "malformed BAM record";
            return 2;
        }
        const uint64_t v0 = voff(o);
        p = o + 4 + sz;
        const uint64_t v1 = voff(p);
        if (ref_id < 0 || pos < 0) {                 // no coordinate: not indexed (samtools wants these last; here they
            if (ref_id < 0) bai.n_no_coor++;         // may sit anywhere, they only end the running chunk)
            last_bin = ~0u;
            continue;
        }
        if ((size_t)ref_id >= bai.refs.size() || ref_id < last_ref || (ref_id == last_ref && pos < last_pos)) {
            err = "the BAM is not coordinate-sorted: cannot index it";
            return 3;
        }
        RefIndex &ri = bai.refs[(size_t)ref_id];
        uint64_t span = (flag & 4u) ? 0 : ref_span(q + 36 + l_read_name, n_cigar);
        if (span == 0) span = 1;
        const int64_t beg = pos, end = (int64_t)pos + (int64_t)span;
        const uint32_t bin = reg2bin(beg, end);
        auto &chunks = ri.bins[bin];
        if (ref_id == last_ref && bin == last_bin && !chunks.empty()) chunks.back().second = v1;
        else chunks.emplace_back(v0, v1);
        const uint64_t w0 = (uint64_t)beg >> 14, w1 = (uint64_t)(end - 1) >> 14;
        if (ri.linear.size() <= w1) ri.linear.resize(w1 + 1, kNoOffset);
        for (uint64_t w = w0; w <= w1; w++)
            if (ri.linear[w] == kNoOffset) ri.linear[w] = v0;
        ri.off_beg = std::min(ri.off_beg, v0);
        ri.off_end = std::max(ri.off_end, v1);
        if (flag & 4u) ri.n_unmapped++; else ri.n_mapped++;
        last_ref = ref_id;
        last_pos = pos;
        last_bin = bin;
    }
    if (p != raw.size()) {
        err = "truncated BAM record";
        return 2;
    }
    ByteVec out;
    out.insert(out.end(), {'B', 'A', 'I', 1});
    wr32(out, (uint32_t)bai.refs.size());
    for (RefIndex &ri : bai.refs) {
        const bool any = ri.off_beg != kNoOffset;
        wr32(out, (uint32_t)ri.bins.size() + (any ? 1u : 0u));
        for (auto &kv : ri.bins) {
            wr32(out, kv.first);
            wr32(out, (uint32_t)kv.second.size());
            for (auto &c : kv.second) {
                wr64(out, c.first);
                wr64(out, c.second);
            }
        }
        if (any) {
            wr32(out, kMetaBin);
            wr32(out, 2);
            wr64(out, ri.off_beg);
            wr64(out, ri.off_end);
            wr64(out, ri.n_mapped);
            wr64(out, ri.n_unmapped);
        }
        for (size_t w = ri.linear.size(); w-- > 0;)
            if (ri.linear[w] == kNoOffset) ri.linear[w] = w + 1 < ri.linear.size() ? ri.linear[w + 1] : 0;
        wr32(out, (uint32_t)ri.linear.size());
        for (uint64_t v : ri.linear) wr64(out, v);
    }
    wr64(out, bai.n_no_coor);
    FILE *fo = std::fopen(bai_path, "wb");
    if (!fo) {
        err = std::string("cannot write ") + bai_path;
        return 1;
    }
    const bool ok = std::fwrite(out.data(), 1, out.size(), fo) == out.size();
    std::fclose(fo);
    if (!ok) {
        err = "short write";
        return 1;
    }
    return 0;
}

// Open only the records of reference `ref_id` that START in [beg, end) (0-based), through the BAI.
// The handle behaves like one from bc_bam_open whose file holds just those records.
inline int bc_bam_open_region_impl(const char *bam_path, const char *bai_path, int32_t ref_id, int64_t beg, int64_t end,
                                   int threads, bc_bam **out, std::string &err)
{
    using namespace bcbam;
    ByteVec idx;
    if (!load_file(bai_path, idx, err)) return 1;
    Bai bai;
    if (!parse_bai(idx, bai, err)) return 2;
    FILE *fh = std::fopen(bam_path, "rb");
    if (!fh) {
        err = std::string("cannot open ") + bam_path;
        return 1;
    }
    std::fseek(fh, 0, SEEK_END);
    const long fsz = std::ftell(fh);
    if (fsz < 0) {
        std::fclose(fh);
        err = "cannot size the file";
        return 1;
    }
    const uint64_t fsize = (uint64_t)fsz;
    auto read_at = [&](uint64_t off, uint64_t len, ByteVec &buf) -> bool {
        len = std::min(len, fsize > off ? fsize - off : 0);
        const size_t old = buf.size();
        buf.resize(old + len);
        if (len == 0) return true;
        if (std::fseek(fh, (long)off, SEEK_SET) != 0) return false;
        return std::fread(buf.data() + old, 1, len, fh) == len;
    };
    bc_bam *b = new bc_bam();
    b->threads = threads > 0 ? threads : (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    auto fail = [&](int code, const char *msg) {
        if (msg) err = msg;
        std::fclose(fh);
        delete b;
        return code;
    };
    // ---- header: inflate leading blocks until it is complete
    {
        ByteVec buf, raw;
        std::vector<Block> blocks;
        std::vector<uint64_t> fpos;
        uint64_t total = 0, consumed = 0, have = 0, done_blocks = 0, p = 0;
        for (uint64_t want = 1u << 18;; want <<= 2) {
            if (!read_at(have, want - have, buf)) return fail(1, "short read");
            have = buf.size();
            if (!scan_blocks_partial(buf, consumed, 0, blocks, fpos, total, consumed, err)) return fail(2, nullptr);
            raw.resize(total);
            if (!inflate_blocks(1, buf, blocks, done_blocks, blocks.size(), raw.data(), 0))
                return fail(2, "corrupt BGZF block (inflate or CRC failed)");
            done_blocks = blocks.size();
            const int rc = parse_header(raw, b->ref_names, b->ref_lens, p);
            if (rc == 0) break;
            if (rc == 2) return fail(2, "not a BAM file (bad magic)");
            if (have >= fsize) return fail(2, "truncated BAM header");
        }
    }
    b->rec_off.clear();
    if (bai.refs.size() != b->ref_names.size()) return fail(2, "the index does not belong to this BAM (reference count differs)");
    if (ref_id < 0 || (size_t)ref_id >= bai.refs.size()) return fail(3, "reference id out of range");
    beg = std::max<int64_t>(beg, 0);
    const RefIndex &ri = bai.refs[(size_t)ref_id];
    const uint64_t w0 = (uint64_t)beg >> 14;
    if (end <= beg || ri.off_beg == kNoOffset || w0 >= ri.linear.size()) {       // nothing starts in the region
        b->rec_off.push_back(0);
        std::fclose(fh);
        *out = b;
        return 0;
    }
    const uint64_t v0 = ri.linear[w0];
    // a first guess of where the region's records end: the first record that overlaps the window behind
    // the one holding end - 1, else the reference's last offset; the walk below extends it if it was short
    const uint64_t w1 = ((uint64_t)(end - 1) >> 14) + 2;
    const uint64_t v1 = w1 < ri.linear.size() ? std::max(ri.linear[w1], v0) : std::max(ri.off_end, v0);
    const uint64_t c_first = v0 >> 16;
    uint64_t c_next = c_first;                     // file offset of the first block not read yet
    uint64_t want_to = std::min<uint64_t>(fsize, (v1 >> 16) + 0x10000ull);
    ByteVec buf;
    std::vector<Block> blocks;
    std::vector<uint64_t> fpos;
    uint64_t total = 0, consumed = 0, done_blocks = 0;
    uint64_t p = v0 & 0xFFFFull;                   // walk position in b->raw
    uint64_t u_end = kNoOffset;                    // where the reference's last record ends, once that block is loaded
    bool finished = false;
    while (!finished) {
        if (want_to > c_next) {
            if (!read_at(c_next, want_to - c_next, buf)) return fail(1, "short read");
            c_next = want_to;
        }
        if (!scan_blocks_partial(buf, consumed, c_first, blocks, fpos, total, consumed, err)) return fail(2, nullptr);
        b->raw.resize(total);
        if (!inflate_blocks(b->threads, buf, blocks, done_blocks, blocks.size(), b->raw.data(), 0))
            return fail(2, "corrupt BGZF block (inflate or CRC failed)");
        done_blocks = blocks.size();
        if (u_end == kNoOffset) {
            const auto it = std::lower_bound(fpos.begin(), fpos.end(), ri.off_end >> 16);
            if (it != fpos.end() && *it == (ri.off_end >> 16)) u_end = blocks[(size_t)(it - fpos.begin())].u0 + (ri.off_end & 0xFFFFull);
        }
        const ByteVec &r = b->raw;
        bool starved = false;
        while (true) {
            if (p >= u_end) {                      // behind the reference's last record
                finished = true;
                break;
            }
            if (p + 4 > r.size()) {
                starved = true;
                break;
            }
            const uint64_t sz = rd32(&r[p]);
            if (sz < 32) return fail(2, "malformed BAM record");
            if (p + 4 + sz > r.size()) {
                starved = true;
                break;
            }
            const int32_t rid = (int32_t)rd32(&r[p + 4]), pos = (int32_t)rd32(&r[p + 8]);
            if (rid != ref_id || pos >= end) {
                if (rid == ref_id || rid > ref_id) {
                    finished = true;
                    break;
                }
                p += 4 + sz;                       // no coordinate (or an earlier reference): not ours, keep walking
                continue;
            }
            if (pos >= beg) {
                b->rec_off.push_back(p);
            }
            p += 4 + sz;
        }
        if (finished) break;
        if (starved) {
            if (c_next >= fsize) {
                if (p != r.size()) return fail(2, "truncated BAM record");
                break;                             // end of file
            }
            want_to = std::min<uint64_t>(fsize, c_next + std::max<uint64_t>(1u << 22, c_next - c_first));
        }
    }
    std::fclose(fh);
    b->rec_off.push_back(p);
    RecView v;
    for (uint64_t i = 0; i + 1 < b->rec_off.size(); i++)
        if (!view(b, i, v)) {
            err = "malformed BAM record";
            delete b;
            return 2;
        }
    *out = b;
    return 0;
}
