// inflate_fast.h -- a raw-DEFLATE (RFC 1951) decoder for BGZF blocks: whole block in, whole block out.
//
// Host code of the BAM path.  A BGZF block is one small independent deflate stream whose
// inflated size is known (ISIZE) and whose content is checked by a CRC-32 afterwards, which allows a decoder
// without zlib's streaming state machine: a 64-bit bit buffer refilled without branches, an 11-bit first-level
// table for the literal/length code (BAM payloads are literal-heavy -- packed bases and qualities -- with 8 to
// 11-bit codes, which zlib's 9-bit root table sends through a second lookup), length and distance extra bits
// taken from the same refill, up to three literals per refill, and word-wise match copies.  Anything unusual
// (malformed stream, output that does not end exactly at ISIZE) returns false and the caller falls back to
// zlib, so zlib stays the arbiter of what is a valid block; every block is CRC-checked either way.
#pragma once
#include <stdint.h>
#include <string.h>

namespace bcbam {

class FastInflater {
public:
    // in[0, in_len): the raw deflate stream; out[0, out_len): exactly the inflated bytes.  in must be readable
    // up to in + in_len + 8 (the BGZF trailer follows it in the file).  The same body is compiled twice: with
    // BMI2 (shrx / bzhi shorten the bit-buffer chain every symbol waits on: 1.3 x -> 1.6 x zlib) and without.
    bool inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
    {
#if defined(__x86_64__)
        static const bool bmi2 = __builtin_cpu_supports("bmi2");
        if (bmi2) return inflate_bmi2(in, in_len, out, out_len);
#endif
        return inflate_body(in, in_len, out, out_len);
    }

private:
#if defined(__x86_64__)
    __attribute__((target("bmi2,bmi"), noinline)) bool inflate_bmi2(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
    {
        return inflate_body(in, in_len, out, out_len);
    }
#endif
    __attribute__((always_inline)) inline bool inflate_body(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
    {
        const uint8_t *ip = in, *const in_end = in + in_len;
        uint8_t *op = out, *const out_end = out + out_len;
        uint64_t bb = 0;                                       // bit buffer, next bit = bit 0
        int bl = 0;                                            // valid bits in bb; negative once a damaged or
                                                               // truncated stream has asked for more than there is
        bool last = false;
        // Refill while the read pointer is inside the stream (the 8-byte load may take in the trailer).  Past the end
        // nothing is loaded any more: a valid stream finishes on the bits it has, anything else runs bl negative and is
        // declined at the next check (block header, code lengths, match, end of block) -- literals are bounded by out_end.
#define BCI_REFILL()                                                                   \
    do {                                                                               \
        if (ip <= in_end) {                                                            \
            if (bl < 0) return false;                                                  \
            uint64_t w;                                                                \
            memcpy(&w, ip, 8);                                                         \
            bb |= w << bl;                                                             \
            ip += (63 - bl) >> 3;                                                      \
            bl |= 56;                                                                  \
        }                                                                              \
    } while (0)
        while (!last) {
            BCI_REFILL();
            last = bb & 1u;
            const unsigned type = (unsigned)(bb >> 1) & 3u;
            bb >>= 3;
            bl -= 3;
            if (bl < 0) return false;
            if (type == 0) {                                   // stored block
                const int drop = bl & 7;
                bb >>= drop;
                bl -= drop;
                // give whole unread bytes back to the input
                ip -= bl >> 3;
                bb = 0;
                bl = 0;
                if (ip + 4 > in_end) return false;
                const unsigned len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
                ip += 4;
                if ((len ^ nlen) != 0xFFFFu || ip + len > in_end || op + len > out_end) return false;
                memcpy(op, ip, len);
                ip += len;
                op += len;
                continue;
            }
            if (type == 3) return false;
            if (type == 1) {
                if (!build_fixed()) return false;
            } else {
                // dynamic block header
                const unsigned hlit = ((unsigned)bb & 31u) + 257u, hdist = ((unsigned)(bb >> 5) & 31u) + 1u,
                               hclen = ((unsigned)(bb >> 10) & 15u) + 4u;
                bb >>= 14;
                bl -= 14;
                if (bl < 0 || hlit > 286u || hdist > 30u) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19];
                memset(cl, 0, sizeof(cl));
                for (unsigned i = 0; i < hclen; i++) {
                    if (bl < 3) BCI_REFILL();
                    cl[order[i]] = (uint8_t)(bb & 7u);
                    bb >>= 3;
                    bl -= 3;
                }
                if (bl < 0) return false;
                if (!build(cl, 19, pre_, 7, kPreSize, nullptr, nullptr, 0)) return false;
                uint8_t lens[286 + 30 + 138];
                unsigned n = 0;
                const unsigned total = hlit + hdist;
                while (n < total) {
                    BCI_REFILL();
                    const uint32_t e = pre_[bb & 127u];
                    const unsigned cb = e & 0xFFu;
                    if (cb == 0) return false;
                    bb >>= cb;
                    bl -= (int)cb;
                    const unsigned sym = e >> 16;
                    if (sym < 16) {
                        lens[n++] = (uint8_t)sym;
                    } else if (sym == 16) {
                        if (n == 0) return false;
                        unsigned r = 3u + ((unsigned)bb & 3u);
                        bb >>= 2;
                        bl -= 2;
                        const uint8_t v = lens[n - 1];
                        while (r--) lens[n++] = v;
                    } else if (sym == 17) {
                        unsigned r = 3u + ((unsigned)bb & 7u);
                        bb >>= 3;
                        bl -= 3;
                        while (r--) lens[n++] = 0;
                    } else {
                        unsigned r = 11u + ((unsigned)bb & 127u);
                        bb >>= 7;
                        bl -= 7;
                        while (r--) lens[n++] = 0;
                    }
                    if (bl < 0) return false;
                }
                if (n != total || lens[256] == 0) return false;
                if (!build(lens, hlit, lit_, kLitBits, kLitSize, kLenBase, kLenExtra, 257)) return false;
                if (!build(lens + hlit, hdist, dist_, kDistBits, kDistSize, kDistBase, kDistExtra, 0)) return false;
            }
            // ---- the block's symbols
            for (;;) {
                BCI_REFILL();                                   // >= 56 bits
                uint32_t e = lit_[bb & ((1u << kLitBits) - 1u)];
                if (e & kSub) e = lit_[(e >> 16) + (((unsigned)(bb >> kLitBits)) & ((1u << ((e >> 8) & 15u)) - 1u))];
                if (e & kLiteral) {                             // up to three literals per refill (3 x 15 bits < 56)
                    if (op + 3 > out_end) {
                        if (op >= out_end) return false;
                        bb >>= (e & 0xFFu);
                        bl -= (int)(e & 0xFFu);
                        *op++ = (uint8_t)(e >> 16);
                        continue;
                    }
                    bb >>= (e & 0xFFu);
                    bl -= (int)(e & 0xFFu);
                    *op++ = (uint8_t)(e >> 16);
                    e = lit_[bb & ((1u << kLitBits) - 1u)];
                    if (e & kSub) e = lit_[(e >> 16) + (((unsigned)(bb >> kLitBits)) & ((1u << ((e >> 8) & 15u)) - 1u))];
                    if (e & kLiteral) {
                        bb >>= (e & 0xFFu);
                        bl -= (int)(e & 0xFFu);
                        *op++ = (uint8_t)(e >> 16);
                        e = lit_[bb & ((1u << kLitBits) - 1u)];
                        if (e & kSub) e = lit_[(e >> 16) + (((unsigned)(bb >> kLitBits)) & ((1u << ((e >> 8) & 15u)) - 1u))];
                        if (e & kLiteral) {
                            bb >>= (e & 0xFFu);
                            bl -= (int)(e & 0xFFu);
                            *op++ = (uint8_t)(e >> 16);
                            continue;
                        }
                    }
                    // e is a length / end-of-block entry; at most 30 bits are gone, 26+ are left: refill for the match
                    BCI_REFILL();
                }
                const unsigned cb = e & 0xFFu;
                if (cb == 0) return false;                      // unused code
                bb >>= cb;
                bl -= (int)cb;
                if (e & kEob) {
                    if (bl < 0) return false;
                    break;
                }
                const unsigned lx = (e >> 8) & 15u;
                unsigned len = (e >> 16) + ((unsigned)bb & ((1u << lx) - 1u));
                bb >>= lx;
                bl -= (int)lx;
                // distance: <= 15 code bits + 13 extra; 56 - 15 - 5 = 36 bits were left at least
                uint32_t d = dist_[bb & ((1u << kDistBits) - 1u)];
                if (d & kSub) d = dist_[(d >> 16) + (((unsigned)(bb >> kDistBits)) & ((1u << ((d >> 8) & 15u)) - 1u))];
                const unsigned db = d & 0xFFu;
                if (db == 0) return false;
                bb >>= db;
                bl -= (int)db;
                const unsigned dx = (d >> 8) & 15u;
                const unsigned dist = (d >> 16) + ((unsigned)bb & ((1u << dx) - 1u));
                bb >>= dx;
                bl -= (int)dx;
                if (bl < 0 || dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return false;
                const uint8_t *src = op - dist;
                if (dist >= 8 && (size_t)(out_end - op) >= len + 8u) {      // word-wise, may write up to 7 bytes past len
                    uint8_t *dst = op;
                    op += len;
                    do {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(dst, &w, 8);
                        src += 8;
                        dst += 8;
                    } while (dst < op);
                } else {
                    while (len--) *op++ = *src++;
                }
            }
            if (ip > in_end + 8) return false;
        }
#undef BCI_REFILL
        // whole bytes still in the bit buffer were not consumed
        if (bl < 0) return false;
        const uint8_t *used = ip - (bl >> 3);
        return op == out_end && used <= in_end;
    }

    static constexpr unsigned kLitBits = 11, kDistBits = 8;
    static constexpr unsigned kLitSize = (1u << kLitBits) + 288u * 16u, kDistSize = (1u << kDistBits) + 32u * 128u,
                              kPreSize = 128;
    // entry: [31:16] literal / base / subtable start, [15] literal, [14] end of block, [13] subtable link,
    //        [11:8] extra bits (link: subtable index bits), [7:0] code bits to consume (0 = unused code)
    static constexpr uint32_t kLiteral = 1u << 15, kEob = 1u << 14, kSub = 1u << 13;
    uint32_t lit_[kLitSize], dist_[kDistSize], pre_[kPreSize];
    bool fixed_built_ = false;
    uint32_t fixed_lit_[kLitSize], fixed_dist_[kDistSize];

    static constexpr uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                              67, 83, 99, 115, 131, 163, 195, 227, 258};
    static constexpr uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static constexpr uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                               1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static constexpr uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

    static unsigned rev(unsigned code, unsigned len)
    {
        unsigned r = 0;
        for (unsigned i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1u - i);
        return r;
    }

    // Canonical Huffman decode table.  Symbols < first_coded are literals (pre-code and distance tables pass
    // base == nullptr / first_coded == 0 accordingly); symbol 256 of the literal/length code is end of block.
    static bool build(const uint8_t *lens, unsigned n, uint32_t *tab, unsigned tbits, unsigned tsize, const uint16_t *base,
                      const uint8_t *extra, unsigned first_coded)
    {
        unsigned count[16] = {0}, next[16];
        for (unsigned i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        unsigned code = 0, maxlen = 0, used = 0;
        for (unsigned l = 1; l < 16; l++) {
            code = (code + count[l - 1]) << 1;
            next[l] = code;
            if (count[l]) maxlen = l;
            used += count[l];
        }
        // over-subscribed codes are malformed; incomplete ones are allowed only as zlib allows them (one code)
        {
            int left = 1;
            for (unsigned l = 1; l < 16; l++) {
                left <<= 1;
                left -= (int)count[l];
                if (left < 0) return false;
            }
            if (left > 0 && used > 1) return false;
        }
        for (unsigned i = 0; i < (1u << tbits); i++) tab[i] = 0;
        unsigned sub_next = 1u << tbits;
        const unsigned sub_bits = maxlen > tbits ? maxlen - tbits : 0;
        for (unsigned s = 0; s < n; s++) {
            const unsigned l = lens[s];
            if (!l) continue;
            const unsigned c = rev(next[l]++, l);
            uint32_t e;
            if (base == nullptr && first_coded == 0 && extra == nullptr) {
                e = (s << 16) | l;                                          // pre-code: plain symbol
            } else if (first_coded && s < 256) {
                e = (s << 16) | kLiteral | l;
            } else if (first_coded && s == 256) {
                e = kEob | l;
            } else {
                const unsigned k = s - first_coded;
                if (k >= (first_coded ? 29u : 30u)) return false;           // codes 286/287, 30/31 must not be used
                e = ((uint32_t)base[k] << 16) | ((uint32_t)extra[k] << 8) | l;
            }
            if (l <= tbits) {
                for (unsigned i = c; i < (1u << tbits); i += 1u << l) tab[i] = e;
            } else {
                const unsigned root = c & ((1u << tbits) - 1u);
                if (!(tab[root] & kSub)) {
                    if (sub_next + (1u << sub_bits) > tsize) return false;
                    tab[root] = (sub_next << 16) | kSub | (sub_bits << 8) | tbits;
                    for (unsigned i = 0; i < (1u << sub_bits); i++) tab[sub_next + i] = 0;
                    sub_next += 1u << sub_bits;
                }
                const unsigned start = tab[root] >> 16;
                for (unsigned i = c >> tbits; i < (1u << sub_bits); i += 1u << (l - tbits)) tab[start + i] = e;
            }
        }
        return true;
    }

    bool build_fixed()
    {
        if (!fixed_built_) {
            uint8_t l[288 + 32];
            for (unsigned i = 0; i < 144; i++) l[i] = 8;
            for (unsigned i = 144; i < 256; i++) l[i] = 9;
            for (unsigned i = 256; i < 280; i++) l[i] = 7;
            for (unsigned i = 280; i < 288; i++) l[i] = 8;
            for (unsigned i = 0; i < 32; i++) l[288 + i] = 5;
            // (symbols 286/287 and distances 30/31 exist in the fixed code but never occur in a valid stream:
            //  build them as unused so that they fail)
            uint8_t ll[288];
            memcpy(ll, l, 288);
            if (!build_allow_reserved(ll, 288, fixed_lit_, kLitBits, kLitSize, kLenBase, kLenExtra, 257, 286)) return false;
            if (!build_allow_reserved(l + 288, 32, fixed_dist_, kDistBits, kDistSize, kDistBase, kDistExtra, 0, 30)) return false;
            fixed_built_ = true;
        }
        memcpy(lit_, fixed_lit_, sizeof(lit_));
        memcpy(dist_, fixed_dist_, sizeof(dist_));
        return true;
    }

    // The fixed code assigns codes to reserved symbols; they take part in the canonical numbering but decode as unused.
    static bool build_allow_reserved(const uint8_t *lens, unsigned n, uint32_t *tab, unsigned tbits, unsigned tsize,
                                     const uint16_t *base, const uint8_t *extra, unsigned first_coded, unsigned first_reserved)
    {
        unsigned count[16] = {0}, next[16];
        for (unsigned i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        unsigned code = 0;
        for (unsigned l = 1; l < 16; l++) {
            code = (code + count[l - 1]) << 1;
            next[l] = code;
        }
        for (unsigned i = 0; i < (1u << tbits); i++) tab[i] = 0;
        (void)tsize;
        for (unsigned s = 0; s < n; s++) {
            const unsigned l = lens[s];
            const unsigned c = rev(next[l]++, l);
            if (s >= first_reserved) continue;
            uint32_t e;
            if (first_coded && s < 256) e = (s << 16) | kLiteral | l;
            else if (first_coded && s == 256) e = kEob | l;
            else {
                const unsigned k = s - first_coded;
                e = ((uint32_t)base[k] << 16) | ((uint32_t)extra[k] << 8) | l;
            }
            for (unsigned i = c; i < (1u << tbits); i += 1u << l) tab[i] = e;       // (all fixed codes are <= 9 bits)
        }
        return true;
    }
};

}  // namespace bcbam
