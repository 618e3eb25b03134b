// cigar_canon.h -- the packers' CIGAR normal form (host code).
//
// The reference's loop (basecount/count.cpp:40-96) distinguishes three kinds of operation: M, = and X count bases
// (count.cpp:51), I advances the read (count.cpp:74), D and N count deletions/skips (count.cpp:80); S, H, P and
// anything else are ignored (count.cpp:92-95; soft clips are already trimmed from the sequence).  The packers
// therefore hand the device a normal form with the same meaning and fewer words: =/X spelled M, N spelled D,
// ignored and zero-length operations dropped, equal neighbours merged.  A short read is then one op (M) or
// three (M, I or D, M), which is what k1_count_fast decodes in straight-line code; anything else -- including
// CIGARs that were never normalised -- takes the general walker.
#pragma once
#include <stdint.h>

namespace bccanon {

// -1 = dropped; otherwise the op code of the normal form (0 M, 1 I, 2 D)
inline int canon_op(uint32_t op)
{
    switch (op) {
    case 0: case 7: case 8: return 0;
    case 1: return 1;
    case 2: case 3: return 2;
    default: return -1;
    }
}

// Normal form of n BAM-native words read through `word(k)`; written to out (may be null) and counted.
template <class Word>
inline uint32_t canon_cigar(uint32_t n, Word word, uint32_t *out)
{
    uint32_t m = 0;
    int last_op = -1;
    uint32_t last_len = 0;
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t w = word(k), len = w >> 4;
        const int op = canon_op(w & 15u);
        if (op < 0 || len == 0) continue;
        if (op == last_op && (uint64_t)last_len + len < (1u << 28)) {
            last_len += len;
            if (out) out[m - 1] = (last_len << 4) | (uint32_t)op;
            continue;
        }
        if (out) out[m] = (len << 4) | (uint32_t)op;
        m++;
        last_op = op;
        last_len = len;
    }
    return m;
}

}  // namespace bccanon
