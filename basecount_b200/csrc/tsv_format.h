// tsv_format.h -- exact native emitter of the per-position TSV rows.
//
// Replaces the reference's row loop (basecount/main.py:456-466): every cell is
// `x if isinstance(x, str) else str(round(x, decimal_places))`.  For the cell types the rows hold
// (main.py:55-78) that is:
//   * int cells (position, coverage, counts, and the zero-coverage sentinels -1 / 1 / 1):
//     round(int, d >= 0) is the int itself -> its decimal digits;
//   * float cells (percentages, entropies): CPython's float.__round__ rounds the EXACT binary value
//     to d decimals, ties to even (dtoa mode 3), converts back to a double and prints its shortest
//     repr.  glibc's printf("%.*f") performs the same correctly-rounded, ties-to-even conversion;
//     for 0 <= d <= 4 and |x| < 1e11 the resulting decimal has at most 15 significant digits, so it
//     round-trips and its shortest repr is that decimal with trailing zeros removed (at least one
//     fractional digit kept, fixed notation because the smallest non-zero magnitude is 1e-4).
// Anything outside that envelope (other decimal_places, huge or non-finite values) is refused and
// the caller keeps the Python path.  Host-only code, thread-parallel over row ranges.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bam_decode.h"   // bcbam::parallel_for

namespace bctsv {

inline char *put_int(char *p, long long v)
{
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// str(round(x, dp)) for a Python float, 0 <= dp <= 4, |x| < 1e11, finite.
// A double is m * 2^e exactly, so x * 10^dp = m * 10^dp / 2^-e is an exact 128-bit integer division
// by a power of two: quotient and remainder give the correctly rounded (ties-to-even) integer
// q = round(x * 10^dp) with no floating-point step at all.
inline char *put_float(char *p, double x, int dp)
{
    static const uint64_t kPow10[5] = {1ull, 10ull, 100ull, 1000ull, 10000ull};
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    if (bits >> 63) *p++ = '-';
    const int be = (int)((bits >> 52) & 0x7FF);
    uint64_t m = bits & ((1ull << 52) - 1);
    int e;                                                    // x = m * 2^e
    if (be == 0) {
        e = -1074;
    } else {
        m |= 1ull << 52;
        e = be - 1075;
    }
    unsigned __int128 q;
    if (e >= 0) {
        q = ((unsigned __int128)m << e) * kPow10[dp];         // |x| < 1e11 < 2^37: fits easily
    } else {
        const unsigned __int128 num = (unsigned __int128)m * kPow10[dp];     // < 2^67
        const int s = -e;
        if (s >= 100) {
            q = 0;                                            // |x| < 2^-46 * ... far below half a unit
        } else {
            q = num >> s;
            const unsigned __int128 rem = num - (q << s), half = (unsigned __int128)1 << (s - 1);
            if (rem > half || (rem == half && (q & 1))) q++;
        }
    }
    uint64_t ip = (uint64_t)(q / kPow10[dp]), fp = (uint64_t)(q % kPow10[dp]);
    p = put_int(p, (long long)ip);
    *p++ = '.';
    if (dp == 0 || fp == 0) {
        *p++ = '0';
        return p;
    }
    char d[4];
    for (int i = dp - 1; i >= 0; i--) {
        d[i] = (char)('0' + fp % 10);
        fp /= 10;
    }
    int n = dp;
    while (n > 1 && d[n - 1] == '0') n--;
    for (int i = 0; i < n; i++) *p++ = d[i];
    return p;
}

}  // namespace bctsv

// counts: n_pos x 6 int64 row-major (A,C,G,T,DS,N); pc: K planes of `pc_stride` doubles; flags: bit0 =
// coverage 0 (pc = int -1, entropy = secondary = int 1), bit1 = secondary coverage 0 (secondary = int 1).
// Returns 0 and a malloc'ed text (rows joined by '\n', no trailing newline), or 1 if a value is
// outside the exactness envelope.
inline int bc_format_tsv_impl(const char *ref_name, uint64_t n_pos, uint64_t first_pos, int K, int long_format,
                              int decimal_places, const int64_t *counts, const int64_t *coverage, const double *pc,
                              uint64_t pc_stride, const double *entropy, const double *secondary, const uint8_t *flags,
                              int threads, char **text, uint64_t *len)
{
    static const char *kBase[6] = {"A", "C", "G", "T", "DS", "N"};
    if (decimal_places < 0 || decimal_places > 4 || K < 1 || K > 6) return 1;
    const size_t name_len = std::strlen(ref_name);
    const uint64_t grain = 1 << 14;
    const uint64_t chunks = (n_pos + grain - 1) / grain;
    std::vector<std::string> parts(chunks);
    std::vector<int> bad(chunks, 0);
    if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    bcbam::parallel_for(threads, n_pos, grain, [&](uint64_t a, uint64_t e) {
        std::string &out = parts[a / grain];
        const size_t per_row = name_len + 24 * (size_t)(3 + 2 * K + 2) + 8;
        out.resize((size_t)(e - a) * per_row * (long_format ? (size_t)K : 1u));
        char *p = &out[0];
        for (uint64_t i = a; i < e; i++) {
            const uint8_t f = flags[i];
            double pcv[6];
            bool okv = std::isfinite(entropy[i]) && std::fabs(entropy[i]) < 1e11 && std::isfinite(secondary[i]) &&
                       std::fabs(secondary[i]) < 1e11;
            for (int k = 0; k < K; k++) {
                pcv[k] = pc[(uint64_t)k * pc_stride + i];
                okv = okv && std::isfinite(pcv[k]) && std::fabs(pcv[k]) < 1e11;
            }
            if (!okv) {
                bad[a / grain] = 1;
                return;
            }
            const int rows = long_format ? K : 1;
            for (int r = 0; r < rows; r++) {
                if (i != a || r != 0) *p++ = '\n';
                std::memcpy(p, ref_name, name_len);
                p += name_len;
                *p++ = '\t';
                p = bctsv::put_int(p, (long long)(first_pos + i));
                *p++ = '\t';
                p = bctsv::put_int(p, (long long)coverage[i]);
                if (long_format) {                                   // main.py:57-68
                    *p++ = '\t';
                    for (const char *c = kBase[r]; *c; c++) *p++ = *c;
                    *p++ = '\t';
                    p = bctsv::put_int(p, (long long)counts[i * 6 + r]);
                    *p++ = '\t';
                    if (f & 1) { *p++ = '-'; *p++ = '1'; } else p = bctsv::put_float(p, pcv[r], decimal_places);
                } else {                                             // main.py:70-78
                    for (int k = 0; k < K; k++) {
                        *p++ = '\t';
                        p = bctsv::put_int(p, (long long)counts[i * 6 + k]);
                    }
                    for (int k = 0; k < K; k++) {
                        *p++ = '\t';
                        if (f & 1) { *p++ = '-'; *p++ = '1'; } else p = bctsv::put_float(p, pcv[k], decimal_places);
                    }
                }
                *p++ = '\t';
                if (f & 1) *p++ = '1'; else p = bctsv::put_float(p, entropy[i], decimal_places);
                *p++ = '\t';
                if (f & 3) *p++ = '1'; else p = bctsv::put_float(p, secondary[i], decimal_places);
            }
        }
        out.resize((size_t)(p - &out[0]));
    });
    uint64_t total = 0;
    for (uint64_t c = 0; c < chunks; c++) {
        if (bad[c]) return 1;
        total += parts[c].size() + (c ? 1 : 0);
    }
    char *buf = (char *)std::malloc(total + 1);
    if (!buf) return 1;
    char *p = buf;
    for (uint64_t c = 0; c < chunks; c++) {
        if (c) *p++ = '\n';
        std::memcpy(p, parts[c].data(), parts[c].size());
        p += parts[c].size();
    }
    *p = 0;
    *text = buf;
    *len = total;
    return 0;
}
