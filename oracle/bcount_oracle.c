/*
 * oracle/bcount_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's native counting operator
 * (reference: basecount/count.cpp:7-99).  It exists so the CUDA path can be
 * checked bit-for-bit on the GPU box, where /root/reference is absent.
 * Nothing in basecount_b200/ may call, link or import this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (a) tests/golden/bcount_kats.json, produced by the *compiled reference*
 *       (oracle/_ref/count*.so built from /root/reference/basecount/count.cpp)
 *       via tests/golden/make_golden.py, and
 *   (b) oracle/_ref itself, live, whenever that .so is present.
 *
 * The reference takes Python lists; here the same information is laid out flat:
 *   seq / qual : the reads' query_alignment_sequence / _qualities, concatenated
 *   seq_off    : n_reads+1 offsets into seq / qual
 *   starts     : reference_start of each read
 *   cigar      : BAM-native words (len << 4 | op), concatenated
 *   cigar_off  : n_reads+1 offsets into cigar
 *   counts     : ref_len x 6 uint32, columns A,C,G,T,DS,N  (count.cpp:16-17)
 *
 * Return: 0 ok; 1 = an increment fell at refPos >= ref_len (the reference's
 * std::out_of_range from .at(), count.cpp:60-64,85 -> Python IndexError);
 * 2 = the CIGAR walks past the end of the read's sequence (undefined
 * behaviour in the reference: operator[] at count.cpp:56,58).
 */
#include <stdint.h>
#include <stddef.h>

int bcount_oracle(uint32_t ref_len, uint32_t min_base_quality, uint64_t n_reads,
                  const uint8_t *seq, const uint8_t *qual, const uint64_t *seq_off,
                  const uint32_t *starts, const uint32_t *cigar,
                  const uint64_t *cigar_off, uint32_t *counts)
{
    for (uint64_t i = 0; i < n_reads; i++) {                 /* count.cpp:22 */
        const uint8_t *rd = seq + seq_off[i];
        const uint8_t *ql = qual + seq_off[i];
        const uint64_t rd_len = seq_off[i + 1] - seq_off[i];
        uint64_t ref_pos = starts[i];                        /* count.cpp:35 */
        uint64_t read_pos = 0;                               /* count.cpp:38 */

        for (uint64_t c = cigar_off[i]; c < cigar_off[i + 1]; c++) {   /* :40 */
            const uint32_t op = cigar[c] & 0xFu;
            const uint32_t len = cigar[c] >> 4;

            if (op == 0 || op == 7 || op == 8) {             /* M,=,X  :51 */
                for (uint32_t j = 0; j < len; j++) {         /* :54 */
                    if (read_pos >= rd_len)
                        return 2;
                    if (ql[read_pos] >= min_base_quality) {  /* :56 */
                        int col = -1;
                        switch (rd[read_pos]) {              /* :58-65 */
                        case 'A': col = 0; break;
                        case 'C': col = 1; break;
                        case 'G': col = 2; break;
                        case 'T': col = 3; break;
                        case 'N': col = 5; break;
                        default: break;                      /* anything else: not counted */
                        }
                        if (col >= 0) {
                            if (ref_pos >= ref_len)
                                return 1;                    /* .at() throws */
                            counts[ref_pos * 6 + (uint64_t)col] += 1;
                        }
                    }
                    read_pos += 1;                           /* :67 */
                    ref_pos += 1;                            /* :68 */
                }
            } else if (op == 1) {                            /* I  :74-75 */
                read_pos += len;
            } else if (op == 2 || op == 3) {                 /* D,N  :80-87 */
                for (uint32_t j = 0; j < len; j++) {
                    if (ref_pos >= ref_len)
                        return 1;
                    counts[ref_pos * 6 + 4] += 1;            /* no quality test */
                    ref_pos += 1;
                }
            }
            /* ops 4,5,6,9 (S,H,P,B) ignored: count.cpp:92-95 */
        }
    }
    return 0;
}

/* "Aligned bases" as BASELINE.md defines the metric: sum of opLen over CIGAR
 * ops in {M,=,X,D,N} (iterations of count.cpp:54 plus count.cpp:83). */
uint64_t bcount_oracle_aligned_bases(uint64_t n_cigar_words, const uint32_t *cigar)
{
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_cigar_words; c++) {
        const uint32_t op = cigar[c] & 0xFu;
        if (op == 0 || op == 7 || op == 8 || op == 2 || op == 3)
            total += cigar[c] >> 4;
    }
    return total;
}
