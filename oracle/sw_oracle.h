/*
 * oracle/sw_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the SWIMM `search` hot path (see sw_oracle.c for
 * the reference file:line each function follows).  Only tests/, the smoke test
 * and bench.py's cpu_baseline / --impl reference legs may link or load this.
 * The product (libswimm_cuda.so, the swimm host binary) never does.
 *
 * Parity status: PINNED -- checked against golden vectors produced by the
 * unmodified reference binary (oracle/_ref/swimm, built by oracle/Makefile
 * from /root/reference) and committed under tests/golden/.
 */
#ifndef SW_ORACLE_H
#define SW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 'A'..'Z' -> 0..23 (J, O, U -> 23). reference sequences.c:165-175, :393-402 */
int swo_encode_residue(int c);
void swo_encode(const char *ascii, int8_t *codes, uint64_t n);

/* exact int32 local alignment score, affine gaps (gap of length k costs go + k*ge).
 * submat is the reference's 24x32 int8 table.  reference CPUsearch.c:605-667 (+ :678-956 widening) */
int32_t swo_score(const int8_t *q, uint32_t m, const int8_t *d, uint32_t n,
                  const int8_t *submat, int go, int ge);

/* all-vs-all: scores[qi*n_seqs + si]; queries/db are flat concatenations with prefix offsets */
/* EXTENSION (no reference counterpart, parity unpinned): score and 0-based inclusive coordinates
 * (q_start, q_end, d_start, d_end) of an optimal alignment; see sw_oracle.c for the tie rules. */
int32_t swo_align_ends(const int8_t *q, uint32_t m, const int8_t *d, uint32_t n, const int8_t *submat, int go, int ge,
                       int32_t *coords);

void swo_search(const int8_t *queries, const uint32_t *q_off, uint64_t q_count,
                const int8_t *db, const uint64_t *db_off, uint64_t n_seqs,
                const int8_t *submat, int go, int ge, int threads, int32_t *scores);

/* restatement of the reference's descending merge sort (utils.c:3-86); idx is permuted with scores */
void swo_sort_scores(int32_t *scores, uint64_t *idx, uint64_t n);

/* closed-form order the merge sort produces: score desc, then index desc */
void swo_top(const int32_t *scores, uint64_t n, uint64_t r, int32_t *top_scores, uint64_t *top_idx);

/* stable ascending length sort permutation (sequences.c:770-865): perm[k] = original position */
void swo_length_order(const uint16_t *lengths, uint64_t n, uint64_t *perm);

#ifdef __cplusplus
}
#endif
#endif
