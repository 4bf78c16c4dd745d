/*
 * oracle/sw_oracle.c -- TEST INFRASTRUCTURE ONLY (see sw_oracle.h).
 *
 * Plain scalar C restatement of what SWIMM's `-S search -m 0` path computes.
 * It is deliberately the simplest correct statement of the arithmetic (one
 * int32 cell at a time), not a transcription of the SIMD code; the reference
 * lines each function is answerable to are cited at the function.
 *
 * Parity: PINNED by tests/test_oracle_golden.py against the JSON fixtures in tests/golden,
 * which were produced by the unmodified reference binary (oracle/_ref/swimm).
 */
#include "sw_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Residue re-encoding.
 * reference sequences.c:165-175 (database) and :393-402 (queries): J, O and U
 * become the dummy symbol 'Z'+1, then the letter is shifted down by 'A' plus the
 * number of removed letters that precede it, giving the 23-letter alphabet
 * A B C D E F G H I K L M N P Q R S T V W X Y Z -> 0..22 and dummy -> 23. */
int swo_encode_residue(int c)
{
    if (c == 'J' || c == 'O' || c == 'U')
        c = 'Z' + 1;
    return c - ('A' + (c > 'J') + (c > 'O') + (c > 'U'));
}

void swo_encode(const char *ascii, int8_t *codes, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++)
        codes[i] = (int8_t)swo_encode_residue((unsigned char)ascii[i]);
}

/* One (query, database sequence) score.
 *
 * reference CPUsearch.c:605-655 is the recurrence (per lane):
 *     cur = max(0, H[i-1][j-1] + SP, maxRow(i), maxCol(j))
 *     maxRow(i) = max(maxRow(i) - ge, cur - (go+ge))      horizontal gap state
 *     maxCol(j) = max(maxCol(j) - ge, cur - (go+ge))      vertical gap state
 *     score = max(score, cur)
 * with both gap states starting at 0, which under the 0 floor is the same as
 * minus infinity.  The 8 -> 16 -> 32 bit re-computation (:678-956) makes the
 * stored value the exact integer, so the oracle simply works in int32.
 * Substitution lookup: submat[q*32 + d] (profile build :581-603); the profile
 * row of query code 23 is forced to zero (:602) and table columns 23..31 are
 * zero, so dummy/padding residues score 0 against everything.
 */
int32_t swo_score(const int8_t *q, uint32_t m, const int8_t *d, uint32_t n,
                  const int8_t *submat, int go, int ge)
{
    if (m == 0 || n == 0)
        return 0;
    const int32_t goe = go + ge;
    int32_t *hrow = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t)); /* H[i-1][*] */
    int32_t *vgap = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t)); /* maxCol */
    int32_t best = 0;
    for (uint32_t i = 0; i < m; i++) {
        const int qi = q[i];
        const int8_t *row = (qi >= 0 && qi < 23) ? submat + 32 * qi : NULL; /* row 23 == all zero */
        int32_t hgap = 0;      /* maxRow for this query row */
        int32_t diag = 0;      /* H[i-1][j-1] */
        for (uint32_t j = 1; j <= n; j++) {
            const int dj = d[j - 1];
            const int32_t s = (row && dj >= 0 && dj < 32) ? row[dj] : 0;
            int32_t cur = diag + s;
            if (cur < hgap) cur = hgap;
            if (cur < vgap[j]) cur = vgap[j];
            if (cur < 0) cur = 0;
            const int32_t open = cur - goe;
            hgap -= ge;    if (hgap < open) hgap = open;
            int32_t v = vgap[j] - ge; if (v < open) v = open;
            vgap[j] = v;
            diag = hrow[j];
            hrow[j] = cur;
            if (best < cur) best = cur;
        }
    }
    free(hrow);
    free(vgap);
    return best;
}

/* ---- start / end coordinates of an optimal alignment (EXTENSION: the reference is score-only, CPUsearch.c:670-676) ----
 * PARITY UNPINNED for this function: there is no reference output to pin it to.  It restates, cell by cell, the
 * definition the product's opt-in coordinate pass (csrc/align_ends.cu) must reproduce:
 *   end   = the cell (i, j) of the forward recurrence above (same scores as swo_score) that holds the best score;
 *           ties: the smallest database position j, then the smallest query position i;
 *   start = the same search on the REVERSED prefixes q[0..i], d[0..j] (an optimal local alignment of the reversed
 *           prefixes has the same score and ends where an optimal alignment of the originals starts), mapped back.
 * Positions are 0-based and inclusive.  A score of 0 gives all -1.
 * best_cell: score and argmax cell of x[0..m) vs y[0..n) read forwards (step = +1) or backwards from (x0, y0). */
static int32_t best_cell(const int8_t *q, int64_t q0, int qstep, uint32_t m, const int8_t *d, int64_t d0, int dstep, uint32_t n,
                         const int8_t *submat, int go, int ge, int64_t *bi, int64_t *bj)
{
    *bi = -1;
    *bj = -1;
    if (m == 0 || n == 0)
        return 0;
    const int32_t goe = go + ge;
    int32_t *hrow = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *vgap = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t best = 0;
    for (uint32_t i = 0; i < m; i++) {
        const int qi = q[q0 + (int64_t)qstep * i];
        const int8_t *row = (qi >= 0 && qi < 23) ? submat + 32 * qi : NULL;
        int32_t hgap = 0, diag = 0;
        for (uint32_t j = 1; j <= n; j++) {
            const int dj = d[d0 + (int64_t)dstep * (j - 1)];
            const int32_t s = (row && dj >= 0 && dj < 32) ? row[dj] : 0;
            int32_t cur = diag + s;
            if (cur < hgap) cur = hgap;
            if (cur < vgap[j]) cur = vgap[j];
            if (cur < 0) cur = 0;
            const int32_t open = cur - goe;
            hgap -= ge;    if (hgap < open) hgap = open;
            int32_t v = vgap[j] - ge; if (v < open) v = open;
            vgap[j] = v;
            diag = hrow[j];
            hrow[j] = cur;
            if (cur > best || (cur == best && cur > 0 && (int64_t)(j - 1) < *bj)) {
                best = cur;
                *bi = i;
                *bj = j - 1;
            }
        }
    }
    free(hrow);
    free(vgap);
    return best;
}

int32_t swo_align_ends(const int8_t *q, uint32_t m, const int8_t *d, uint32_t n, const int8_t *submat, int go, int ge,
                       int32_t *coords /* q_start, q_end, d_start, d_end */)
{
    int64_t ei, ej, si, sj;
    const int32_t score = best_cell(q, 0, 1, m, d, 0, 1, n, submat, go, ge, &ei, &ej);
    coords[0] = coords[1] = coords[2] = coords[3] = -1;
    if (score <= 0)
        return 0;
    const int32_t back = best_cell(q, ei, -1, (uint32_t)ei + 1, d, ej, -1, (uint32_t)ej + 1, submat, go, ge, &si, &sj);
    if (back != score)
        return -1;              /* cannot happen: the reversed prefixes hold the same optimal alignment */
    coords[0] = (int32_t)(ei - si);
    coords[1] = (int32_t)ei;
    coords[2] = (int32_t)(ej - sj);
    coords[3] = (int32_t)ej;
    return score;
}

/* Every query against every database sequence.
 * reference CPUsearch.c:540-548: tasks are (query, lane-group); the score of
 * sequence s for query q lands at scores[q*N + s] with s the position in the
 * length-sorted database. */
void swo_search(const int8_t *queries, const uint32_t *q_off, uint64_t q_count,
                const int8_t *db, const uint64_t *db_off, uint64_t n_seqs,
                const int8_t *submat, int go, int ge, int threads, int32_t *scores)
{
    const int64_t total = (int64_t)(q_count * n_seqs);
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
#endif
    for (int64_t t = 0; t < total; t++) {
        /* longest first, like the reference's reversed task index (:543-544) */
        const uint64_t u = (uint64_t)(total - 1 - t);
        const uint64_t qi = u % q_count, si = u / q_count;
        scores[qi * n_seqs + si] =
            swo_score(queries + q_off[qi], q_off[qi + 1] - q_off[qi],
                      db + db_off[si], (uint32_t)(db_off[si + 1] - db_off[si]), submat, go, ge);
    }
}

/* Descending merge sort exactly as the reference performs it.
 * reference utils.c:44-68 (recursion: halves of size/2 and size-size/2, a pair
 * is swapped when first <= second) and utils.c:3-41 (merge: the left element is
 * taken only when strictly greater).  The threaded wrapper (:71-86) splits at
 * the same size/2 points, so the result does not depend on the thread count. */
static void merge_halves(int32_t *s, uint64_t *ix, uint64_t n, int32_t *ts, uint64_t *ti)
{
    const uint64_t half = n / 2;
    uint64_t a = 0, b = half, o = 0;
    while (a < half && b < n) {
        if (s[a] > s[b]) { ts[o] = s[a]; ti[o] = ix[a]; a++; }
        else             { ts[o] = s[b]; ti[o] = ix[b]; b++; }
        o++;
    }
    for (; a < half; a++, o++) { ts[o] = s[a]; ti[o] = ix[a]; }
    for (; b < n; b++, o++)    { ts[o] = s[b]; ti[o] = ix[b]; }
    memcpy(s, ts, n * sizeof(int32_t));
    memcpy(ix, ti, n * sizeof(uint64_t));
}

static void sort_rec(int32_t *s, uint64_t *ix, uint64_t n, int32_t *ts, uint64_t *ti)
{
    if (n == 2) {
        if (s[0] <= s[1]) {
            int32_t a = s[0]; s[0] = s[1]; s[1] = a;
            uint64_t b = ix[0]; ix[0] = ix[1]; ix[1] = b;
        }
    } else if (n > 2) {
        sort_rec(s, ix, n / 2, ts, ti);
        sort_rec(s + n / 2, ix + n / 2, n - n / 2, ts, ti);
        merge_halves(s, ix, n, ts, ti);
    }
}

void swo_sort_scores(int32_t *scores, uint64_t *idx, uint64_t n)
{
    if (n < 2) return;
    int32_t *ts = (int32_t *)malloc(n * sizeof(int32_t));
    uint64_t *ti = (uint64_t *)malloc(n * sizeof(uint64_t));
    sort_rec(scores, idx, n, ts, ti);
    free(ts);
    free(ti);
}

/* Closed form of the order above: score descending, ties by database index
 * descending.  tests check it equals swo_sort_scores. */
typedef struct { int32_t s; uint64_t i; } hit_t;

static int hit_cmp(const void *pa, const void *pb)
{
    const hit_t *a = (const hit_t *)pa, *b = (const hit_t *)pb;
    if (a->s != b->s) return a->s > b->s ? -1 : 1;
    if (a->i != b->i) return a->i > b->i ? -1 : 1;
    return 0;
}

void swo_top(const int32_t *scores, uint64_t n, uint64_t r, int32_t *top_scores, uint64_t *top_idx)
{
    hit_t *h = (hit_t *)malloc((n ? n : 1) * sizeof(hit_t));
    for (uint64_t i = 0; i < n; i++) { h[i].s = scores[i]; h[i].i = i; }
    qsort(h, n, sizeof(hit_t), hit_cmp);
    if (r > n) r = n;
    for (uint64_t k = 0; k < r; k++) { top_scores[k] = h[k].s; top_idx[k] = h[k].i; }
    free(h);
}

/* Stable ascending order of sequence lengths.
 * reference sequences.c:770-865: merge takes the left element on <= (:780) and a
 * pair is swapped only on > (:826), i.e. a stable ascending merge sort. */
void swo_length_order(const uint16_t *lengths, uint64_t n, uint64_t *perm)
{
    uint64_t *tmp = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) perm[i] = i;
    for (uint64_t w = 1; w < n; w *= 2) {
        for (uint64_t lo = 0; lo < n; lo += 2 * w) {
            uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            uint64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi)
                tmp[o++] = (lengths[perm[a]] <= lengths[perm[b]]) ? perm[a++] : perm[b++];
            while (a < mid) tmp[o++] = perm[a++];
            while (b < hi) tmp[o++] = perm[b++];
        }
        memcpy(perm, tmp, n * sizeof(uint64_t));
    }
    free(tmp);
}
