/*
 * utils.c -- wall clock and the host-side top-r merge of per-GPU hit lists.
 *
 * The reference sorts all n scores of a query on the host (utils.c:3-86) and prints the first `top`.
 * Here each GPU returns its `top` best hits as 64-bit keys (score << 32 | database index), already in
 * the reference's order (largest key first = score descending, then index descending), so the host
 * only merges `parts` short descending lists.
 */
#include "swimm_host.h"

#include <stdlib.h>
#include <sys/time.h>

double swg_walltime(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

/* keys: [parts][r] descending lists; out: the r largest overall, descending */
void swg_merge_top_keys(const uint64_t *keys, int parts, uint64_t r, uint64_t *out)
{
    uint64_t *head = (uint64_t *)calloc((size_t)(parts > 0 ? parts : 1), sizeof(uint64_t));
    for (uint64_t k = 0; k < r; k++) {
        int best = -1;
        uint64_t bestkey = 0;
        for (int p = 0; p < parts; p++) {
            if (head[p] < r) {
                const uint64_t v = keys[(uint64_t)p * r + head[p]];
                if (best < 0 || v > bestkey) { best = p; bestkey = v; }
            }
        }
        out[k] = bestkey;
        if (best >= 0)
            head[best]++;
    }
    free(head);
}
