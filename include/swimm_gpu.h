/*
 * swimm_gpu.h -- C ABI of libswimm_cuda.so, the B200 (sm_100a) replacement for SWIMM's search kernels.
 *
 * What it replaces in the reference (enzorucci/SWIMM v1.1.3):
 *   cpu_search_avx2_sp() / cpu_search_sse_sp()   CPUsearch.h:32-39, CPUsearch.c:6-967   (called from swimm.c:66-76)
 *   assemble_single_chunk_db()'s lane interleave sequences.c:618-734                    (now a device-side layout build)
 *   sort_scores() + the top-r print loop         utils.c:3-86, swimm.c:150-160          (now a device-side top-r)
 *
 * Conventions: plain C, plain pointers and sizes, caller owns every host buffer, the library owns
 * all device memory.  Every function returns 0 on success and a negative swg_status otherwise; the
 * library never calls exit().  There is NO CPU fallback: without a CUDA device every entry point
 * that computes returns SWG_ERR_NO_DEVICE / SWG_ERR_CUDA.
 *
 * Residues are the reference's preprocessed codes (sequences.c:165-175): 0..22 = ABCDEFGHIKLMNPQRSTVWXYZ,
 * 23 = dummy (J/O/U).  Database arrays are exactly the contents of <db>.seq: `lengths` ascending
 * (stable), `residues` concatenated in that order.  A database index in any result is the position
 * in that length-sorted order, i.e. the line number in <db>.desc -- the same index the reference's
 * `scores` array uses (CPUsearch.c:548).
 */
#ifndef SWIMM_GPU_H
#define SWIMM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swg_ctx swg_ctx;

typedef enum {
    SWG_OK = 0,
    SWG_ERR_NO_DEVICE = -1,
    SWG_ERR_CUDA = -2,
    SWG_ERR_ARG = -3,
    SWG_ERR_STATE = -4,
    SWG_ERR_NOMEM = -5
} swg_status;

/* key of one hit: (uint32 score << 32) | database index; larger key = earlier in the hit list.
 * This reproduces the reference order exactly: score descending, then index descending
 * (utils.c:12 strict '>' in the merge, utils.c:52 '<=' swap). */
#define SWG_KEY(score, index) (((uint64_t)(uint32_t)(score) << 32) | (uint64_t)(uint32_t)(index))
#define SWG_KEY_SCORE(key)    ((int32_t)((key) >> 32))
#define SWG_KEY_INDEX(key)    ((uint64_t)((key) & 0xffffffffu))

typedef struct {
    double device_seconds;     /* CUDA-event time of the last run: profile build + all search kernels (+ top-r) */
    double search_seconds;     /* the part of it spent in the alignment kernels (the reference's workTime region) */
    double topr_seconds;       /* the part spent in top-r selection */
    uint64_t cells;            /* sum over queries of query_length * local database residues */
    uint64_t padded_cells;     /* cell updates actually executed (tile padding, pipeline fill, lane pairing) */
    uint64_t launches;         /* kernels launched by the last run */
    uint64_t rescored;         /* (query, sequence) pairs that left the 16-bit range and were redone in 32 bits */
    uint64_t db_bytes;         /* bytes of the resident tiled database on this device */
    uint64_t h2d_bytes;        /* host->device bytes of the last search call */
    uint64_t d2h_bytes;        /* device->host bytes of the last search call */
    uint64_t pair_launches;    /* launches of the query-pair kernel in the last run (one per pass of a query pair) */
    uint64_t stream_bytes;     /* algorithmic HBM bytes of the last run's search launches: the tiled database once per
                                  launch + the pass lines the query-pair kernel parks and re-reads (8 B per column each way) */
} swg_stats;

/* ---- context ---- */
int swg_gpu_device_count(int *count);
int swg_gpu_create(int device, swg_ctx **ctx);                 /* one context per GPU (per process or per host thread) */
void swg_gpu_destroy(swg_ctx *ctx);
const char *swg_gpu_last_error(const swg_ctx *ctx);           /* static string when ctx == NULL */

/* ---- database (replaces assemble_single_chunk_db, sequences.c:618-734) ----
 * Uploads the flat database and builds the tiled, pair-interleaved device layout on the GPU.
 * shard/num_shards: this context keeps tiles (16 consecutive sorted sequences) t with
 * t % num_shards == shard, so every shard gets the same residue count and length mix. */
int swg_gpu_load_db(swg_ctx *ctx, const uint16_t *lengths, const signed char *residues,
                    uint64_t n_sequences, uint64_t n_residues, int shard, int num_shards);
/* same, when the caller already holds the prefix sums of the lengths (offsets[i] = first residue of sequence i,
 * offsets[n_sequences] = n_residues; what the host keeps after reading <db>.seq): a shard then visits only its own
 * tiles instead of walking the whole length array -- one host thread per GPU can load all shards at the same time. */
int swg_gpu_load_db_offsets(swg_ctx *ctx, const uint16_t *lengths, const uint64_t *offsets, const signed char *residues,
                            uint64_t n_sequences, uint64_t n_residues, int shard, int num_shards);
/* same, when the caller holds only this shard's sequences (tiles shard, shard+num_shards, ... of the whole
 * length-sorted database, back to back): what one process per GPU loads in a multi-GPU job. */
int swg_gpu_load_db_shard(swg_ctx *ctx, const uint16_t *local_lengths, const signed char *local_residues,
                          uint64_t n_local, uint64_t n_local_residues, int shard, int num_shards, uint64_t n_total);
/* same, from the reference's 32- or 16-lane interleaved arrays (what swimm.c:74-76 hands to
 * cpu_search_avx2_sp): vect_db[disp[g] + j*vector_length + k], padded with code 24. */
int swg_gpu_load_db_interleaved(swg_ctx *ctx, const signed char *vect_db, const uint16_t *vect_lengths,
                                uint64_t vect_count, const uint64_t *vect_disp, int vector_length,
                                uint64_t n_sequences, int shard, int num_shards);
uint64_t swg_gpu_db_local_sequences(const swg_ctx *ctx);
uint64_t swg_gpu_db_local_residues(const swg_ctx *ctx);

/* ---- search (replaces cpu_search_avx2_sp + sort_scores) ----
 * queries:    residue codes, concatenated; query i is queries[q_disp[i] .. q_disp[i] + q_lengths[i])
 * q_lengths:  REAL lengths (no even padding needed; a trailing dummy residue is harmless)
 * submat:     24x32 signed bytes, row = query code, column = database code (reference submat.c)
 * top:        hits per query wanted; only min(top, n_sequences) exist (reference swimm.c:51)
 * scores:     NULL, or [q_count][n_sequences] int32 -- entries of sequences this shard holds are written
 * top_keys:   NULL, or [q_count][top] SWG_KEYs of this shard's best hits, descending, rows padded with 0
 * work_seconds: NULL, or receives swg_stats.search_seconds (the reference's *workTime) */
int swg_gpu_search(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                   uint64_t q_count, const signed char *submat, int open_gap, int extend_gap, uint64_t top,
                   int32_t *scores, uint64_t *top_keys, double *work_seconds);

/* the same call in three steps, so that several GPUs can be driven from one host thread and so
 * that the kernels can be timed with the inputs already resident in HBM */
int swg_gpu_set_queries(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                        uint64_t q_count, const signed char *submat, int open_gap, int extend_gap);   /* H2D */
/* keep_scores != 0: every query's score row stays on the device for swg_gpu_fetch(scores != NULL).  0: only hit lists
 * are wanted; the run may then search the batch in chunks of queries so that its score rows fit option
 * "score_budget_mb" (fetching scores after such a run returns SWG_ERR_STATE). */
int swg_gpu_run(swg_ctx *ctx, uint64_t top, int keep_scores);     /* enqueue all kernels; returns immediately */
int swg_gpu_fetch(swg_ctx *ctx, int32_t *scores, uint64_t *top_keys);   /* wait + D2H */
int swg_gpu_sync(swg_ctx *ctx);                                   /* wait only */

/* ---- opt-in: alignment coordinates of the hits (no counterpart in the reference, which is score-only:
 * CPUsearch.c:670-676; definition and tie rules: oracle/sw_oracle.c swo_align_ends) ----
 * After swg_gpu_run / swg_gpu_fetch of the current queries with top > 0: coords[q][k][0..3] = q_start, q_end, d_start,
 * d_end (0-based, inclusive) of an optimal alignment behind hit k of query q, in the order of swg_gpu_fetch's top_keys
 * (this shard's hits); rows are `top` entries apart, entries without a hit hold -1.  Only the hits are aligned again
 * (one warp each), so the cost is top alignments per query, not n. */
int swg_gpu_align_ends(swg_ctx *ctx, int32_t *coords);

/* ---- streaming: keep the database resident and feed batches as they arrive (replaces the one-shot flow of
 * swimm.c:38-160 for a long-lived process) ----
 * swg_gpu_submit enqueues the upload of a batch (copy stream), all its kernels and the download of its hit lists, and
 * returns at once with a ticket.  Two batches may be in flight: while batch k computes, batch k+1 is uploaded and its
 * launches queue up behind, so the GPU does not idle between batches.  swg_gpu_poll hands a finished batch's hit lists
 * out (wait != 0: blocks until it is finished); *done = 1 when they were delivered.  top_keys: [q_count][top] as in
 * swg_gpu_search; device_seconds (may be NULL): CUDA-event time from the batch's first kernel to its last download. */
int swg_gpu_submit(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                   uint64_t q_count, const signed char *submat, int open_gap, int extend_gap, uint64_t top, int *ticket);
int swg_gpu_poll(swg_ctx *ctx, int ticket, int wait, uint64_t *top_keys, double *device_seconds, int *done);

int swg_gpu_get_stats(swg_ctx *ctx, swg_stats *out);
/* device time (CUDA events) of each query's kernels in the last completed run, in the order the queries were given */
int swg_gpu_get_query_seconds(swg_ctx *ctx, double *seconds, uint64_t max_queries);

/* which kernel searched each query in the last run, in the order the queries were given: 0 = sequence-pair kernel
 * (one query, two database sequences per register), 1 = query-pair kernel (two queries of a batch per register) */
int swg_gpu_get_query_kernels(swg_ctx *ctx, int32_t *kind, uint64_t max_queries);

/* The schedule swg_gpu_run would use for a batch of queries with these lengths on a shard of n_sequences sequences /
 * n_residues residues whose longest sequence has longest_sequence residues, as text (one line per query or launch,
 * the format of option "verbose").  Pure host code: works without a GPU.  query_pairing as the option (0/1/2).
 * At most capacity - 1 characters are written (NUL-terminated); returns SWG_OK or SWG_ERR_ARG. */
int swg_plan_describe(const uint16_t *q_lengths, uint64_t q_count, uint64_t n_sequences, uint64_t n_residues,
                      uint32_t longest_sequence, int query_pairing, char *text, uint64_t capacity);

/* The column chunks swg_gpu_run cuts long tiles into for a query of m <= 1024 rows (pure host code, works without a
 * GPU; what the exactness argument rests on is tested through this entry): tile_cols[n_tiles] = columns of every long
 * tile (multiples of 8), smax = largest entry of the substitution matrix, option as "chunk_columns".  chunks receives
 * up to `capacity` triples (tile, first column, columns), *n_chunks their number (0: no chunking), *span_bound the
 * largest number of columns an alignment of the query can span. */
int swg_plan_column_chunks(uint32_t m, int smax, int extend_gap, long option, const uint32_t *tile_cols, uint32_t n_tiles,
                           uint32_t *chunks, uint64_t capacity, uint64_t *n_chunks, uint64_t *span_bound);

/* Measured issue rates of the search kernel's instruction mix (the integer roofline the search is
 * reported against).  ginstr_per_s[p] = 1e9 thread-instructions per second on the whole GPU for probe p,
 * sm_mhz[p] = the SM clock during it, names[p] = static strings.  max_probes >= 64 is enough. */
int swg_gpu_pipebench(swg_ctx *ctx, int max_probes, double *ginstr_per_s, double *sm_mhz, const char **names,
                      int *n_probes, int *sm_count);

/* diagnostics: copy an internal device buffer to the host.  name in {"db", "tile_off", "tile_cols", "profile",
 * "profile32", "scores", "counters"}; *bytes receives the buffer's size, at most max_bytes are copied. */
int swg_gpu_debug_read(swg_ctx *ctx, const char *name, void *out, uint64_t max_bytes, uint64_t *bytes);

/* tuning knobs (all optional): name in {"long_threshold" (0 default: the long-tile threshold is estimated per query
 * from the run times; > 0: tiles with more columns than this run the long-sequence kernel), "long_kernel" (1 default:
 * long tiles run the cross-warp wavefront kernel; 0: the 32-thread shape of the sequence-pair kernel), "xw_warps",
 * "xw_rows" (forced shape of the long-sequence kernel), "score_budget_mb" (bytes of score rows a run may hold when only
 * hit lists are wanted, default 4096: larger batches are searched in chunks of queries), "force_group", "force_rows", "block_threads",
 * "query_pairing" (0 never / 1 planner / 2 always pair the queries of a batch for the query-pair kernel),
 * "q2_group", "q2_rows" (forced shape of the query-pair kernel), "grid_blocks" (CTAs per search launch, 0 = one per SM),
 * "pass_lines" (1 default; 0 = never allocate the query-pair kernel's pass lines, 8 bytes per database column:
 * only single-launch pairs are formed, longer queries run the sequence-pair kernel), "verbose" (1 = print every
 * run's schedule to stderr)} */
int swg_gpu_set_option(swg_ctx *ctx, const char *name, long value);

/* Drop-in with the reference signature (CPUsearch.h:37-39).  n_threads is read as the number of GPUs
 * to use (0 or less = all visible; more than visible = all visible): the database is sharded over them by tiles and
 * every GPU fills its own entries of `scores`.  cpu_block_size is ignored.  Fills scores[q*vect_count*vector_length + s] for every
 * lane (padded lanes get 0) and *workTime, exactly as cpu_search_avx2_sp does.  vector_length is 32. */
int swimm_gpu_search_avx2_compat(char *query_sequences, unsigned short int *query_sequences_lengths,
                                 unsigned long int query_sequences_count, unsigned int *query_disp,
                                 char *vect_sequences_db, unsigned short int *vect_sequences_db_lengths,
                                 unsigned short int *vect_sequences_db_blocks, unsigned long int vect_sequences_db_count,
                                 unsigned long int *vect_sequences_db_disp, char *submat, int open_gap, int extend_gap,
                                 int n_threads, int cpu_block_size, int *scores, double *workTime);

#ifdef __cplusplus
}
#endif
#endif
