/*
 * swimm.c -- driver of the B200 swimm build: `swimm -S preprocess ...` and `swimm -S search ... -m 3`.
 *
 * Mirrors the reference driver (swimm.c:9-207): parse arguments, preprocess OR (load queries, load the
 * database, search, load headers, print the top hits of every query, print time / GCUPS / mode), with
 * the same stdout layout.  The search itself is the C ABI of libswimm_cuda.so (include/swimm_gpu.h).
 *
 * With -x N the database is sharded over N GPUs: one host thread per GPU uploads its shard (the reference's
 * one-thread-per-device pattern, MICsearch.c:53,74-75), every GPU sees all queries, and the per-GPU hit lists
 * are merged here.
 *
 * Streaming: `-q a,b,c` or `-q -` (file names on stdin, one per line) keeps the process and the database
 * resident and feeds the files as batches through swg_gpu_submit / swg_gpu_poll: while batch k computes,
 * file k+1 is parsed, uploaded and queued behind it.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/swimm_gpu.h"
#include "arguments.h"
#include "swimm_host.h"

static void die_gpu(const char *what, swg_ctx *ctx, int st)
{
    printf("SWIMM: %s failed (%d): %s\n", what, st, swg_gpu_last_error(ctx));
    exit(4);
}

/* ---- one host thread per GPU: context + shard upload ---- */
typedef struct {
    int gpu, ngpu;
    const swg_seqset *db;
    swg_ctx *ctx;
    int st;
    char err[512];
} load_job;

static void *load_thread(void *arg)
{
    load_job *j = (load_job *)arg;
    j->st = swg_gpu_create(j->gpu, &j->ctx);
    if (j->st == SWG_OK)
        j->st = swg_gpu_load_db_offsets(j->ctx, j->db->lengths, j->db->offsets, j->db->codes, j->db->count, j->db->residues,
                                        j->gpu, j->ngpu);
    if (j->st != SWG_OK)
        snprintf(j->err, sizeof(j->err), "%s", swg_gpu_last_error(j->ctx));
    return NULL;
}

/* ---- one batch = one query file in flight ---- */
typedef struct {
    char *name;
    swg_seqset q;
    uint32_t *q_disp;
    int *tickets;            /* one per GPU */
    time_t when;
    double t_submit;
    /* --coordinates: the batch is searched with the blocking calls and keeps its per-GPU results here */
    uint64_t *part_keys;     /* [ngpu][q][top] */
    int32_t *part_coords;    /* [ngpu][q][top][4] */
    double work;
} batch;

static void batch_free(batch *b)
{
    free(b->name);
    free(b->q_disp);
    free(b->tickets);
    free(b->part_keys);
    free(b->part_coords);
    swg_seqset_free(&b->q);
    memset(b, 0, sizeof(*b));
}

typedef struct {
    batch *b;
    swg_ctx *ctx;
    const swg_options *opt;
    unsigned long top;
    int *ticket;
    int st;
} submit_job;

static void *submit_thread(void *arg)
{
    submit_job *j = (submit_job *)arg;
    j->st = swg_gpu_submit(j->ctx, j->b->q.codes, j->b->q.lengths, j->b->q_disp, j->b->q.count, swg_submat_table(j->opt->submat),
                           j->opt->open_gap, j->opt->extend_gap, j->top, j->ticket);
    return NULL;
}

/* parse a query file and queue it on every GPU; returns 0, or the reference's exit code */
static int batch_submit(batch *b, const char *qfile, swg_ctx **ctx, int ngpu, const swg_options *opt, unsigned long top)
{
    memset(b, 0, sizeof(*b));
    /* queries: parsed, sorted by ascending length, encoded (reference sequences.c:223-423) */
    if (swg_read_fasta(qfile, &b->q) != 0) {
        printf("SWIMM: An error occurred while opening input sequence file.\n");
        return 2;
    }
    b->name = strdup(qfile);
    b->when = time(NULL);
    b->q_disp = (uint32_t *)malloc((b->q.count + 1) * sizeof(uint32_t));
    b->tickets = (int *)malloc((size_t)ngpu * sizeof(int));
    for (uint64_t i = 0; i <= b->q.count; i++)
        b->q_disp[i] = (uint32_t)b->q.offsets[i];
    b->t_submit = swg_walltime();
    if (opt->coordinates) {
        /* opt-in coordinate pass: blocking calls (search on every GPU first, then results + coordinates per GPU) */
        const uint64_t nq = b->q.count, t1 = top ? top : 1;
        b->part_keys = (uint64_t *)calloc((size_t)ngpu * nq * t1, sizeof(uint64_t));
        b->part_coords = (int32_t *)calloc((size_t)ngpu * nq * t1 * 4, sizeof(int32_t));
        for (int g = 0; g < ngpu; g++) {
            int st = swg_gpu_set_queries(ctx[g], b->q.codes, b->q.lengths, b->q_disp, nq, swg_submat_table(opt->submat),
                                         opt->open_gap, opt->extend_gap);
            if (st == SWG_OK)
                st = swg_gpu_run(ctx[g], top, 0);
            if (st != SWG_OK)
                die_gpu("search", ctx[g], st);
        }
        for (int g = 0; g < ngpu; g++) {
            int st = swg_gpu_fetch(ctx[g], NULL, b->part_keys + (size_t)g * nq * top);
            swg_stats ss;
            swg_gpu_get_stats(ctx[g], &ss);
            if (ss.device_seconds > b->work)
                b->work = ss.device_seconds;
            if (st == SWG_OK)
                st = swg_gpu_align_ends(ctx[g], b->part_coords + (size_t)g * nq * top * 4);
            if (st != SWG_OK)
                die_gpu("result download", ctx[g], st);
        }
        return 0;
    }
    /* every GPU gets all queries and searches its shard of the database; one host thread per GPU queues the batch
     * (planning, the first call's one-off loading of the kernels' code) so that no device waits for another's host work */
    submit_job *jobs = (submit_job *)calloc((size_t)ngpu, sizeof(submit_job));
    pthread_t *threads = (pthread_t *)calloc((size_t)ngpu, sizeof(pthread_t));
    for (int g = 0; g < ngpu; g++) {
        jobs[g].b = b;
        jobs[g].ctx = ctx[g];
        jobs[g].opt = opt;
        jobs[g].top = top;
        jobs[g].ticket = &b->tickets[g];
        if (ngpu == 1 || pthread_create(&threads[g], NULL, submit_thread, &jobs[g]) != 0) {
            submit_thread(&jobs[g]);
            threads[g] = 0;
        }
    }
    for (int g = 0; g < ngpu; g++) {
        if (threads[g])
            pthread_join(threads[g], NULL);
        if (jobs[g].st != SWG_OK)
            die_gpu("search", ctx[g], jobs[g].st);
    }
    free(jobs);
    free(threads);
    return 0;
}

/* wait for a batch, merge the per-GPU hit lists and print the report (reference swimm.c:150-190) */
static void batch_report(batch *b, swg_ctx **ctx, int ngpu, const swg_seqset *db, const swg_options *opt, unsigned long top)
{
    const uint64_t nq = b->q.count, t1 = top ? top : 1;
    uint64_t *part_keys = b->part_keys ? b->part_keys : (uint64_t *)calloc((size_t)ngpu * nq * t1, sizeof(uint64_t));
    uint64_t *keys = (uint64_t *)calloc(nq * t1, sizeof(uint64_t));
    double work = b->work;
    for (int g = 0; g < ngpu && !b->part_keys; g++) {
        int done = 0;
        double secs = 0;
        int st = swg_gpu_poll(ctx[g], b->tickets[g], 1, part_keys + (size_t)g * nq * top, &secs, &done);
        if (st != SWG_OK || !done)
            die_gpu("result download", ctx[g], st);
        if (secs > work)
            work = secs;                   /* the slowest GPU, like the reference's single workTime */
    }
    const double wall = swg_walltime() - b->t_submit;
    /* merge the per-GPU lists query by query */
    uint64_t *tmp = (uint64_t *)malloc((size_t)ngpu * t1 * sizeof(uint64_t));
    for (uint64_t i = 0; i < nq; i++) {
        for (int g = 0; g < ngpu; g++)
            memcpy(tmp + (size_t)g * top, part_keys + ((size_t)g * nq + i) * top, top * sizeof(uint64_t));
        swg_merge_top_keys(tmp, ngpu, top, keys + i * top);
    }
    free(tmp);

    printf("Query filename:\t\t\t%s\n", b->name);
    /* queries are reported in ascending-length order like the reference (sequences.c:344); --keep-input-order reports
     * them in the order of the file instead */
    uint64_t *show = (uint64_t *)malloc((nq ? nq : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < nq; i++)
        show[opt->keep_input_order && b->q.input_pos ? b->q.input_pos[i] : i] = i;
    uint64_t Q = 0;
    for (uint64_t n = 0; n < nq; n++) {
        const uint64_t i = show[n];
        Q += b->q.lengths[i] + (b->q.lengths[i] & 1u);      /* the reference counts odd lengths padded to even (sequences.c:370) */
        printf("\nQuery no.\t\t\t%d\n", (int)(n + 1));
        printf("Query description: \t\t%s\n", b->q.titles[i][0] ? b->q.titles[i] + 1 : "");
        printf("Query length:\t\t\t%d residues\n", b->q.lengths[i]);
        printf(b->part_coords ? "\nScore\tQuery range\tSequence range\tSequence description\n" : "\nScore\tSequence description\n");
        for (unsigned long j = 0; j < top; j++) {
            const uint64_t k = keys[i * top + j];
            const char *title = db->titles[SWG_KEY_INDEX(k)];
            if (b->part_coords) {
                /* --coordinates: query and database range of an optimal alignment, 1-based inclusive; the hit's
                 * coordinates sit beside its key in the list of the GPU that found it */
                const int32_t *c = NULL;
                for (int g = 0; g < ngpu && !c; g++)
                    for (unsigned long jj = 0; jj < top && !c; jj++)
                        if (part_keys[((size_t)g * nq + i) * top + jj] == k)
                            c = b->part_coords + (((size_t)g * nq + i) * top + jj) * 4;
                if (c && c[0] >= 0)
                    printf("%d\t%d-%d\t%d-%d\t%s\n", SWG_KEY_SCORE(k), c[0] + 1, c[1] + 1, c[2] + 1, c[3] + 1, title[0] ? title + 1 : "");
                else
                    printf("%d\t-\t-\t%s\n", SWG_KEY_SCORE(k), title[0] ? title + 1 : "");
                continue;
            }
            printf("%d\t%s\n", SWG_KEY_SCORE(k), title[0] ? title + 1 : "");
        }
    }
    free(show);
    /* Search time: CUDA-event time of the whole batch on the slowest GPU -- profile builds, all search kernels, top-r
     * selection and the hit-list download; the one-off loading of kernel code by a process's first batch and the
     * host-side merge of N x top keys are not in it (the reference's workTime covers its kernel call only, its sort
     * is outside too: swimm.c:151-160) */
    printf("\nSearch date:\t\t\t%s", ctime(&b->when));
    printf("Search time:\t\t\t%lf seconds\n", work);
    printf("Search speed:\t\t\t%.2lf GCUPS\n", ((double)Q * (double)db->residues) / (work * 1000000000.0));
    printf("Execution mode:\t\t\tB200 GPU only (%d GPU%s, submit-to-report %lf seconds)\n", ngpu, ngpu == 1 ? "" : "s", wall);
    printf("Profile technique:\t\tQuery Profile (shared memory)\n");
    printf("Instruction set:\t\tsm_100a DPX s16x2 (two sequences, or two queries of a batch, per 32-bit lane), 32-bit recomputation on overflow\n");
    fflush(stdout);
    if (!b->part_keys)
        free(part_keys);
    free(keys);
}

int main(int argc, char **argv)
{
    swg_options opt;
    swg_parse_arguments(argc, argv, &opt);

    if (strcmp(opt.op, "preprocess") == 0)
        return swg_preprocess_db(opt.input_filename, opt.output_filename, opt.cpu_threads, 1);

    printf("\nSWIMM v%s \n\n", SWG_VERSION);
    printf("Database file:\t\t\t%s\n", opt.sequences_filename);

    const int from_stdin = strcmp(opt.queries_filename, "-") == 0;
    /* like the reference (swimm.c:38), an unreadable query file is reported before anything is loaded */
    if (!from_stdin) {
        char *probe = strdup(opt.queries_filename);
        for (char *qfile = strtok(probe, ","); qfile; qfile = strtok(NULL, ",")) {
            FILE *f = fopen(qfile, "r");
            if (!f) {
                printf("SWIMM: An error occurred while opening input sequence file.\n");
                return 2;
            }
            fclose(f);
        }
        free(probe);
    }
    swg_seqset db;
    int rc = swg_load_db(opt.sequences_filename, &db);
    if (rc != 0) {
        printf("SWIMM: An error occurred while opening %s file.\n", rc == -1 ? "info/sequence" : "a damaged sequence");
        return 2;
    }
    unsigned long top = db.count < opt.top ? db.count : opt.top;      /* reference swimm.c:51 */
    unsigned max_len = 0;
    for (uint64_t i = 0; i < db.count; i++)
        if (db.lengths[i] > max_len)
            max_len = db.lengths[i];

    printf("Database size:\t\t\t%ld sequences (%ld residues) \n", (long)db.count, (long)db.residues);
    printf("Longest database sequence: \t%d residues\n", max_len);
    printf("Substitution matrix:\t\t%s\n", swg_submat_shown(opt.submat));
    printf("Gap open penalty:\t\t%d\n", opt.open_gap);
    printf("Gap extend penalty:\t\t%d\n", opt.extend_gap);

    /* GPUs: one host thread each creates the context and uploads its shard */
    int visible = 0, st = swg_gpu_device_count(&visible);
    if (st != SWG_OK)
        die_gpu("GPU discovery", NULL, st);
    int ngpu = opt.num_gpus == 0 ? visible : opt.num_gpus;
    if (ngpu > visible) {
        printf("SWIMM: %d GPUs requested, %d visible.\n", ngpu, visible);
        return 4;
    }
    swg_ctx **ctx = (swg_ctx **)calloc((size_t)ngpu, sizeof(swg_ctx *));
    load_job *jobs = (load_job *)calloc((size_t)ngpu, sizeof(load_job));
    pthread_t *threads = (pthread_t *)calloc((size_t)ngpu, sizeof(pthread_t));
    const double t_load = swg_walltime();
    for (int g = 0; g < ngpu; g++) {
        jobs[g].gpu = g;
        jobs[g].ngpu = ngpu;
        jobs[g].db = &db;
        if (ngpu == 1 || pthread_create(&threads[g], NULL, load_thread, &jobs[g]) != 0) {
            load_thread(&jobs[g]);
            threads[g] = 0;
        }
    }
    for (int g = 0; g < ngpu; g++) {
        if (threads[g])
            pthread_join(threads[g], NULL);
        ctx[g] = jobs[g].ctx;
        if (jobs[g].st != SWG_OK) {
            printf("SWIMM: database upload to GPU %d failed (%d): %s\n", g, jobs[g].st, jobs[g].err);
            return 4;
        }
    }
    if (opt.verbose)
        fprintf(stderr, "[swimm] %d shard%s resident after %.3f s\n", ngpu, ngpu == 1 ? "" : "s", swg_walltime() - t_load);
    free(jobs);
    free(threads);

    if (swg_load_db_headers(opt.sequences_filename, &db) != 0) {
        printf("SWIMM: An error occurred while opening sequence description file.\n");
        return 3;
    }

    /* -q takes one FASTA file like the reference, several separated by commas, or "-" (file names on stdin): the
     * database stays resident on the GPUs, every file is a batch, and two batches are in flight -- file k+1 is parsed,
     * uploaded and queued while batch k computes ("keep DB resident, stream query files") */
    batch pending, next;
    int have_pending = 0;
    char *qlist = from_stdin ? NULL : strdup(opt.queries_filename);
    char *cursor = qlist, line[4096];
    for (;;) {
        const char *qfile = NULL;
        if (from_stdin) {
            if (!fgets(line, sizeof(line), stdin))
                break;
            line[strcspn(line, "\r\n")] = '\0';
            if (!line[0])
                continue;
            qfile = line;
        } else {
            qfile = strtok(cursor, ",");
            cursor = NULL;
            if (!qfile)
                break;
        }
        rc = batch_submit(&next, qfile, ctx, ngpu, &opt, top);
        if (rc != 0) {
            if (!from_stdin)
                return rc;
            continue;                     /* a server keeps going when one file is unreadable */
        }
        if (have_pending) {
            batch_report(&pending, ctx, ngpu, &db, &opt, top);
            batch_free(&pending);
        }
        pending = next;
        have_pending = 1;
    }
    if (have_pending) {
        batch_report(&pending, ctx, ngpu, &db, &opt, top);
        batch_free(&pending);
    }
    free(qlist);
    for (int g = 0; g < ngpu; g++)
        swg_gpu_destroy(ctx[g]);
    free(ctx);
    swg_seqset_free(&db);
    return 0;
}
