/*
 * swimm.c -- driver of the B200 swimm build: `swimm -S preprocess ...` and `swimm -S search ... -m 3`.
 *
 * Mirrors the reference driver (swimm.c:9-207): parse arguments, preprocess OR (load queries, load the
 * database, search, load headers, print the top hits of every query, print time / GCUPS / mode), with
 * the same stdout layout.  The search itself is the C ABI of libswimm_cuda.so (include/swimm_gpu.h);
 * with -x N the database is sharded over N GPUs and the per-GPU hit lists are merged here.
 */
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

int main(int argc, char **argv)
{
    swg_options opt;
    swg_parse_arguments(argc, argv, &opt);

    if (strcmp(opt.op, "preprocess") == 0)
        return swg_preprocess_db(opt.input_filename, opt.output_filename, opt.cpu_threads, 1);

    printf("\nSWIMM v%s \n\n", SWG_VERSION);
    printf("Database file:\t\t\t%s\n", opt.sequences_filename);

    /* like the reference (swimm.c:38), an unreadable query file is reported before anything is loaded */
    {
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

    /* GPUs */
    int visible = 0, st = swg_gpu_device_count(&visible);
    if (st != SWG_OK)
        die_gpu("GPU discovery", NULL, st);
    int ngpu = opt.num_gpus == 0 ? visible : opt.num_gpus;
    if (ngpu > visible) {
        printf("SWIMM: %d GPUs requested, %d visible.\n", ngpu, visible);
        return 4;
    }
    swg_ctx **ctx = (swg_ctx **)calloc((size_t)ngpu, sizeof(swg_ctx *));
    for (int g = 0; g < ngpu; g++) {
        if ((st = swg_gpu_create(g, &ctx[g])) != SWG_OK)
            die_gpu("GPU context creation", NULL, st);
        if ((st = swg_gpu_load_db(ctx[g], db.lengths, db.codes, db.count, db.residues, g, ngpu)) != SWG_OK)
            die_gpu("database upload", ctx[g], st);
    }

    if (swg_load_db_headers(opt.sequences_filename, &db) != 0) {
        printf("SWIMM: An error occurred while opening sequence description file.\n");
        return 3;
    }

    /* -q takes one FASTA file like the reference, or several separated by commas: the database stays resident on
     * the GPUs and every file is searched and reported in turn ("keep DB resident, stream query files") */
    char *qlist = strdup(opt.queries_filename);
    for (char *qfile = strtok(qlist, ","); qfile; qfile = strtok(NULL, ",")) {
        /* queries: parsed, sorted by ascending length, encoded (reference sequences.c:223-423) */
        swg_seqset q;
        if (swg_read_fasta(qfile, &q) != 0) {
            printf("SWIMM: An error occurred while opening input sequence file.\n");
            return 2;
        }
        time_t current_time = time(NULL);
        printf("Query filename:\t\t\t%s\n", qfile);

        /* search: every GPU gets all queries and its shard of the database */
        uint32_t *q_disp = (uint32_t *)malloc((q.count + 1) * sizeof(uint32_t));
        for (uint64_t i = 0; i <= q.count; i++)
            q_disp[i] = (uint32_t)q.offsets[i];
        uint64_t *part_keys = (uint64_t *)calloc((size_t)ngpu * q.count * (top ? top : 1), sizeof(uint64_t));
        uint64_t *keys = (uint64_t *)calloc(q.count * (top ? top : 1), sizeof(uint64_t));
        const double t0 = swg_walltime();
        for (int g = 0; g < ngpu; g++) {
            st = swg_gpu_set_queries(ctx[g], q.codes, q.lengths, q_disp, q.count, swg_submat_table(opt.submat), opt.open_gap,
                                     opt.extend_gap);
            if (st == SWG_OK)
                st = swg_gpu_run(ctx[g], top, 0);
            if (st != SWG_OK)
                die_gpu("search", ctx[g], st);
        }
        double work = 0;
        for (int g = 0; g < ngpu; g++) {
            if ((st = swg_gpu_fetch(ctx[g], NULL, part_keys + (size_t)g * q.count * top)) != SWG_OK)
                die_gpu("result download", ctx[g], st);
            swg_stats s;
            swg_gpu_get_stats(ctx[g], &s);
            if (s.search_seconds > work)
                work = s.search_seconds;           /* the slowest GPU, like the reference's single workTime */
        }
        const double wall = swg_walltime() - t0;
        /* merge the per-GPU lists query by query */
        uint64_t *tmp = (uint64_t *)malloc((size_t)ngpu * (top ? top : 1) * sizeof(uint64_t));
        for (uint64_t i = 0; i < q.count; i++) {
            for (int g = 0; g < ngpu; g++)
                memcpy(tmp + (size_t)g * top, part_keys + ((size_t)g * q.count + i) * top, top * sizeof(uint64_t));
            swg_merge_top_keys(tmp, ngpu, top, keys + i * top);
        }
        free(tmp);
        uint64_t Q = 0;
        for (uint64_t i = 0; i < q.count; i++) {
            Q += q.lengths[i];
            printf("\nQuery no.\t\t\t%d\n", (int)(i + 1));
            printf("Query description: \t\t%s\n", q.titles[i][0] ? q.titles[i] + 1 : "");
            printf("Query length:\t\t\t%d residues\n", q.lengths[i]);
            printf("\nScore\tSequence description\n");
            for (unsigned long j = 0; j < top; j++) {
                const uint64_t k = keys[i * top + j];
                const char *title = db.titles[SWG_KEY_INDEX(k)];
                printf("%d\t%s\n", SWG_KEY_SCORE(k), title[0] ? title + 1 : "");
            }
        }
        printf("\nSearch date:\t\t\t%s", ctime(&current_time));
        printf("Search time:\t\t\t%lf seconds\n", work);
        printf("Search speed:\t\t\t%.2lf GCUPS\n", ((double)Q * (double)db.residues) / (work * 1000000000.0));
        printf("Execution mode:\t\t\tB200 GPU only (%d GPU%s, end-to-end %lf seconds)\n", ngpu, ngpu == 1 ? "" : "s", wall);
        printf("Profile technique:\t\tQuery Profile (shared memory)\n");
        printf("Instruction set:\t\tsm_100a DPX s16x2 (two sequences, or two queries of a batch, per 32-bit lane), 32-bit recomputation on overflow\n");


        free(q_disp);
        free(part_keys);
        free(keys);
        swg_seqset_free(&q);
    }
    free(qlist);
    for (int g = 0; g < ngpu; g++)
        swg_gpu_destroy(ctx[g]);
    free(ctx);
    swg_seqset_free(&db);
    return 0;
}
