/*
 * arguments.h -- command line of the B200 swimm build.  Same flags as the reference (arguments.c:15-38);
 * `-m 3` selects the GPU execution mode this build adds, `-x` then counts GPUs instead of Xeon Phis.
 */
#ifndef SWG_ARGUMENTS_H
#define SWG_ARGUMENTS_H

#define SWG_MODE_CPU_ONLY      0
#define SWG_MODE_MIC_ONLY      1
#define SWG_MODE_HETEROGENEOUS 2
#define SWG_MODE_GPU           3

typedef struct {
    const char *op;                 /* "preprocess" | "search" */
    const char *input_filename;     /* -i */
    const char *output_filename;    /* -o */
    const char *queries_filename;   /* -q */
    const char *sequences_filename; /* -d */
    int submat;                     /* index into the matrix table, -s */
    int open_gap, extend_gap;       /* -g, -e */
    int execution_mode;             /* -m */
    int cpu_threads;                /* -c (host threads; only preprocessing uses them) */
    int num_gpus;                   /* -x (0 = all visible) */
    int mic_threads;                /* -t accepted, ignored */
    char profile;                   /* -p accepted, ignored */
    int query_length_threshold;     /* -u accepted, ignored */
    int vector_length;              /* -v accepted, ignored */
    unsigned long top;              /* -r */
    unsigned long max_chunk_size;   /* -k accepted, ignored */
    int block_size;                 /* -b accepted, ignored */
    int keep_input_order;           /* --keep-input-order: report the queries of a file in file order, not by ascending length */
    int verbose;                    /* --verbose: progress notes on stderr */
    int coordinates;                /* --coordinates: print the query / sequence range of every hit's alignment (extension) */
} swg_options;

void swg_parse_arguments(int argc, char **argv, swg_options *opt);

#endif
