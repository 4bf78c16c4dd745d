/*
 * arguments.c -- GNU argp command line, flag-for-flag the reference's (arguments.c:15-38, :44-172),
 * with execution mode 3 (B200 GPUs) added.  Modes 0-2 (Xeon / Xeon Phi / hybrid) are recognised and
 * refused: this build has no CPU or Xeon Phi kernels.
 */
#include "arguments.h"

#include <argp.h>
#include <stdlib.h>
#include <string.h>

#include "swimm_host.h"

const char *argp_program_bug_address = "<swimm-b200 maintainers>";

static char doc[] =
    "\nSWIMM (B200 build) accelerates Smith-Waterman protein database search on NVIDIA B200 GPUs; "
    "command line and preprocessed database format are those of SWIMM 1.1.3";

static struct argp_option options[] = {
    {0, 0, 0, 0, "SWIMM execution", 1},
    {0, 'S', "<string>", 0, "'preprocess' for database preprocessing, 'search' for database search. [REQUIRED]", 1},
    {0, 0, 0, 0, "preprocess", 2},
    {"input", 'i', "<string>", 0, "Input sequence filename (must be in FASTA format). [REQUIRED]", 2},
    {"output", 'o', "<string>", 0, "Output filename. [REQUIRED]", 2},
    {0, 0, 0, 0, "search", 3},
    {"query", 'q', "<string>", 0,
     "Input query sequence filename (must be in FASTA format); several files separated by commas, or - to read file "
     "names from standard input: the database stays resident and the files are searched as a stream of batches. [REQUIRED]", 3},
    {"db", 'd', "<string>", 0, "Preprocessed database output filename. [REQUIRED]", 3},
    {"sm", 's', "<string>", 0,
     "Substitution matrix. Supported values: blosum45, blosum50, blosum62, blosum80, blosum90, pam30, pam70, pam250 "
     "(default: blosum62).", 3},
    {"gap_open", 'g', "<integer>", 0, "Gap open penalty (default: 10).", 3},
    {"gap_extend", 'e', "<integer>", 0, "Gap extend penalty (default: 2).", 3},
    {"execution_mode", 'm', "<integer>", 0,
     "Execution mode: 3 for B200 GPUs (default). 0 (Xeon), 1 (Xeon Phi) and 2 (Xeon+Xeon Phi) are not part of this build.", 3},
    {"cpu_threads", 'c', "<integer>", 0, "Number of host threads (default: 4).", 3},
    {"num_mics", 'x', "<integer>", 0, "Number of GPUs, 0 for all visible (default: 1).", 3},
    {"mic_threads", 't', "<integer>", 0, "Accepted for compatibility, ignored.", 3},
    {"mic_profile", 'p', "<char>", 0, "Accepted for compatibility, ignored (the GPU kernels use a query profile).", 3},
    {"query_length_threshold", 'u', "<integer>", 0, "Accepted for compatibility, ignored.", 3},
    {"vector_length", 'v', "<integer>", 0, "Accepted for compatibility (16 or 32), ignored.", 3},
    {"top", 'r', "<integer>", 0, "Number of scores to show (default: 10).", 3},
    {"max_chunk_size", 'k', "<integer>", 0, "Accepted for compatibility, ignored.", 3},
    {"block_size", 'b', "<integer>", 0, "Accepted for compatibility, ignored.", 3},
    {"keep-input-order", 1001, 0, 0,
     "Report the queries of a file in the order of the file.  Default: ascending length, the order the reference "
     "reports (it sorts the queries, sequences.c:344; its own lengths array is copied before that sort, :276, so it "
     "requires multi-query files that are already sorted -- this build accepts any order).", 3},
    {"verbose", 1002, 0, 0, "Progress notes on stderr.", 3},
    {"coordinates", 1003, 0, 0,
     "Also print, for every hit, the query range and the sequence range (1-based) of an optimal alignment.  An "
     "extension: the reference reports scores only; without this flag the output keeps the reference's layout.", 3},
    {0}};

typedef struct {
    swg_options *opt;
    int argc;
} parse_state;

static int parse_opt(int key, char *arg, struct argp_state *state)
{
    parse_state *ps = (parse_state *)state->input;
    swg_options *o = ps->opt;
    switch (key) {
    case 'S':
        if (strcmp(arg, "preprocess") != 0 && strcmp(arg, "search") != 0)
            argp_failure(state, 1, 0, "%s is not a valid option for execution.", arg);
        else
            o->op = arg;
        break;
    case 'i': o->input_filename = arg; break;
    case 'o': o->output_filename = arg; break;
    case 'q': o->queries_filename = arg; break;
    case 'd': o->sequences_filename = arg; break;
    case 'm':
        o->execution_mode = atoi(arg);
        if (o->execution_mode < SWG_MODE_CPU_ONLY || o->execution_mode > SWG_MODE_GPU)
            argp_failure(state, 1, 0, "%d is not a valid option for execution mode.", o->execution_mode);
        break;
    case 's':
        o->submat = swg_submat_index(arg);
        if (o->submat < 0)
            argp_failure(state, 1, 0, "%s is not a valid option for substitution matrix.", arg);
        break;
    case 'g':
        o->open_gap = atoi(arg);
        if (o->open_gap < 0 || o->open_gap > 127)
            argp_failure(state, 1, 0, "%s is not a valid option for gap open penalty.", arg);
        break;
    case 'e':
        o->extend_gap = atoi(arg);
        if (o->extend_gap < 0 || o->extend_gap > 127)
            argp_failure(state, 1, 0, "%s is not a valid option for gap extend penalty.", arg);
        break;
    case 'c':
        o->cpu_threads = atoi(arg);
        if (o->cpu_threads < 0)
            argp_failure(state, 1, 0, "The number of host threads must be greater than 0.");
        break;
    case 'x':
        o->num_gpus = atoi(arg);
        if (o->num_gpus < 0)
            argp_failure(state, 1, 0, "The number of GPUs must not be negative.");
        break;
    case 't': o->mic_threads = atoi(arg); break;
    case 'p':
        if (strcmp(arg, "Q") != 0 && strcmp(arg, "S") != 0 && strcmp(arg, "A") != 0)
            argp_failure(state, 1, 0, "%s is not a valid option for profile technique.", arg);
        else
            o->profile = arg[0];
        break;
    case 'u': o->query_length_threshold = atoi(arg); break;
    case 'v':
        o->vector_length = atoi(arg);
        if (o->vector_length != 16 && o->vector_length != 32)
            argp_failure(state, 1, 0, "%d is not a valid option for vector length.", o->vector_length);
        break;
    case 'r': {
        long t = atol(arg);
        if (t < 0)
            argp_failure(state, 1, 0, "The number of scores to show must be greater than 0.");
        o->top = (unsigned long)t;
        break;
    }
    case 'k': o->max_chunk_size = strtoul(arg, NULL, 10); break;
    case 'b': o->block_size = atoi(arg); break;
    case 1001: o->keep_input_order = 1; break;
    case 1002: o->verbose = 1; break;
    case 1003: o->coordinates = 1; break;
    case ARGP_KEY_END:
        if (ps->argc == 1)
            argp_failure(state, 1, 0, "Missing options");
        if (o->op == NULL)
            argp_failure(state, 1, 0, "SWIMM execution option is required");
        else if (strcmp(o->op, "preprocess") == 0) {
            if (o->input_filename == NULL)
                argp_failure(state, 1, 0, "Input sequence filename is required");
            if (o->output_filename == NULL)
                argp_failure(state, 1, 0, "Output filename is required");
        } else {
            if (o->sequences_filename == NULL)
                argp_failure(state, 1, 0, "Database filename is required");
            if (o->queries_filename == NULL)
                argp_failure(state, 1, 0, "Query sequences filename is required");
            if (o->execution_mode != SWG_MODE_GPU)
                argp_failure(state, 1, 0,
                             "Execution mode %d needs the Xeon / Xeon Phi kernels of the original SWIMM; this build "
                             "only has the B200 GPU mode (-m 3) and no CPU fallback.", o->execution_mode);
        }
        break;
    default:
        break;
    }
    return 0;
}

void swg_parse_arguments(int argc, char **argv, swg_options *o)
{
    memset(o, 0, sizeof(*o));
    o->submat = swg_submat_index("blosum62");
    o->open_gap = 10;
    o->extend_gap = 2;
    o->execution_mode = SWG_MODE_GPU;
    o->cpu_threads = 4;
    o->num_gpus = 1;
    o->mic_threads = 240;
    o->vector_length = 32;
    o->top = 10;
    o->max_chunk_size = 100663296ul;
    parse_state ps = {o, argc};
    struct argp argp = {options, parse_opt, 0, doc};
    argp_parse(&argp, argc, argv, 0, 0, &ps);
}
