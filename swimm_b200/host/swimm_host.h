/*
 * swimm_host.h -- host-side (plain C) data layer of the B200 swimm build.
 *
 * Same role as the reference's sequences.h / submat.h / utils.h / arguments.h: FASTA parsing,
 * `-S preprocess` (byte-compatible .info/.seq/.desc), query loading, database loading, hit
 * ordering and the report.  It knows nothing about CUDA; the search itself goes through the C ABI
 * in include/swimm_gpu.h.
 */
#ifndef SWIMM_HOST_H
#define SWIMM_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWG_VERSION        "1.1.3-b200"
#define SWG_ALPHABET       23      /* A B C D E F G H I K L M N P Q R S T V W X Y Z */
#define SWG_SUBMAT_ROWS    24      /* reference submat.h:4 */
#define SWG_SUBMAT_COLS    32      /* reference submat.h:5 */
#define SWG_SUBMAT_ELEMS   768
#define SWG_NUM_MATRICES   8
#define SWG_DUMMY_CODE     23      /* J, O, U and anything that is not a letter */
#define SWG_PAD_CODE       24      /* reference sequences.h:17 (database padding) */
#define SWG_MAX_SEQ_LEN    65535   /* lengths are stored as unsigned short (reference sequences.c:7,202) */

/* ---- substitution matrices (submat.c) ---- */
int swg_submat_index(const char *name);              /* -1 if unknown */
const signed char *swg_submat_table(int k);          /* 24x32, 64-byte aligned */
const char *swg_submat_key(int k);
const char *swg_submat_shown(int k);

/* ---- sequences (sequences.c) ---- */
typedef struct {
    uint64_t count;          /* number of sequences */
    uint64_t residues;       /* total residues */
    uint16_t *lengths;       /* [count]  ascending, stable w.r.t. input order */
    uint64_t *offsets;       /* [count+1] prefix sums of lengths */
    signed char *codes;      /* [residues] codes 0..23, concatenated in the same order */
    char **titles;           /* [count] header lines WITH the leading '>' and without '\n' (may be NULL) */
    int max_title;           /* longest header line incl. '\n', +1 (reference sequences.c:41,96) */
    uint64_t *input_pos;     /* [count] position in the input file of each (sorted) sequence; NULL for a preprocessed database */
} swg_seqset;

int swg_encode_residue(int c);                                           /* reference sequences.c:165-175 */
int swg_read_fasta(const char *path, swg_seqset *out);                   /* parsed, length-sorted, encoded */
void swg_seqset_free(swg_seqset *s);

/* `-S preprocess`: FASTA -> <prefix>.info/.seq/.desc (reference sequences.c:4-220). 0 on success. */
int swg_preprocess_db(const char *fasta_path, const char *out_prefix, int threads, int verbose);

/* read <prefix>.info + .seq (flat residues, no interleave: the GPU builds its own layout) */
int swg_load_db(const char *prefix, swg_seqset *out);
/* read <prefix>.desc into s->titles (reference sequences.c:736-767) */
int swg_load_db_headers(const char *prefix, swg_seqset *s);

/* ---- utilities (utils.c) ---- */
double swg_walltime(void);
/* merge `parts` descending key lists (each `r` long, key = score<<32 | index) into the `r` largest */
void swg_merge_top_keys(const uint64_t *keys, int parts, uint64_t r, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
