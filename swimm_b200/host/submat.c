/*
 * Substitution matrices for the swimm host (drop-in for the reference's submat.c).
 *
 * The reference keeps eight 24x32 signed-byte tables (submat.c:4-227; rows and columns in the
 * order A B C D E F G H I K L M N P Q R S T V W X Y Z, row 23 and columns 23..31 all zero so
 * that the dummy residue and database padding score 0).  Here only the lower triangles are
 * stored (submat_tri.inc, generated data) and the padded 24x32 tables -- the layout the
 * search kernels and the C ABI expect -- are expanded on first use.
 */
#include "swimm_host.h"

#include <string.h>
#include <strings.h>

#define SWG_TRI_BEGIN(name) static const signed char tri_##name[] = {
#define SWG_TRI_END(name)   };
#include "submat_tri.inc"
#undef SWG_TRI_BEGIN
#undef SWG_TRI_END

typedef struct {
    const char *key;          /* value accepted by -s */
    const char *shown;        /* value printed in the report (reference arguments.c:76-83) */
    const signed char *tri;
} matrix_entry;

static const matrix_entry k_matrices[SWG_NUM_MATRICES] = {
    {"blosum45", "BLOSUM45", tri_blosum45}, {"blosum50", "BLOSUM50", tri_blosum50},
    {"blosum62", "BLOSUM62", tri_blosum62}, {"blosum80", "BLOSUM80", tri_blosum80},
    {"blosum90", "BLOSUM90", tri_blosum90}, {"pam30", "PAM30", tri_pam30},
    {"pam70", "PAM70", tri_pam70},          {"pam250", "PAM250", tri_pam250},
};

static signed char g_tables[SWG_NUM_MATRICES][SWG_SUBMAT_ELEMS] __attribute__((aligned(64)));
static int g_ready[SWG_NUM_MATRICES];

static void expand(int k)
{
    signed char *t = g_tables[k];
    const signed char *tri = k_matrices[k].tri;
    memset(t, 0, SWG_SUBMAT_ELEMS);
    for (int r = 0, p = 0; r < SWG_ALPHABET; r++)
        for (int c = 0; c <= r; c++, p++) {
            t[r * SWG_SUBMAT_COLS + c] = tri[p];
            t[c * SWG_SUBMAT_COLS + r] = tri[p];
        }
    g_ready[k] = 1;
}

int swg_submat_index(const char *name)
{
    for (int k = 0; k < SWG_NUM_MATRICES; k++)
        if (strcasecmp(name, k_matrices[k].key) == 0)
            return k;
    return -1;
}

const signed char *swg_submat_table(int k)
{
    if (k < 0 || k >= SWG_NUM_MATRICES)
        return NULL;
    if (!g_ready[k])
        expand(k);
    return g_tables[k];
}

const char *swg_submat_key(int k)   { return (k < 0 || k >= SWG_NUM_MATRICES) ? NULL : k_matrices[k].key; }
const char *swg_submat_shown(int k) { return (k < 0 || k >= SWG_NUM_MATRICES) ? NULL : k_matrices[k].shown; }
