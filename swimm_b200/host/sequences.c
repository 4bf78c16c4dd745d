/*
 * sequences.c -- FASTA parsing, `-S preprocess`, query and database loading for the B200 swimm host.
 *
 * Same role and same on-disk format as the reference's sequences.c (preprocess_db :4-220,
 * load_query_sequences :223-423, load_database_headers :736-767), rebuilt around one pass over a
 * file image and a counting sort on the 16-bit length (stable, ascending -- the order the reference's
 * merge sort produces, sequences.c:770-865).  The lane interleave of assemble_single_chunk_db
 * (sequences.c:618-734) is NOT done here: the GPU library builds its own tiled layout from the flat
 * arrays (include/swimm_gpu.h, swg_gpu_load_db).
 */
#include "swimm_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* 'A'..'Z' -> 0..22, J/O/U (and anything that is not a letter) -> 23.  reference sequences.c:165-175 */
int swg_encode_residue(int c)
{
    if (c >= 'a' && c <= 'z')
        c -= 'a' - 'A';
    if (c < 'A' || c > 'Z' || c == 'J' || c == 'O' || c == 'U')
        return SWG_DUMMY_CODE;
    return c - 'A' - (c > 'J') - (c > 'O') - (c > 'U');
}

static char *read_whole_file(const char *path, size_t *size)
{
    FILE *f = fopen(path, "rb");
    if (!f)
        return NULL;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return NULL; }
    long n = ftell(f);
    if (n < 0) { fclose(f); return NULL; }
    rewind(f);
    char *buf = (char *)malloc((size_t)n + 2);
    if (!buf) { fclose(f); return NULL; }
    size_t got = fread(buf, 1, (size_t)n, f);
    fclose(f);
    buf[got] = '\n';            /* sentinel: every line is terminated */
    buf[got + 1] = '\0';
    *size = got;
    return buf;
}

void swg_seqset_free(swg_seqset *s)
{
    if (!s)
        return;
    free(s->lengths);
    free(s->offsets);
    free(s->codes);
    free(s->input_pos);
    if (s->titles) {
        for (uint64_t i = 0; i < s->count; i++)
            free(s->titles[i]);
        free(s->titles);
    }
    memset(s, 0, sizeof(*s));
}

/* One record of the file image: where its header and its residue lines are. */
typedef struct {
    size_t title_at, title_len;     /* header line without the line terminator, with the leading '>' */
    size_t body_at, body_end;
    uint32_t length;                /* residues */
} fasta_rec;

int swg_read_fasta(const char *path, swg_seqset *out)
{
    memset(out, 0, sizeof(*out));
    size_t size = 0;
    char *img = read_whole_file(path, &size);
    if (!img)
        return -1;

    /* pass 1 over the image: record boundaries and lengths */
    size_t cap = 1024, n = 0;
    fasta_rec *rec = (fasta_rec *)malloc(cap * sizeof(fasta_rec));
    int rc = 0;
    size_t p = 0;
    while (rec && p < size) {
        size_t eol = p;
        while (img[eol] != '\n')
            eol++;
        if (img[p] == '>') {
            if (n == cap) {
                cap *= 2;
                fasta_rec *r2 = (fasta_rec *)realloc(rec, cap * sizeof(fasta_rec));
                if (!r2) { rc = -2; break; }
                rec = r2;
            }
            size_t tl = eol - p;
            if (tl && img[p + tl - 1] == '\r')
                tl--;
            rec[n].title_at = p;
            rec[n].title_len = tl;
            rec[n].body_at = eol + 1 < size ? eol + 1 : size;
            rec[n].body_end = rec[n].body_at;
            rec[n].length = 0;
            n++;
        } else if (n) {
            uint32_t k = 0;
            for (size_t i = p; i < eol; i++)
                k += (img[i] > ' ');
            rec[n - 1].length += k;
            rec[n - 1].body_end = eol;
        }
        p = eol + 1;
    }
    if (!rec)
        rc = -2;
    for (size_t i = 0; rc == 0 && i < n; i++)
        if (rec[i].length > SWG_MAX_SEQ_LEN)
            rc = -3;               /* lengths are stored as unsigned short (reference sequences.c:7,202) */
    if (rc) { free(rec); free(img); return rc; }

    /* stable counting sort by length: order[k] = input position of the k-th sorted sequence */
    uint64_t *bucket = (uint64_t *)calloc((size_t)SWG_MAX_SEQ_LEN + 2, sizeof(uint64_t));
    uint64_t *order = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    out->lengths = (uint16_t *)malloc((n ? n : 1) * sizeof(uint16_t));
    out->offsets = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    out->titles = (char **)calloc(n ? n : 1, sizeof(char *));
    if (!bucket || !order || !out->lengths || !out->offsets || !out->titles) {
        free(bucket); free(order); free(rec); free(img);
        swg_seqset_free(out);
        return -2;
    }
    for (size_t i = 0; i < n; i++)
        bucket[rec[i].length + 1]++;
    for (size_t l = 1; l <= (size_t)SWG_MAX_SEQ_LEN + 1; l++)
        bucket[l] += bucket[l - 1];
    for (size_t i = 0; i < n; i++)
        order[bucket[rec[i].length]++] = i;
    free(bucket);

    out->count = n;
    uint64_t total = 0;
    int max_title = 0;
    for (size_t k = 0; k < n; k++) {
        const fasta_rec *r = &rec[order[k]];
        out->lengths[k] = (uint16_t)r->length;
        out->offsets[k] = total;
        total += r->length;
        /* reference sequences.c:41: strlen(header line incl. '\n') + 1 */
        if ((int)r->title_len + 2 > max_title)
            max_title = (int)r->title_len + 2;
    }
    out->offsets[n] = total;
    out->residues = total;
    out->max_title = max_title;
    out->codes = (signed char *)malloc(total ? total : 1);
    if (!out->codes) { free(order); free(rec); free(img); swg_seqset_free(out); return -2; }
    for (size_t k = 0; k < n; k++) {
        const fasta_rec *r = &rec[order[k]];
        signed char *dst = out->codes + out->offsets[k];
        for (size_t i = r->body_at; i < r->body_end; i++)
            if (img[i] > ' ')
                *dst++ = (signed char)swg_encode_residue((unsigned char)img[i]);
        out->titles[k] = (char *)malloc(r->title_len + 1);
        if (!out->titles[k]) { rc = -2; break; }
        memcpy(out->titles[k], img + r->title_at, r->title_len);
        out->titles[k][r->title_len] = '\0';
    }
    out->input_pos = order;
    free(rec);
    free(img);
    if (rc) { swg_seqset_free(out); return rc; }
    return 0;
}

/* `-S preprocess`.  Files (reference sequences.c:128-205):
 *   <prefix>.desc  one header line per sequence (with '>'), sorted order
 *   <prefix>.info  "%ld %ld %d" = count, residues, max title length; no newline
 *   <prefix>.seq   uint16 lengths[count] then int8 codes[residues]                                    */
int swg_preprocess_db(const char *fasta_path, const char *out_prefix, int threads, int verbose)
{
    (void)threads;
    const double t0 = swg_walltime();
    swg_seqset s;
    int rc = swg_read_fasta(fasta_path, &s);
    if (rc) {
        if (rc == -3)
            printf("SWIMM: a sequence is longer than %d residues.\n", SWG_MAX_SEQ_LEN);
        else
            printf("SWIMM: An error occurred while opening input sequence file.\n");
        return 2;
    }
    size_t plen = strlen(out_prefix);
    char *name = (char *)malloc(plen + 8);
    FILE *f;
    sprintf(name, "%s.desc", out_prefix);
    if (!(f = fopen(name, "w"))) {
        printf("SWIMM: An error occurred while opening sequence header file.\n");
        free(name); swg_seqset_free(&s);
        return 2;
    }
    for (uint64_t i = 0; i < s.count; i++)
        fprintf(f, "%s\n", s.titles[i]);
    fclose(f);
    sprintf(name, "%s.info", out_prefix);
    if (!(f = fopen(name, "w"))) {
        printf("SWIMM: An error occurred while opening info file.\n");
        free(name); swg_seqset_free(&s);
        return 2;
    }
    fprintf(f, "%ld %ld %d", (long)s.count, (long)s.residues, s.max_title);
    fclose(f);
    sprintf(name, "%s.seq", out_prefix);
    if (!(f = fopen(name, "wb"))) {
        printf("SWIMM: An error occurred while opening sequence file.\n");
        free(name); swg_seqset_free(&s);
        return 2;
    }
    fwrite(s.lengths, sizeof(uint16_t), s.count, f);
    fwrite(s.codes, 1, s.residues, f);
    fclose(f);
    free(name);
    if (verbose) {
        printf("\nSWIMM v%s\n\n", SWG_VERSION);
        printf("Database file:\t\t\t %s\n", fasta_path);
        printf("Database size:\t\t\t%ld sequences (%ld residues) \n", (long)s.count, (long)s.residues);
        printf("Preprocessed database name:\t%s\n", out_prefix);
        printf("Preprocessing time:\t\t%lf seconds\n\n", swg_walltime() - t0);
    }
    swg_seqset_free(&s);
    return 0;
}

int swg_load_db(const char *prefix, swg_seqset *out)
{
    memset(out, 0, sizeof(*out));
    size_t plen = strlen(prefix);
    char *name = (char *)malloc(plen + 8);
    sprintf(name, "%s.info", prefix);
    FILE *f = fopen(name, "r");
    if (!f) { free(name); return -1; }
    long n = 0, d = 0;
    int mt = 0;
    int got = fscanf(f, "%ld %ld %d", &n, &d, &mt);
    fclose(f);
    if (got != 3 || n < 0 || d < 0) { free(name); return -4; }
    sprintf(name, "%s.seq", prefix);
    f = fopen(name, "rb");
    free(name);
    if (!f)
        return -1;
    out->count = (uint64_t)n;
    out->residues = (uint64_t)d;
    out->max_title = mt;
    out->lengths = (uint16_t *)malloc((n ? n : 1) * sizeof(uint16_t));
    out->offsets = (uint64_t *)malloc(((size_t)n + 1) * sizeof(uint64_t));
    out->codes = (signed char *)malloc(d ? d : 1);
    if (!out->lengths || !out->offsets || !out->codes) { fclose(f); swg_seqset_free(out); return -2; }
    size_t a = fread(out->lengths, sizeof(uint16_t), (size_t)n, f);
    size_t b = fread(out->codes, 1, (size_t)d, f);
    fclose(f);
    if (a != (size_t)n || b != (size_t)d) { swg_seqset_free(out); return -4; }
    uint64_t acc = 0;
    for (long i = 0; i < n; i++) {
        out->offsets[i] = acc;
        acc += out->lengths[i];
    }
    out->offsets[n] = acc;
    if (acc != (uint64_t)d) { swg_seqset_free(out); return -4; }
    return 0;
}

/* <prefix>.desc -> s->titles (header lines WITH '>' and without '\n').  reference sequences.c:736-767 */
int swg_load_db_headers(const char *prefix, swg_seqset *s)
{
    size_t plen = strlen(prefix), size = 0;
    char *name = (char *)malloc(plen + 8);
    sprintf(name, "%s.desc", prefix);
    char *img = read_whole_file(name, &size);
    free(name);
    if (!img)
        return -1;
    if (!s->titles)
        s->titles = (char **)calloc(s->count ? s->count : 1, sizeof(char *));
    size_t p = 0;
    for (uint64_t i = 0; i < s->count; i++) {
        size_t eol = p;
        while (eol < size && img[eol] != '\n')
            eol++;
        size_t len = p < size ? eol - p : 0;
        free(s->titles[i]);
        s->titles[i] = (char *)malloc(len + 1);
        if (len)
            memcpy(s->titles[i], img + p, len);
        s->titles[i][len] = '\0';
        p = eol + 1;
    }
    free(img);
    return 0;
}
