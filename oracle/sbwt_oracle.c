/*
 * sbwt_oracle.c -- TEST INFRASTRUCTURE ONLY (see sbwt_oracle.h).
 * Plain-C CPU restatement of the plain-matrix SBWT query path of algbio/SBWT.
 * Parity status: PINNED against the reference (tests/test_oracle.py).
 */
#include "sbwt_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------ load */

static int rd(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }

static int fail(char *err, size_t errlen, const char *msg) {
    if (err && errlen) snprintf(err, errlen, "%s", msg);
    return -1;
}

/* load_string, src/globals.cpp:56-62: [i64 length][bytes]. */
static int load_string(FILE *f, char *dst, size_t cap) {
    int64_t n;
    if (rd(f, &n, 8) || n < 0 || (size_t)n >= cap) return -1;
    if (rd(f, dst, (size_t)n)) return -1;
    dst[n] = 0;
    return 0;
}

/* sdsl int_vector<1>::load, int_vector.hpp:1614-1628: [u64 bit size][capacity/64 words],
 * capacity = size rounded up to 64 (int_vector.hpp:417-420). One zero padding
 * word is kept in memory so rank(size) may touch data[size/64]
 * (memory_management.hpp:351-368). */
static int load_bitvector(FILE *f, uint64_t *len, uint64_t **words) {
    uint64_t n;
    if (rd(f, &n, 8)) return -1;
    uint64_t nw = (n + 63) >> 6;
    uint64_t *w = (uint64_t *)calloc(nw + 1, 8);
    if (!w) return -1;
    if (nw && rd(f, w, nw * 8)) { free(w); return -1; }
    *len = n;
    *words = w;
    return 0;
}

/* rank_support_v5::load (rank_support_v5.hpp:151-155) = int_vector<64>::load:
 * [u64 bit size = 64*W][W words]. */
static int load_u64vector(FILE *f, uint64_t *nwords, uint64_t **words) {
    uint64_t nbits;
    if (rd(f, &nbits, 8) || (nbits & 63)) return -1;
    uint64_t nw = nbits >> 6;
    uint64_t *w = (uint64_t *)calloc(nw + 2, 8);
    if (!w) return -1;
    if (nw && rd(f, w, nw * 8)) { free(w); return -1; }
    *nwords = nw;
    *words = w;
    return 0;
}

/* sbwt_search.cpp:194-199 (variant string) + SBWT::load SBWT.hh:501-516 +
 * SubsetMatrixRank::load SubsetMatrixRank.hh:102-125 + load_std_vector SBWT.hh:451-459. */
int sbwt_oracle_load(const char *path, sbwt_oracle_index *idx, char *err, size_t errlen) {
    memset(idx, 0, sizeof *idx);
    FILE *f = fopen(path, "rb");
    if (!f) return fail(err, errlen, "cannot open index file");
    char s[64];
    int rc = -1;
    if (load_string(f, s, sizeof s)) { fail(err, errlen, "truncated variant string"); goto out; }
    if (strcmp(s, "plain-matrix")) { fail(err, errlen, "not a plain-matrix index"); goto out; }
    if (load_string(f, s, sizeof s)) { fail(err, errlen, "truncated version string"); goto out; }
    if (strcmp(s, "v0.1")) {
        fail(err, errlen, "Error: Corrupt index file, or the index was constructed with an incompatible version of SBWT.");
        goto out;
    }
    for (int c = 0; c < 4; c++)
        if (load_bitvector(f, &idx->bits_len[c], &idx->bits[c])) { fail(err, errlen, "truncated bit vector"); goto out; }
    for (int c = 0; c < 4; c++)
        if (load_u64vector(f, &idx->rs_words[c], &idx->rs[c])) { fail(err, errlen, "truncated rank support"); goto out; }
    if (load_bitvector(f, &idx->sgs_len, &idx->sgs)) { fail(err, errlen, "truncated streaming support"); goto out; }
    int64_t nbytes;
    if (rd(f, &nbytes, 8) || nbytes != 32 || rd(f, idx->C, 32)) { fail(err, errlen, "bad C array"); goto out; }
    if (rd(f, &nbytes, 8) || nbytes < 0 || (nbytes & 15)) { fail(err, errlen, "bad precalc table"); goto out; }
    idx->n_precalc = nbytes / 16;
    idx->precalc = (int64_t *)malloc(nbytes ? (size_t)nbytes : 16);
    if (!idx->precalc || (nbytes && rd(f, idx->precalc, (size_t)nbytes))) { fail(err, errlen, "truncated precalc table"); goto out; }
    if (rd(f, &idx->precalc_k, 8) || rd(f, &idx->n_nodes, 8) || rd(f, &idx->n_kmers, 8) || rd(f, &idx->k, 8)) {
        fail(err, errlen, "truncated trailer");
        goto out;
    }
    if (fgetc(f) != EOF) { fail(err, errlen, "trailing bytes after index"); goto out; }
    for (int c = 0; c < 4; c++)
        if ((int64_t)idx->bits_len[c] != idx->n_nodes) { fail(err, errlen, "bit vector length != n_nodes"); goto out; }
    if (idx->precalc_k < 0 || idx->precalc_k > 20 || idx->n_precalc != (idx->precalc_k ? (int64_t)1 << (2 * idx->precalc_k) : 0)) {
        fail(err, errlen, "precalc table size does not match precalc_k");
        goto out;
    }
    rc = 0;
out:
    fclose(f);
    if (rc) sbwt_oracle_free(idx);
    return rc;
}

void sbwt_oracle_free(sbwt_oracle_index *idx) {
    for (int c = 0; c < 4; c++) { free(idx->bits[c]); free(idx->rs[c]); }
    free(idx->sgs);
    free(idx->precalc);
    memset(idx, 0, sizeof *idx);
}

/* ------------------------------------------------------------------ rank */

/* get_char_idx, SBWT.hh:49-57 / DNA_to_char_idx, globals.hh:38-47 (case-sensitive). */
static inline int char_idx(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }

/* rank_support_v5<1,1>::rank, rank_support_v5.hpp:116-134: superblock of 2048 bits
 * holds an absolute count and five 11-bit relative counts of 384-bit blocks at
 * shifts 48,36,24,12,0; then <= 5 full words and one masked word
 * (rank_support.hpp:148-156). */
static int64_t rank_v5(const uint64_t *data, const uint64_t *bb, uint64_t pos) {
    const uint64_t *p = bb + ((pos >> 10) & 0xFFFFFFFFFFFFFFFEULL);
    uint64_t res = p[0] + ((p[1] >> (60 - 12 * ((pos & 0x7FF) / 384))) & 0x7FFULL);
    uint64_t off = pos & 63;
    res += off ? (uint64_t)popc64(data[pos >> 6] & ((1ULL << off) - 1)) : 0; /* word_rank */
    uint64_t w = pos >> 6;                /* index of the partial word */
    unsigned todo = (unsigned)((w & 31) % 6);
    while (todo) { res += (uint64_t)popc64(data[--w]); todo--; }
    return (int64_t)res;
}

int64_t sbwt_oracle_rank(const sbwt_oracle_index *idx, int64_t pos, char c) {
    int ci = char_idx(c);
    if (ci < 0) return 0; /* SubsetMatrixRank.hh:36 */
    return rank_v5(idx->bits[ci], idx->rs[ci], (uint64_t)pos);
}

int64_t sbwt_oracle_rank_naive(const sbwt_oracle_index *idx, int64_t pos, char c) {
    int ci = char_idx(c);
    if (ci < 0) return 0;
    const uint64_t *d = idx->bits[ci];
    int64_t r = 0;
    for (int64_t w = 0; w < (pos >> 6); w++) r += popc64(d[w]);
    if (pos & 63) r += popc64(d[pos >> 6] & ((1ULL << (pos & 63)) - 1));
    return r;
}

/* ---------------------------------------------------------------- search */

/* SBWT.hh:423-437. Note the validity test is on the raw byte (line 427). */
void sbwt_oracle_update_interval(const sbwt_oracle_index *idx, const char *s, int64_t len,
                                 int64_t *l, int64_t *r) {
    if (*l == -1) return;
    for (int64_t i = 0; i < len; i++) {
        int ci = char_idx(s[i]);
        if (ci < 0) { *l = *r = -1; return; }
        *l = idx->C[ci] + rank_v5(idx->bits[ci], idx->rs[ci], (uint64_t)*l);
        *r = idx->C[ci] + rank_v5(idx->bits[ci], idx->rs[ci], (uint64_t)(*r + 1)) - 1;
        if (*l > *r) { *l = *r = -1; return; }
    }
}

/* SBWT.hh:390-415. */
int64_t sbwt_oracle_search(const sbwt_oracle_index *idx, const char *kmer) {
    int64_t l, r;
    if (idx->precalc_k > 0) {
        uint64_t pi = 0;
        for (int64_t i = 0; i < idx->precalc_k; i++) {
            int ci = char_idx(kmer[idx->precalc_k - 1 - i]);
            if (ci < 0) return -1;
            pi = (pi << 2) | (uint64_t)ci;
        }
        l = idx->precalc[2 * pi];
        r = idx->precalc[2 * pi + 1];
        sbwt_oracle_update_interval(idx, kmer + idx->precalc_k, idx->k - idx->precalc_k, &l, &r);
    } else {
        l = 0;
        r = idx->n_nodes - 1;
        sbwt_oracle_update_interval(idx, kmer, idx->k, &l, &r);
    }
    if (l != r) {
        fprintf(stderr, "Bug: k-mer search did not give a singleton interval: %lld %lld\n", (long long)l, (long long)r);
        exit(1);
    }
    return l;
}

/* ------------------------------------------------------ index from arrays */

/* rank_support_v5's constructor (rank_support_v5.hpp:65-109): per 2048-bit superblock one word with the ones
 * before it and one word with the ones of its first 6, 12, 18, 24, 30 words at shifts 48, 36, 24, 12, 0. */
static uint64_t *build_v5(const uint64_t *data, uint64_t nbits, uint64_t *nwords_out) {
    const uint64_t W = (nbits + 63) / 64, nsb = (W * 64 >> 11) + 1;
    uint64_t *bb = (uint64_t *)calloc(nsb * 2, 8);
    if (!bb) return NULL;
    uint64_t total = 0;
    for (uint64_t sb = 0; sb < nsb; sb++) {
        bb[2 * sb] = total;
        uint64_t second = 0, sum = 0;
        for (uint64_t j = 0; j < 32; j++) {
            const uint64_t wi = 32 * sb + j;
            if (j && j % 6 == 0 && wi <= W) second |= sum << (60 - 12 * (j / 6));
            if (wi < W) sum += (uint64_t)popc64(data[wi]);
        }
        bb[2 * sb + 1] = second;
        total += sum;
    }
    *nwords_out = nsb * 2;
    return bb;
}

/* SBWT(A, C, G, T, streaming_support, k, n_kmers, precalc_k) (SBWT.hh:336-353): copies the four n_nodes-bit vectors
 * (and suffix_group_starts, may be NULL), builds the rank directories, the C array (:345-349) and the p-mer table
 * (do_kmer_prefix_precalc, :617-645). For indexes that exist only in memory (tests of very large layouts). */
int sbwt_oracle_from_arrays(sbwt_oracle_index *idx, const uint64_t *const bits[4], const uint64_t *sgs, int64_t n_nodes,
                            int64_t n_kmers, int64_t k, int64_t precalc_k) {
    memset(idx, 0, sizeof *idx);
    const uint64_t W = ((uint64_t)n_nodes + 63) / 64;
    idx->n_nodes = n_nodes; idx->n_kmers = n_kmers; idx->k = k; idx->precalc_k = precalc_k;
    for (int c = 0; c < 4; c++) {
        idx->bits_len[c] = (uint64_t)n_nodes;
        idx->bits[c] = (uint64_t *)calloc(W + 1, 8); /* one zero padding word: rank(n_nodes) may read it */
        if (!idx->bits[c]) { sbwt_oracle_free(idx); return -1; }
        memcpy(idx->bits[c], bits[c], W * 8);
        idx->rs[c] = build_v5(idx->bits[c], (uint64_t)n_nodes, &idx->rs_words[c]);
        if (!idx->rs[c]) { sbwt_oracle_free(idx); return -1; }
    }
    if (sgs) {
        idx->sgs_len = (uint64_t)n_nodes;
        idx->sgs = (uint64_t *)calloc(W + 1, 8);
        if (!idx->sgs) { sbwt_oracle_free(idx); return -1; }
        memcpy(idx->sgs, sgs, W * 8);
    }
    idx->C[0] = 1; /* one ghost dollar into the root */
    for (int c = 0; c < 3; c++) idx->C[c + 1] = idx->C[c] + rank_v5(idx->bits[c], idx->rs[c], (uint64_t)n_nodes);
    if (precalc_k > 0) {
        static const char alphabet[4] = {'A', 'C', 'G', 'T'};
        idx->n_precalc = (int64_t)1 << (2 * precalc_k);
        idx->precalc = (int64_t *)malloc((size_t)idx->n_precalc * 16);
        if (!idx->precalc) { sbwt_oracle_free(idx); return -1; }
        char kmer[32];
        for (int64_t i = 0; i < idx->n_precalc; i++) {
            for (int64_t j = 0; j < precalc_k; j++) kmer[j] = alphabet[(i >> (2 * j)) & 3]; /* first character = lowest digit */
            int64_t l = 0, r = n_nodes - 1;
            sbwt_oracle_update_interval(idx, kmer, precalc_k, &l, &r);
            idx->precalc[2 * i] = l;
            idx->precalc[2 * i + 1] = r;
        }
    }
    return 0;
}

/* ------------------------------------------------- other read-only queries */

/* SubsetMatrixRank::contains, SubsetMatrixRank.hh:39-48. */
int sbwt_oracle_contains(const sbwt_oracle_index *idx, int64_t pos, char c) {
    int ci = char_idx(c);
    if (ci < 0) return 0;
    return (int)((idx->bits[ci][pos >> 6] >> (pos & 63)) & 1);
}

/* SBWT::forward, SBWT.hh:369-381. Returns -2 when the index has no streaming
 * support (the reference throws). */
int64_t sbwt_oracle_forward(const sbwt_oracle_index *idx, int64_t node, char c) {
    if (idx->sgs_len == 0) return -2;
    while (!((idx->sgs[node >> 6] >> (node & 63)) & 1)) node--; /* the first node is always marked */
    int64_t r1 = sbwt_oracle_rank(idx, node, c), r2 = sbwt_oracle_rank(idx, node + 1, c);
    if (r1 == r2) return -1;
    return idx->C[char_idx(c)] + r1;
}

/* SBWT::partial_search, SBWT.hh:526-537: the interval of the longest prefix of
 * input that is found, and its length. */
int64_t sbwt_oracle_partial_search(const sbwt_oracle_index *idx, const char *input, int64_t len,
                                   int64_t *l_out, int64_t *r_out) {
    int64_t l = 0, r = idx->n_nodes - 1;
    for (int64_t i = 0; i < len; i++) {
        char c = (input[i] >= 'a' && input[i] <= 'z') ? (char)(input[i] - 32) : input[i]; /* toupper */
        int64_t nl = l, nr = r;
        sbwt_oracle_update_interval(idx, &c, 1, &nl, &nr);
        if (nl == -1) { *l_out = l; *r_out = r; return i; }
        l = nl; r = nr;
    }
    *l_out = l; *r_out = r;
    return len;
}

/* SBWT::get_kmer, SBWT.hh:701-725: the label of a node, '$'-padded on the
 * left; the backward step is a binary search over rank (no select support). */
void sbwt_oracle_get_kmer(const sbwt_oracle_index *idx, int64_t colex_rank, char *buf) {
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    for (int64_t i = 0; i < idx->k; i++) {
        if (colex_rank == 0) {
            buf[idx->k - 1 - i] = '$';
        } else {
            int ci = 0;
            while (ci + 1 < 4 && colex_rank >= idx->C[ci + 1]) ci++;
            char c = alphabet[ci];
            buf[idx->k - 1 - i] = c;
            int64_t rel = colex_rank - idx->C[ci], p = 0, step = idx->n_nodes;
            while (step > 0) {
                while (p + step <= idx->n_nodes && sbwt_oracle_rank(idx, p + step, c) <= rel) p += step;
                step /= 2;
            }
            colex_rank = p;
        }
    }
}

/* SBWT::ascii_export_sets, SBWT.hh:750-773: per column its characters (the last
 * one lower-cased) or '$' for an empty set, then one newline. buf must hold
 * 4 * n_nodes + 1 bytes; returns the bytes written. */
int64_t sbwt_oracle_export_sets(const sbwt_oracle_index *idx, char *buf) {
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    int64_t n = 0;
    for (int64_t col = 0; col < idx->n_nodes; col++) {
        int64_t start = n;
        for (int ci = 0; ci < 4; ci++)
            if ((idx->bits[ci][col >> 6] >> (col & 63)) & 1) buf[n++] = alphabet[ci];
        if (n == start) buf[n++] = '$';
        else buf[n - 1] = (char)(buf[n - 1] + 32); /* tolower */
    }
    buf[n++] = '\n';
    return n;
}

static inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

/* SBWT.hh:545-581. */
int64_t sbwt_oracle_streaming_search(const sbwt_oracle_index *idx, const char *input,
                                     int64_t len, int64_t *out) {
    if (idx->sgs_len == 0) return -2;
    int64_t k = idx->k, n = 0;
    if (len < k) return 0;
    out[n++] = sbwt_oracle_search(idx, input);
    for (int64_t i = 1; i < len - k + 1; i++) {
        if (out[n - 1] == -1) {
            out[n++] = sbwt_oracle_search(idx, input + i);
        } else {
            int64_t column = out[n - 1];
            while (((idx->sgs[column >> 6] >> (column & 63)) & 1) == 0) column--;
            char c = up(input[i + k - 1]); /* toupper, SBWT.hh:565 */
            int ci = char_idx(c);
            if (ci < 0) {
                out[n++] = -1;
            } else {
                int64_t nl = idx->C[ci] + rank_v5(idx->bits[ci], idx->rs[ci], (uint64_t)column);
                int64_t nr = idx->C[ci] + rank_v5(idx->bits[ci], idx->rs[ci], (uint64_t)(column + 1)) - 1;
                out[n++] = (nl == nr) ? nl : -1;
            }
        }
    }
    return n;
}

/* sbwt_search.cpp:74-88. */
int64_t sbwt_oracle_search_all(const sbwt_oracle_index *idx, const char *input,
                               int64_t len, int64_t *out) {
    int64_t n = 0;
    for (int64_t i = 0; i < len - idx->k + 1; i++) out[n++] = sbwt_oracle_search(idx, input + i);
    return n;
}

int64_t sbwt_oracle_query_batch(const sbwt_oracle_index *idx, const char *ascii,
                                const int64_t *offsets, int64_t n_reads, int streaming,
                                int64_t *out) {
    int64_t n = 0, cap = 0;
    char *buf = NULL; /* the reference works on NUL-terminated copies (SeqIO read_buf) */
    for (int64_t i = 0; i < n_reads; i++) {
        int64_t len = offsets[i + 1] - offsets[i];
        if (len + 1 > cap) { cap = 2 * (len + 1); buf = (char *)realloc(buf, (size_t)cap); }
        memcpy(buf, ascii + offsets[i], (size_t)len);
        buf[len] = 0;
        int64_t m = streaming ? sbwt_oracle_streaming_search(idx, buf, len, out + n)
                              : sbwt_oracle_search_all(idx, buf, len, out + n);
        if (m < 0) { free(buf); return m; }
        n += m;
    }
    free(buf);
    return n;
}

/* ---------------------------------------------------------------- output */

/* print_vector, sbwt_search.cpp:21-43. A value of 0 prints as an empty field. */
size_t sbwt_oracle_format_line(const int64_t *v, int64_t n, char *buf) {
    char *p = buf;
    for (int64_t j = 0; j < n; j++) {
        int64_t x = v[j];
        char tmp[32];
        int i = 0;
        if (x == -1) {
            tmp[0] = '1'; tmp[1] = '-'; i = 2;
        } else {
            while (x > 0) { tmp[i++] = (char)('0' + (x % 10)); x /= 10; }
        }
        while (i) *p++ = tmp[--i];
        *p++ = ' ';
    }
    *p++ = '\n';
    return (size_t)(p - buf);
}

/* ----------------------------------------------------------------- SeqIO */

/* Buffered_ifstream::get / getline / eof, buffered_streams.hh:67-111, over a
 * whole-file buffer. eof becomes true only after a get() past the end. */
typedef struct { char *data; size_t size, pos; int is_eof; } bytestream;

static int bs_get(bytestream *s, char *c) {
    if (s->is_eof) return 0;
    if (s->pos < s->size) { *c = s->data[s->pos++]; return 1; }
    s->is_eof = 1;
    return 0;
}

typedef struct { char *p; size_t n, cap; } strbuf;
static void sb_push(strbuf *b, char c) {
    if (b->n + 1 > b->cap) { b->cap = b->cap ? 2 * b->cap : 256; b->p = (char *)realloc(b->p, b->cap); }
    b->p[b->n++] = c;
}

static int bs_getline(bytestream *s, strbuf *line) {
    line->n = 0;
    for (;;) {
        char c = 0;
        bs_get(s, &c);
        if (s->is_eof) return line->n > 0;
        if (c == '\n') return 1;
        sb_push(line, c);
    }
}

static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && !strcmp(s + a - b, suf);
}

/* figure_out_file_format, SeqIO.hh:62-93. Returns 0 FASTA, 1 FASTQ, -1 unknown. */
static int file_format(const char *name, int *gz) {
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s", name);
    *gz = 0;
    if (ends_with(tmp, ".gz")) { tmp[strlen(tmp) - 3] = 0; *gz = 1; }
    const char *dot = strrchr(tmp, '.');
    if (!dot) return -1;
    static const char *fa[] = {".fasta", ".fna", ".ffn", ".faa", ".frn", ".fa"};
    static const char *fq[] = {".fastq", ".fq"};
    for (size_t i = 0; i < 6; i++) if (!strcmp(dot, fa[i])) return 0;
    for (size_t i = 0; i < 2; i++) if (!strcmp(dot, fq[i])) return 1;
    return -1;
}

static int slurp(const char *path, int gz, bytestream *s) {
    memset(s, 0, sizeof *s);
    size_t cap = 1 << 20;
    s->data = (char *)malloc(cap);
    if (gz) {
        gzFile g = gzopen(path, "rb");
        if (!g) return -1;
        for (;;) {
            if (s->size == cap) { cap *= 2; s->data = (char *)realloc(s->data, cap); }
            int n = gzread(g, s->data + s->size, (unsigned)((cap - s->size) > (1u << 30) ? (1u << 30) : (cap - s->size)));
            if (n < 0) { gzclose(g); return -1; }
            if (n == 0) break;
            s->size += (size_t)n;
        }
        gzclose(g);
    } else {
        FILE *f = fopen(path, "rb");
        if (!f) return -1;
        for (;;) {
            if (s->size == cap) { cap *= 2; s->data = (char *)realloc(s->data, cap); }
            size_t n = fread(s->data + s->size, 1, cap - s->size, f);
            if (n == 0) break;
            s->size += n;
        }
        fclose(f);
    }
    return 0;
}

/* Reader::get_next_read_to_buffer, SeqIO.hh:255-360. Returns the read length,
 * 0 at end of file, -1 on the conditions where the reference throws. The
 * upper-cased read is left in seq. */
static int64_t next_read(bytestream *s, int fastq, strbuf *seq, strbuf *tmp, char *err, size_t errlen) {
    if (s->is_eof) return 0;
    if (!fastq) {
        seq->n = 0;
        char c = 0;
        bs_getline(s, tmp); /* header */
        if (s->is_eof) return fail(err, errlen, "FASTA file ended unexpectedly.");
        bs_get(s, &c);
        if (c == '\n') return fail(err, errlen, "Empty line in FASTA file.");
        if (c == '>') return fail(err, errlen, "Empty sequence in FASTA file.");
        while (c != '>') {
            sb_push(seq, c);
            bs_getline(s, tmp);
            if (s->is_eof) return fail(err, errlen, "FASTA file ended unexpectedly.");
            for (size_t i = 0; i < tmp->n; i++) sb_push(seq, tmp->p[i]);
            bs_get(s, &c);
            if (c == '\n') return fail(err, errlen, "Empty line inside sequence in file.");
            if (s->is_eof) break;
        }
    } else {
        strbuf *lines[4] = {tmp, seq, tmp, tmp};
        for (int i = 0; i < 4; i++) {
            bs_getline(s, lines[i]);
            if (s->is_eof) return fail(err, errlen, "FASTQ file ended unexpectedly.");
        }
        char c;
        bs_get(s, &c); /* '@' of the next record, or sets eof */
        if (seq->n == 0) return fail(err, errlen, "Error: empty sequence in FASTQ file.");
    }
    for (size_t i = 0; i < seq->n; i++) seq->p[i] = up(seq->p[i]); /* upper_case_table, SeqIO.hh:44-45 */
    sb_push(seq, 0);
    seq->n--;
    return (int64_t)seq->n;
}

/* run_file / run_queries_streaming / run_queries_not_streaming,
 * sbwt_search.cpp:45-105, plain-text output. */
int64_t sbwt_oracle_search_file(const sbwt_oracle_index *idx, const char *query_path,
                                const char *out_path, char *err, size_t errlen) {
    int gz = 0, fmt = file_format(query_path, &gz);
    if (fmt < 0) return fail(err, errlen, "Unknown file format");
    bytestream s;
    if (slurp(query_path, gz, &s)) { free(s.data); return fail(err, errlen, "Error opening file"); }
    FILE *out = fopen(out_path, "wb");
    if (!out) { free(s.data); return fail(err, errlen, "cannot open output file"); }
    int64_t nq = -1;
    strbuf seq = {0, 0, 0}, tmp = {0, 0, 0};
    int64_t *vals = NULL; char *text = NULL; int64_t cap = 0;
    char c = 0;
    bs_get(&s, &c); /* read_first_char_and_sanity_check, SeqIO.hh:178-189 */
    if ((fmt == 0 && c != '>') || (fmt == 1 && c != '@')) {
        fail(err, errlen, fmt == 0 ? "ERROR: FASTA file does not start with '>'" : "ERROR: FASTQ file does not start with '@'");
        goto done;
    }
    nq = 0;
    for (;;) {
        int64_t len = next_read(&s, fmt, &seq, &tmp, err, errlen);
        if (len < 0) { nq = -1; goto done; }
        if (len == 0) break;
        int64_t m = len - idx->k + 1;
        if (m < 0) m = 0;
        if (m + 1 > cap) {
            cap = 2 * (m + 1);
            vals = (int64_t *)realloc(vals, (size_t)cap * 8);
            text = (char *)realloc(text, (size_t)cap * 21 + 2);
        }
        int64_t n = idx->sgs_len ? sbwt_oracle_streaming_search(idx, seq.p, len, vals)
                                 : sbwt_oracle_search_all(idx, seq.p, len, vals);
        nq += n;
        size_t nb = sbwt_oracle_format_line(vals, n, text);
        fwrite(text, 1, nb, out);
    }
done:
    fclose(out);
    free(s.data); free(seq.p); free(tmp.p); free(vals); free(text);
    return nq;
}
