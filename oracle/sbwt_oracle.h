/*
 * sbwt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the plain-matrix SBWT k-mer
 * query path of algbio/SBWT.  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg can check the CUDA path; nothing under
 * sbwt_b200/ links, imports or calls it, and the product has no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against (1) the reference's only hard-coded known answer for this path
 * (tests/test_CLI.hh:90), (2) outputs of the reference's own classes compiled
 * from /root/reference into oracle/_ref (see oracle/Makefile, ref_driver.cpp)
 * and committed as fixtures under tests/golden/.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference root).
 */
#ifndef SBWT_ORACLE_H
#define SBWT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sbwt_oracle_index {
    int64_t n_nodes;    /* number of columns (subsets)              SBWT.hh:43  */
    int64_t n_kmers;    /*                                          SBWT.hh:44  */
    int64_t k;          /*                                          SBWT.hh:45  */
    int64_t precalc_k;  /* p; 0 = no table                          SBWT.hh:41  */
    int64_t C[4];       /* cumulative character counts, C[0] = 1    SBWT.hh:38  */
    uint64_t bits_len[4];   /* bit length of A,C,G,T vectors (== n_nodes) */
    uint64_t *bits[4];      /* LSB-first words, one zero padding word appended */
    uint64_t rs_words[4];   /* number of 64-bit words of each rank_support_v5 */
    uint64_t *rs[4];        /* the reference's own serialized rank directory */
    uint64_t sgs_len;       /* suffix_group_starts bit length, 0 if absent */
    uint64_t *sgs;
    int64_t n_precalc;      /* 4^p entries */
    int64_t *precalc;       /* pairs (l, r), entry i at [2i], [2i+1] */
} sbwt_oracle_index;

/* Load a serialized plain-matrix .sbwt file (variant string included).
 * Returns 0 on success; on failure returns nonzero and writes a message. */
int sbwt_oracle_load(const char *path, sbwt_oracle_index *idx, char *err, size_t errlen);
/* An index from four bit vectors in memory (SBWT.hh:336-353): rank directories (rank_support_v5.hpp:65-109), C array
 * and p-mer table are built here. sgs may be NULL. Returns 0 on success. */
int sbwt_oracle_from_arrays(sbwt_oracle_index *idx, const uint64_t *const bits[4], const uint64_t *sgs, int64_t n_nodes,
                            int64_t n_kmers, int64_t k, int64_t precalc_k);
void sbwt_oracle_free(sbwt_oracle_index *idx);

/* rank_c(pos) through the rank_support_v5 arithmetic and the directory that
 * was stored in the file (rank_support_v5.hpp:116-134). c is 'A','C','G','T';
 * any other byte returns 0 (SubsetMatrixRank.hh:31-37). */
int64_t sbwt_oracle_rank(const sbwt_oracle_index *idx, int64_t pos, char c);
/* Same quantity by definition: number of 1-bits in [0,pos). O(pos/64). */
int64_t sbwt_oracle_rank_naive(const sbwt_oracle_index *idx, int64_t pos, char c);

/* SBWT::update_sbwt_interval (SBWT.hh:423-437). */
void sbwt_oracle_update_interval(const sbwt_oracle_index *idx, const char *s, int64_t len,
                                 int64_t *l, int64_t *r);
/* SBWT::search(const char*) (SBWT.hh:390-415): colex rank or -1. Reads k bytes. */
int64_t sbwt_oracle_search(const sbwt_oracle_index *idx, const char *kmer);
/* SBWT::streaming_search (SBWT.hh:545-581). out must hold max(0,len-k+1)
 * values. Returns the number of values written, or -2 when the index has no
 * streaming support (the reference throws std::runtime_error there). */
int64_t sbwt_oracle_streaming_search(const sbwt_oracle_index *idx, const char *input,
                                     int64_t len, int64_t *out);
/* The per-read loop of run_queries_not_streaming (sbwt_search.cpp:67-91):
 * search() on every k-mer start. Returns the number of values written. */
int64_t sbwt_oracle_search_all(const sbwt_oracle_index *idx, const char *input,
                               int64_t len, int64_t *out);

/* Batch helper for tests and the cpu_baseline leg: reads are concatenated in
 * `ascii`, read i occupying [offsets[i], offsets[i+1]).  Writes
 * sum(max(0,len_i-k+1)) values. streaming != 0 selects streaming_search.
 * Returns the number of values, or a negative error code. */
int64_t sbwt_oracle_query_batch(const sbwt_oracle_index *idx, const char *ascii,
                                const int64_t *offsets, int64_t n_reads, int streaming,
                                int64_t *out);

/* The other read-only queries of the index. */
int sbwt_oracle_contains(const sbwt_oracle_index *idx, int64_t pos, char c);      /* SubsetMatrixRank.hh:39-48 */
int64_t sbwt_oracle_forward(const sbwt_oracle_index *idx, int64_t node, char c);  /* SBWT.hh:369-381; -2 = no streaming support */
int64_t sbwt_oracle_partial_search(const sbwt_oracle_index *idx, const char *input, int64_t len,
                                   int64_t *l_out, int64_t *r_out);               /* SBWT.hh:526-537; returns the matched length */
void sbwt_oracle_get_kmer(const sbwt_oracle_index *idx, int64_t colex_rank, char *buf);   /* SBWT.hh:701-725; k bytes */
int64_t sbwt_oracle_export_sets(const sbwt_oracle_index *idx, char *buf);          /* SBWT.hh:750-773; 4 n_nodes + 1 bytes */

/* print_vector (sbwt_search.cpp:21-43): "<v> " per value then '\n'.
 * Returns the number of bytes written to buf (buf must hold 21*n+1 bytes). */
size_t sbwt_oracle_format_line(const int64_t *v, int64_t n, char *buf);

/* `sbwt search` for one (query file, output file) pair
 * (sbwt_search.cpp:45-105 + SeqIO.hh:255-360): FASTA/FASTQ by extension, .gz
 * input through zlib, plain-text output only. Returns the number of queries or
 * a negative value with a message in err. */
int64_t sbwt_oracle_search_file(const sbwt_oracle_index *idx, const char *query_path,
                                const char *out_path, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
