/*
 * sbwt_oracle_main.c -- TEST INFRASTRUCTURE ONLY.
 * Command-line front end of the plain-C oracle:
 *   sbwt_oracle search -i index.sbwt -q reads.(fa|fq)[.gz] -o out.txt
 * mirrors `sbwt search` (src/CLI/sbwt_search.cpp:143-260) for one plain-matrix
 * index and one query file, plain-text output.
 */
#include <stdio.h>
#include <string.h>

#include "sbwt_oracle.h"

static const char *arg(int argc, char **argv, const char *name) {
    for (int i = 0; i + 1 < argc; i++)
        if (!strcmp(argv[i], name)) return argv[i + 1];
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 2 || strcmp(argv[1], "search")) {
        fprintf(stderr, "usage: sbwt_oracle search -i index -q queries -o out\n");
        return 1;
    }
    const char *i = arg(argc, argv, "-i"), *q = arg(argc, argv, "-q"), *o = arg(argc, argv, "-o");
    if (!i || !q || !o) {
        fprintf(stderr, "missing -i/-q/-o\n");
        return 1;
    }
    char err[256];
    sbwt_oracle_index idx;
    if (sbwt_oracle_load(i, &idx, err, sizeof err)) {
        fprintf(stderr, "Runtime error: %s\n", err);
        return 1;
    }
    long long n = sbwt_oracle_search_file(&idx, q, o, err, sizeof err);
    sbwt_oracle_free(&idx);
    if (n < 0) {
        fprintf(stderr, "Runtime error: %s\n", err);
        return 1;
    }
    fprintf(stderr, "queries: %lld\n", n);
    return 0;
}
