/*
 * sbwt_b200.h -- C ABI of the B200-native plain-matrix SBWT k-mer query path.
 *
 * This is the drop-in boundary for ONE path of algbio/SBWT (paths relative to
 * the reference root): the plain-matrix index's k-mer membership queries,
 *   SubsetMatrixRank::rank            include/sbwt/SubsetMatrixRank.hh:31-37
 *   SBWT::update_sbwt_interval        include/sbwt/SBWT.hh:423-437
 *   SBWT::search                      include/sbwt/SBWT.hh:390-415
 *   SBWT::streaming_search            include/sbwt/SBWT.hh:545-581
 *   SBWT::load (plain-matrix file)    include/sbwt/SBWT.hh:501-516, SubsetMatrixRank.hh:102-125
 *   get_char_idx / DNA_to_char_idx    include/sbwt/SBWT.hh:49-57, globals.hh:38-47
 * The reference has no FFI; the C++ mirror of its template surface
 * (sbwt_b200/csrc/SBWT.hh) and the `sbwt search` command line
 * (sbwt_b200/csrc/cli_search.cpp) are thin callers of exactly these entry
 * points. INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success and nonzero on failure, with the message available from
 * sbwt_gpu_last_error() (thread-local); no C++ exception crosses the boundary;
 * there is NO CPU fallback -- without a CUDA device every compute entry point
 * fails.  Results are bit-exact with the reference: one int64 per k-mer, the
 * colex rank of its column, or -1.
 */
#ifndef SBWT_B200_H
#define SBWT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBWT_B200_ABI_VERSION 1

/* Which reference entry point a batch call mirrors. */
#define SBWT_GPU_MODE_SEARCH 0    /* SBWT::search on every k-mer start (sbwt_search.cpp:67-91)      */
#define SBWT_GPU_MODE_STREAMING 1 /* SBWT::streaming_search per read   (sbwt_search.cpp:45-65)      */

/* How read bytes are interpreted (the "packer", reference row H). */
#define SBWT_GPU_CASE_UPPER 0 /* a-z folded to A-Z first: what the index sees through `sbwt search`,
                                 where seq_io::Reader upper-cases every base (SeqIO.hh:294-297,330-333) */
#define SBWT_GPU_CASE_EXACT 1 /* only the bytes 'A','C','G','T' are valid: SBWT::search() called
                                 directly on a caller's buffer (SBWT.hh:427, globals.hh:38-47)         */
#define SBWT_GPU_CASE_API 2   /* the direct API on raw bytes, to the letter: MODE_SEARCH = CASE_EXACT; MODE_STREAMING =
                                 SBWT::streaming_search(const char*, len), whose from-scratch searches take the bytes
                                 as they are while a streaming step upper-cases its new character (SBWT.hh:565): a k-mer
                                 covering a lower-case base is a miss unless the k-mer before it was found. On an index
                                 that violates the edge invariant (hand-made files only) it is CASE_EXACT.        */

typedef struct sbwt_gpu_index sbwt_gpu_index;     /* one device-resident index (index, device)         */
typedef struct sbwt_gpu_session sbwt_gpu_session; /* scratch + streams for batches against one index   */

/* Message of the last failure on the calling thread ("" if none). */
const char *sbwt_gpu_last_error(void);
/* Number of CUDA devices, or 0 when there is no usable device / driver. */
int sbwt_gpu_device_count(void);
int sbwt_gpu_abi_version(void);

/* ---- index ------------------------------------------------------------- */

/* Replaces SBWT(A,C,G,T,streaming_support,k,n_kmers,precalc_k) + the four rank_support_v5
 * constructors (SBWT.hh:336-353, rank_support_v5.hpp:65-109): copies the four n_nodes-bit
 * vectors (LSB-first 64-bit words, ceil(n_nodes/64) words each) to `device` and re-lays them
 * there as interleaved count+payload sectors. suffix_group_starts may be NULL (index built with
 * --no-streaming-support). precalc_lr holds 4^precalc_k pairs (l,r) as in
 * SBWT::kmer_prefix_precalc; with precalc_k > 0 and precalc_lr == NULL the table is computed on
 * the device (SBWT::do_kmer_prefix_precalc, SBWT.hh:617-645). The caller keeps its arrays. */
int sbwt_gpu_index_create(const uint64_t *const bits[4], const uint64_t *suffix_group_starts,
                          int64_t n_nodes, int64_t n_kmers, int64_t k, const int64_t C[4],
                          const int64_t *precalc_lr, int64_t precalc_k, int device,
                          sbwt_gpu_index **out);
/* Replaces the variant-string read of sbwt_search.cpp:194-199 + plain_matrix_sbwt_t::load:
 * parses a serialized plain-matrix .sbwt file bit-for-bit and creates the device index. */
int sbwt_gpu_index_load(const char *path, int device, sbwt_gpu_index **out);
void sbwt_gpu_index_destroy(sbwt_gpu_index *idx);

/* Accessors mirroring SBWT.hh:111-157,253. */
int64_t sbwt_gpu_index_k(const sbwt_gpu_index *idx);
int64_t sbwt_gpu_index_n_nodes(const sbwt_gpu_index *idx);   /* number_of_subsets() */
int64_t sbwt_gpu_index_n_kmers(const sbwt_gpu_index *idx);   /* number_of_kmers()   */
int64_t sbwt_gpu_index_precalc_k(const sbwt_gpu_index *idx); /* get_precalc_k()     */
int sbwt_gpu_index_has_streaming_support(const sbwt_gpu_index *idx);
int sbwt_gpu_index_device(const sbwt_gpu_index *idx);
void sbwt_gpu_index_C(const sbwt_gpu_index *idx, int64_t C[4]); /* get_C_array() */
/* Bytes of device memory held by the re-laid index (sectors + tables + suffix-group bits). */
int64_t sbwt_gpu_index_device_bytes(const sbwt_gpu_index *idx);
/* 1 if every non-suffix-group-start column has an empty subset (true for every index the
 * reference builds; lets the streaming step use one sector instead of a walk-back). */
int sbwt_gpu_index_edges_only_at_group_starts(const sbwt_gpu_index *idx);
/* Nonzero if the walk kernels answer ranks from a compact one-hot layout (two bits per column; for narrow indexes in
 * which at most 5 % of the blocks hold a column with no outgoing edge or several; SBWT_B200_COMPACT=0 disables it, =2
 * forces it): 2 = csector64 (64 columns per sector, absolute counts inline), 1 = csector96 (96 columns, relative counts;
 * chosen when only the denser format keeps the structure on chip; SBWT_B200_LAYOUT=c64|c96 overrides). *flagged_fraction
 * (may be NULL) receives the fraction of blocks that are answered from the classic sectors, -1 if the layout was not
 * built. Results never depend on the layout. Replaces nothing in the reference: it is the device-side counterpart of
 * choosing the SubsetMatrixRank bit-vector representation (SubsetMatrixRank.hh:19-37). */
int sbwt_gpu_index_compact_layout(const sbwt_gpu_index *idx, double *flagged_fraction);
/* Bytes of L2 this index asked the device to set aside for its persisting (evict_last) lines: done for a one-hot
 * index whose csectors fit on chip (cudaLimitPersistingL2CacheSize; SBWT_B200_L2_SET_ASIDE_MB overrides, 0 = never). */
int64_t sbwt_gpu_index_l2_set_aside(const sbwt_gpu_index *idx);

/* The walk starts from a device-side table of the intervals of all tp-mers. By default tp is a
 * few characters longer than the file's precalc length (runtime only: built on the device from
 * the bit vectors, never serialized; results are identical because a table row is exactly the
 * interval the walk would reach, SBWT.hh:617-645). set_table_length rebuilds it (0 = no table,
 * every search walks all k characters as with precalc_k == 0). If the file's own table does not
 * follow from its bit vectors only tp == precalc_k is accepted. */
int sbwt_gpu_index_set_table_length(sbwt_gpu_index *idx, int tp);
int sbwt_gpu_index_table_length(const sbwt_gpu_index *idx);

/* get_precalc(): copies the 4^p (l,r) pairs of the device's table to out_lr (2*4^p values). */
int sbwt_gpu_index_get_precalc(const sbwt_gpu_index *idx, int64_t *out_lr);

/* SubsetMatrixRank::rank(pos, c) for n host-side queries (positions in [0, n_nodes], chars
 * as bytes; any byte outside ACGT gives 0). Small-batch diagnostic / parity entry point. */
int sbwt_gpu_rank(sbwt_gpu_index *idx, const int64_t *pos, const char *chars, int64_t n, int64_t *out);

/* ---- the other read-only queries of SBWT.hh, batched ---------------------
 * Host arrays in, host arrays out, one kernel launch per call (device scratch is kept by the index). Strings are
 * ascii[offsets[i] .. offsets[i+1]). Each entry answers with the arithmetic of the reference function it names. */

/* SBWT::update_sbwt_interval (SBWT.hh:423-437) on n (string, interval) pairs: l[i], r[i] are read and overwritten.
 * Raw bytes (only 'A','C','G','T' are valid); an interval with l == -1 passes through; a failure gives {-1,-1}. */
int sbwt_gpu_update_interval_batch(sbwt_gpu_index *idx, const char *ascii, const int64_t *offsets, int64_t n,
                                   int64_t *l, int64_t *r);
/* SBWT::partial_search (SBWT.hh:526-537): the interval of the longest prefix of each string that is found
 * (lower case folded to upper) and its length. */
int sbwt_gpu_partial_search_batch(sbwt_gpu_index *idx, const char *ascii, const int64_t *offsets, int64_t n,
                                  int64_t *l, int64_t *r, int64_t *matched);
/* SBWT::forward (SBWT.hh:369-381): the node reached from nodes[i] over the edge chars[i], or -1. Fails like the
 * reference when the index has no streaming support. */
int sbwt_gpu_forward_batch(sbwt_gpu_index *idx, const int64_t *nodes, const char *chars, int64_t n, int64_t *out);
/* SubsetMatrixRank::contains (SubsetMatrixRank.hh:39-48): out[i] = 1 if column pos[i] has an edge chars[i]. */
int sbwt_gpu_contains_batch(sbwt_gpu_index *idx, const int64_t *pos, const char *chars, int64_t n, uint8_t *out);
/* SBWT::get_kmer (SBWT.hh:701-725): the k-character label of node colex_ranks[i] ('$'-padded on the left) at
 * out[i * k .. (i+1) * k), not NUL-terminated. */
int sbwt_gpu_get_kmer_batch(sbwt_gpu_index *idx, const int64_t *colex_ranks, int64_t n, char *out);
/* SBWT::ascii_export_sets (SBWT.hh:750-773): per column its characters (the last one lower-cased) or '$', then one
 * newline. *n_bytes receives the text length; with out == NULL or capacity too small only the length is returned
 * (and the call fails). At most 4 * n_nodes + 1 bytes. */
int sbwt_gpu_ascii_export_sets(sbwt_gpu_index *idx, char *out, int64_t capacity, int64_t *n_bytes);

/* ---- batches ----------------------------------------------------------- */

/* A session owns the device scratch (packed reads, plan, staging) for batches of at most
 * max_bases read bytes and max_reads reads per device-side launch; host-side calls of any size
 * are chunked to fit. One session per host thread; sessions on one index are independent. */
int sbwt_gpu_session_create(sbwt_gpu_index *idx, int64_t max_bases, int64_t max_reads,
                            sbwt_gpu_session **out);
void sbwt_gpu_session_destroy(sbwt_gpu_session *s);

/* Number of results a batch produces: sum over reads of max(0, len - k + 1). Host arrays. */
int64_t sbwt_gpu_count_outputs(const int64_t *read_offsets, int64_t n_reads, int64_t k);

/* Host-buffer batch (the call `sbwt search` makes): reads are concatenated in `ascii`, read i
 * is ascii[read_offsets[i] .. read_offsets[i+1]) (read_offsets[0] need not be 0). Writes
 * sbwt_gpu_count_outputs() int64 values to `out`, read after read -- the concatenation of the
 * vectors SBWT::streaming_search / the search() loop return. Copies H2D, packs, walks and copies
 * D2H through pinned staging buffers in a double-buffered pipeline. MODE_STREAMING on an index
 * without streaming support fails like the reference ("streaming search support not built").
 * Every read must fit the session (max_bases) and the offsets must not decrease: both are checked chunk by
 * chunk as the batch is cut, so such a read is reported when its chunk is reached; after ANY non-zero return
 * nothing is in flight any more (the call drains its pipeline) and the contents of the result buffers are
 * unspecified. The same holds for the _i32, _hits, _text and _sharded forms below. */
int sbwt_gpu_query_host(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets,
                        int64_t n_reads, int mode, int case_mode, int64_t *out);

/* Hits-only results, for callers that do not need a dense array: hit_mask receives one bit per result, numbered over
 * the whole batch (bit i of word i / 32 = result i is found; (count_outputs + 31) / 32 words), and `hits` (may be
 * NULL: membership only) the found values in result order, *n_hits of them (at most count_outputs; int32, so only for
 * an index with fewer than 2^31 columns -- larger indexes get the bitmap alone). A miss is always -1
 * (SBWT.hh:390-415, :545-581), so sbwt_gpu_query_host's array follows from the two. One bit per k-mer and four bytes
 * per found k-mer cross PCIe, and with a pinned `hits` buffer no host thread touches the results: this is the call
 * that scales with the number of GPUs on one host. */
int sbwt_gpu_query_host_hits(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets, int64_t n_reads,
                             int mode, int case_mode, uint32_t *hit_mask, int32_t *hits, int64_t *n_hits);

/* Same, with int32 results: half the device-to-host bytes (the PCIe copy of the results is what bounds
 * an end-to-end batch). Only for an index with fewer than 2^31 columns (every value, and -1, fits);
 * fails otherwise. Values are the same numbers sbwt_gpu_query_host returns. */
int sbwt_gpu_query_host_i32(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets,
                            int64_t n_reads, int mode, int case_mode, int32_t *out);

/* Device-buffer batch: d_ascii / d_read_offsets / d_out are device pointers on the index's
 * device, read_offsets[0] == 0, n_bases == read_offsets[n_reads] <= max_bases, n_reads <=
 * max_reads; d_out holds n_out == count_outputs values. Enqueues pack -> plan -> walk on
 * `cuda_stream` (a cudaStream_t, NULL = default stream) and returns without synchronising. */
int sbwt_gpu_query_device(sbwt_gpu_session *s, const char *d_ascii, const int64_t *d_read_offsets,
                          int64_t n_reads, int64_t n_bases, int mode, int case_mode,
                          int64_t *d_out, int64_t n_out, void *cuda_stream);

int sbwt_gpu_query_device_i32(sbwt_gpu_session *s, const char *d_ascii, const int64_t *d_read_offsets,
                              int64_t n_reads, int64_t n_bases, int mode, int case_mode,
                              int32_t *d_out, int64_t n_out, void *cuda_stream);

/* The two names of the reference API, as thin wrappers of sbwt_gpu_query_host with CASE_UPPER. */
int sbwt_gpu_search_batch(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets,
                          int64_t n_reads, int64_t *out);
int sbwt_gpu_streaming_batch(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets,
                             int64_t n_reads, int64_t *out);

/* ---- text output ---------------------------------------------------------
 * print_vector (src/CLI/sbwt_search.cpp:21-43) on the device: per read "<value> " for every k-mer and
 * then '\n'; -1 prints as "-1"; any other value <= 0 prints as an empty field, as in the reference. */

/* Bytes that always hold the text of n_values results of this index in n_reads lines. */
int64_t sbwt_gpu_text_capacity(const sbwt_gpu_index *idx, int64_t n_values, int64_t n_reads);

/* Formats device-resident results (int64, or int32 when vals_are_i32) of the reads described by
 * d_read_offsets (as given to sbwt_gpu_query_device; read i yields max(0, len_i - k + 1) values, consecutive
 * in d_vals) into d_text. *d_text_bytes (device, may be NULL) receives the text length; when it exceeds
 * text_capacity nothing is written. n_reads <= the session's max_reads. Asynchronous on cuda_stream. */
int sbwt_gpu_format_device(sbwt_gpu_session *s, const void *d_vals, int vals_are_i32, const int64_t *d_read_offsets,
                           int64_t n_reads, char *d_text, int64_t text_capacity, int64_t *d_text_bytes, void *cuda_stream);

/* Receives consecutive pieces of the output text, in order; a nonzero return aborts the call. */
typedef int (*sbwt_gpu_text_sink)(void *user, const char *text, int64_t n_bytes);

/* run_queries_* + print_vector for a host batch (what `sbwt search` does per file, sbwt_search.cpp:45-91):
 * H2D -> pack -> walk -> format on the device -> D2H of the text, double-buffered; the text reaches `sink`
 * in order. *n_lookups (may be NULL) receives the number of k-mers answered. */
int sbwt_gpu_query_host_text(sbwt_gpu_session *s, const char *ascii, const int64_t *read_offsets, int64_t n_reads,
                             int mode, int case_mode, sbwt_gpu_text_sink sink, void *user, int64_t *n_lookups);

/* Page-locked host memory for the buffers handed to sbwt_gpu_query_host: pinned buffers are
 * DMA'd directly, pageable ones are staged through the session's own pinned buffers. */
int sbwt_gpu_host_alloc(size_t bytes, void **out);
/* The same with the device whose context does the allocation named (a thread that never selected a device would
 * otherwise create a context on device 0). The memory is usable with every device. */
int sbwt_gpu_host_alloc_on(int device, size_t bytes, void **out);
void sbwt_gpu_host_free(void *p);

/* The packer alone (device buffers): 2-bit codes, 32 bases per u64 word, and one invalid bit
 * per base, 32 per u32 word; both arrays need n_bases/32 + 4 words. Asynchronous. */
int sbwt_gpu_pack_device(const char *d_ascii, int64_t n_bases, int case_mode, uint64_t *d_codes,
                         uint32_t *d_invalid, void *cuda_stream);

/* ---- measurement ------------------------------------------------------- */

typedef struct sbwt_gpu_stats {
    int64_t lookups;      /* results produced                                                       */
    int64_t hits;         /* results >= 0                                                           */
    int64_t rank_ops;     /* rank evaluations with the reference's accounting: 2 per interval step  */
    int64_t index_sectors; /* distinct 32-byte index sectors the walk had to read (incl. table rows) */
    int64_t kernel_launches; /* kernels of this library launched by the counted call               */
} sbwt_gpu_stats;

/* Same work as sbwt_gpu_query_device but with counting kernels; synchronises and fills `st`.
 * For the roofline's algorithmic bytes; never inside a timed region. */
int sbwt_gpu_query_device_counted(sbwt_gpu_session *s, const char *d_ascii, const int64_t *d_read_offsets,
                                  int64_t n_reads, int64_t n_bases, int mode, int case_mode,
                                  int64_t *d_out, int64_t n_out, void *cuda_stream, sbwt_gpu_stats *st);
/* Per-kernel timing of device-buffer batches: when enabled, sbwt_gpu_query_device records CUDA
 * events on the launching stream before the packer, before the walk kernel and after it;
 * last_timing waits for the last batch and returns the two intervals in milliseconds. */
int sbwt_gpu_session_set_timing(sbwt_gpu_session *s, int enable);
int sbwt_gpu_session_last_timing(sbwt_gpu_session *s, double *prep_ms, double *walk_ms);
/* Kernels launched by this library on the calling thread since the last reset. */
int64_t sbwt_gpu_launch_count(int reset);

/* Random 32-byte (or 64-byte) sector gather micro-benchmark: `n_loads` independent loads of
 * `bytes_per_load` (32 or 64) at uniformly random aligned offsets inside a device buffer of
 * buffer_bytes, repeated `iters` times; returns the best time in milliseconds. Gives the
 * random-sector ceilings (DRAM-resident and L2-resident buffers) the roofline is quoted against. */
int sbwt_gpu_sector_probe(int device, int64_t buffer_bytes, int64_t n_loads, int bytes_per_load,
                          int iters, double *best_ms);

/* Multi-GPU inside one process (the reference is single-threaded; this is the batched form of running
 * run_queries_streaming / run_queries_not_streaming, sbwt_search.cpp:45-91, over several devices): `sessions`
 * are sessions on replicas of ONE index, normally one per device (sbwt_gpu_index_load(path, device = d)). The
 * reads are cut into n_sessions contiguous ranges of (almost) equal total bases, one host thread per session
 * answers its range with sbwt_gpu_query_host, and the results land in order in `out` -- exactly what one
 * sbwt_gpu_query_host call over the whole batch returns. No collective, nothing is exchanged between devices. */
int sbwt_gpu_query_host_sharded(sbwt_gpu_session *const *sessions, int n_sessions, const char *ascii,
                                const int64_t *read_offsets, int64_t n_reads, int mode, int case_mode, int64_t *out);

/* Host half of the 32-bit result wire format (no device work): sbwt_gpu_query_host on an index with fewer
 * than 2^31 columns lets the kernel write int32, copies those over PCIe and sign-extends them into the
 * caller's int64 array with `threads` host threads (SBWT_B200_WIDEN_THREADS; default = hardware threads /
 * GPUs of the job, at most 10; below 4 the int64 values are copied directly). This entry runs that widening
 * step alone: out[i] = in[i] for i < n. Values are what SBWT::search returns (SBWT.hh:390-415). */
int sbwt_gpu_widen_i32(const int32_t *in, int64_t *out, int64_t n, int threads);
/* Host half of the SPARSE result wire format (the default of sbwt_gpu_query_host / _i32 when host threads are
 * available; SBWT_B200_WIRE=dense selects the format above): the device sends, per chunk, one hit bit per
 * result (masks[g] bit i = result 32 g + i is >= 0), the hits only (packed, int32, result order inside every
 * 4096-result block) and where each block's hits start in `packed` (block_base[b]); a miss is always -1
 * (SBWT.hh:390-415, :545-581), so nothing is lost. This entry rebuilds n results into out (int64 if out_is_i64,
 * else int32) with `threads` host threads; no device work. `packed` must be readable for 32 bytes past its last
 * value (hits are moved eight at a time). */
int sbwt_gpu_expand_sparse(const uint32_t *masks, const uint32_t *block_base, const int32_t *packed, int64_t n,
                           void *out, int out_is_i64, int threads);
/* Host threads this session's sbwt_gpu_query_host calls widen with: 0 = int64 values cross PCIe, -1 = not
 * decided yet (decided by the first sbwt_gpu_query_host call). */
int sbwt_gpu_session_widen_threads(const sbwt_gpu_session *s);

#ifdef __cplusplus
}
#endif
#endif
