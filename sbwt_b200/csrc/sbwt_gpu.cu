// sbwt_gpu.cu -- the extern "C" layer of include/sbwt_b200.h: index creation, sessions,
// kernel launches and the host<->device pipeline. No torch types, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sbwt_b200.h"
#include "aux_kernels.cuh"
#include "device_index.cuh"
#include "format_kernels.cuh"
#include "host_widen.hpp"
#include "long_kmer_kernels.cuh"
#include "query_kernels.cuh"
#include "sbwt_file.hpp"
#include "walk_kernel.cuh"

using namespace sbwt_b200;

#ifndef SBWT_B200_DEFAULT_LAYOUT
#define SBWT_B200_DEFAULT_LAYOUT LAY_C64
#endif
static constexpr int kDefaultCompactLayout = SBWT_B200_DEFAULT_LAYOUT; // csector format of eligible (one-hot, narrow) indexes

// ------------------------------------------------------------------ errors

static thread_local std::string g_last_error;
static thread_local int64_t g_launches = 0;

static int set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

#define LAUNCHED() (g_launches++)

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------ objects

struct sbwt_gpu_index {
    int device = 0;
    int64_t n_nodes = 0, n_kmers = 0, k = 0, precalc_k = 0;
    int64_t C[4] = {0, 0, 0, 0};
    bool has_sgs = false;
    int sm_count = 0;
    int64_t device_bytes = 0;
    DeviceIndexView view{};
    void* d_sectors = nullptr;
    void* d_compact = nullptr; // one-hot layout (device_index.cuh), nullptr if the index is not eligible
    void* d_cbase = nullptr;
    double flagged_fraction = -1.0; // blocks of the compact layout that fall back to the classic sectors (-1: not built)
    void* d_sbbase = nullptr;
    void* d_precalc = nullptr;
    void* d_table = nullptr;
    int64_t table_bytes = 0;
    bool table_from_bits = true; // the file's table equals what the bit vectors imply, so any table length is admissible
    bool compact_in_search = false; // the per-k-mer search path also walks the compact layout (default: classic sectors)
    int64_t l2_set_aside = 0;       // bytes of L2 set aside for persisting (evict_last) lines by this index, 0 = none
    uint32_t probe_stride = 0;      // streaming walk: distance between the probes of a range of presumed misses (0 = none)
    void* d_sgs = nullptr;
    // grow-only device scratch of the small-batch entry points (rank, forward, partial_search ...): no cudaMalloc per call
    std::mutex scratch_mutex;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

struct Scratch {
    int64_t max_bases = 0, max_reads = 0, max_items = 0;
    uint64_t* codes = nullptr;
    uint32_t* invalid = nullptr;
    uint32_t* lower = nullptr;  // CASE_API only: one bit per base, the byte is a lower-case a/c/g/t (allocated on first use)
    int64_t* n_out = nullptr;   // per read; becomes the exclusive scan
    int64_t* n_win = nullptr;   // per read; becomes the exclusive scan
    int64_t* partials = nullptr;
    int64_t* totals = nullptr;  // [0] = total results, [1] = total items
    WalkItem* items = nullptr;
    int64_t n_words = 0;        // u64 words allocated for codes (u32 words for invalid)
    unsigned long long* stats = nullptr;
};

struct HostSlot {
    Scratch sc;
    char* d_ascii = nullptr;
    int64_t* d_offsets = nullptr;
    int64_t* d_out = nullptr;
    char* d_text = nullptr;       // print_vector text of the slot's batch (text queries only)
    int64_t text_capacity = 0;
    char* h_ascii = nullptr;      // pinned staging, used only for pageable caller buffers
    int64_t* h_offsets = nullptr;
    int64_t* h_out = nullptr;
    int64_t* h_totals = nullptr;  // pinned copy of sc.totals
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    // what to do when the slot's work has finished
    int64_t out_bytes = 0;
    void* out_dst = nullptr;
    bool out_staged = false;
    // 32-bit wire format (host_widen.hpp): pinned int32 staging + the job a stream callback hands to the pool
    int32_t* h_out32 = nullptr;
    struct WidenJob {
        WidenPool* pool = nullptr;
        const int32_t* src = nullptr;
        int64_t* dst = nullptr;
        size_t n = 0;
        WidenTicket* ticket = nullptr;
    };
    std::vector<WidenJob> widen_jobs; // one per D2H piece of the slot's batch; stable until slot_finish
    WidenPool* widen_pool = nullptr;  // non-null while widening jobs of this slot may be outstanding
    WidenTicket widen_ticket;
    // sparse wire format (aux_kernels.cuh sparse_pack_kernel, host_widen.hpp): masks + block bases + hit count come
    // back first (ev_small); the hits follow once the host knows how many there are (slot_back)
    uint32_t* d_masks = nullptr;
    uint32_t* d_bbase = nullptr;
    unsigned long long* d_total = nullptr;
    uint32_t* h_masks = nullptr;
    uint32_t* h_bbase = nullptr;
    unsigned long long* h_total = nullptr;
    cudaEvent_t ev_small = nullptr;
    bool back_pending = false;
    // hits-only results (sbwt_gpu_query_host_hits): per-block hit counts / bases, the chunk's first mask word
    int64_t* d_bcount = nullptr;
    int64_t* d_bpart = nullptr;
    int64_t* d_btotal = nullptr;
    int64_t* h_btotal = nullptr;
    struct SparseJob {
        WidenPool* pool = nullptr;
        const uint32_t* masks = nullptr;
        const uint32_t* bbase = nullptr;
        const int32_t* packed = nullptr;
        size_t n = 0;
        void* dst = nullptr;
        bool dst64 = true;
        WidenTicket* ticket = nullptr;
    } sparse_job;
};

struct sbwt_gpu_session {
    sbwt_gpu_index* idx = nullptr;
    int64_t max_bases = 0, max_reads = 0;
    int window = 256;
    Scratch sc;          // for device-buffer batches
    bool timing = false; // record events around the walk kernel of device-buffer batches
    cudaEvent_t ev_start = nullptr, ev_walk0 = nullptr, ev_walk1 = nullptr;
    bool host_ready = false;
    static constexpr int kSlots = 3; // H2D of chunk i+1, kernels + D2H of chunk i, host widening of chunk i-1
    HostSlot slots[kSlots];
    WidenPool* widen_pool = nullptr; // created by the first int64 host-buffer call on a narrow index
    int widen_threads = -1;          // -1 = not decided yet, 0 = results travel as int64
    char* h_text[2] = {nullptr, nullptr}; // pinned staging the text leaves the device through, piece by piece
    cudaEvent_t text_ev[2] = {nullptr, nullptr};
};

// ------------------------------------------------------------------ small helpers

static int dmalloc(void** p, size_t bytes, int64_t* account = nullptr) {
    CU(cudaMalloc(p, bytes ? bytes : 16));
    if (account) *account += (int64_t)bytes;
    return 0;
}

static int exclusive_scan_inplace(int64_t* d_data, int64_t n, int64_t* d_partials, int64_t* d_total, cudaStream_t st) {
    if (n == 0) {
        CU(cudaMemsetAsync(d_total, 0, 8, st));
        return 0;
    }
    const unsigned nb = grid_for(n, kScanTile);
    scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(d_data, n, d_partials); LAUNCHED();
    scan_partials_kernel<<<1, kScanThreads, 0, st>>>(d_partials, nb, d_total); LAUNCHED();
    scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(d_data, n, d_partials); LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

static int64_t scan_partials_needed(int64_t n) { return (n + kScanTile - 1) / kScanTile + 1; }

// ------------------------------------------------------------------ API: misc

extern "C" const char* sbwt_gpu_last_error(void) { return g_last_error.c_str(); }
extern "C" int sbwt_gpu_abi_version(void) { return SBWT_B200_ABI_VERSION; }
extern "C" int sbwt_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int64_t sbwt_gpu_launch_count(int reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

static int64_t count_outputs_range(const int64_t* off, int64_t i0, int64_t i1, int64_t k) {
    int64_t t = 0;
    for (int64_t i = i0; i < i1; i++) { // (branch-free: the compiler unrolls and vectorises it)
        const int64_t v = off[i + 1] - off[i] - (k - 1);
        t += v > 0 ? v : 0;
    }
    return t;
}

extern "C" int64_t sbwt_gpu_count_outputs(const int64_t* off, int64_t n_reads, int64_t k) {
    if (n_reads < (1 << 20)) return count_outputs_range(off, 0, n_reads, k);
    // a pass over tens of MB of offsets is bound by one core's memory bandwidth: four threads for large batches
    constexpr int T = 4;
    int64_t part[T] = {0, 0, 0, 0};
    std::thread th[T - 1];
    for (int j = 1; j < T; j++) th[j - 1] = std::thread([&, j] { part[j] = count_outputs_range(off, n_reads * j / T, n_reads * (j + 1) / T, k); });
    part[0] = count_outputs_range(off, 0, n_reads / T, k);
    for (std::thread& x : th) x.join();
    return part[0] + part[1] + part[2] + part[3];
}

// ------------------------------------------------------------------ API: index

extern "C" int sbwt_gpu_index_set_table_length(sbwt_gpu_index* ix, int tp);
constexpr int kMaxTableLength = 16; // the table index is the first word of the k-mer window: 16 characters

// The persisting-L2 limit is per device: it follows the largest request among the indexes alive on the device and is given
// back (with the persisting lines) when the last of them goes.
static std::mutex g_l2_mutex;
static std::vector<sbwt_gpu_index*> g_l2_indexes;
static void l2_set_aside_update(int device) { // (device is current)
    std::lock_guard<std::mutex> lock(g_l2_mutex);
    int64_t want = 0;
    for (const sbwt_gpu_index* x : g_l2_indexes)
        if (x->device == device) want = std::max(want, x->l2_set_aside);
    size_t have = 0;
    cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
    if ((int64_t)have != want) {
        if (want == 0) cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want);
    }
    cudaGetLastError();
}

static int index_create_impl(const uint64_t* const bits[4], const uint64_t* sgs, int64_t n_nodes, int64_t n_kmers,
                             int64_t k, const int64_t C[4], const int64_t* precalc, int64_t p, int device,
                             sbwt_gpu_index** out) {
    *out = nullptr;
    if (sbwt_gpu_device_count() <= 0) return set_error("no CUDA device available: the SBWT GPU query path has no CPU fallback");
    if (device < 0 || device >= sbwt_gpu_device_count()) return set_error("invalid device %d", device);
    if (n_nodes < 1 || k < 1) return set_error("invalid index: n_nodes=%lld k=%lld", (long long)n_nodes, (long long)k);
    if (p < 0 || p > k || p > 14) return set_error("precalc length %lld is not supported (0 <= p <= min(k,14))", (long long)p);
    DeviceGuard guard(device);
    sbwt_gpu_index* ix = new sbwt_gpu_index();
    { std::lock_guard<std::mutex> lock(g_l2_mutex); g_l2_indexes.push_back(ix); }
    auto fail = [&](int rc) { sbwt_gpu_index_destroy(ix); return rc; };
    ix->device = device;
    ix->n_nodes = n_nodes; ix->n_kmers = n_kmers; ix->k = k; ix->precalc_k = p;
    for (int c = 0; c < 4; c++) ix->C[c] = C[c];
    ix->has_sgs = sgs != nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(set_error("cudaGetDeviceProperties failed"));
    ix->sm_count = prop.multiProcessorCount;

    const char* force_wide = getenv("SBWT_B200_FORCE_WIDE"); // tests: exercise the > 2^32-column layout on small indexes
    const int wide = (n_nodes >= (1ll << 32) - 256 || (force_wide && atoi(force_wide) > 0)) ? 1 : 0;
    const int sb_shift = (force_wide && atoi(force_wide) > 0) ? atoi(force_wide) : kDefaultSbShift;
    const int64_t n_blocks = n_nodes / kBlockCols + 1;
    const int64_t n_sb = wide ? ((n_blocks - 1) >> sb_shift) + 1 : 1;
    const int64_t words_per_vec = n_blocks * kPayloadWords; // u32 words, zero padded
    const int64_t host_words64 = (n_nodes + 63) / 64;

    uint32_t* d_raw = nullptr;
    int64_t* d_counts = nullptr;
    int64_t* d_partials = nullptr;
    int64_t* d_total = nullptr;
    int* d_flag = nullptr;
    auto cleanup_tmp = [&]() { cudaFree(d_raw); cudaFree(d_counts); cudaFree(d_partials); cudaFree(d_total); cudaFree(d_flag); };
#define CUI(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            cleanup_tmp();                                                                            \
            return fail(set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__))); \
        }                                                                                             \
    } while (0)
    CUI(cudaMalloc(&d_raw, (size_t)(4 * words_per_vec + 8) * 4));
    CUI(cudaMemset(d_raw, 0, (size_t)(4 * words_per_vec + 8) * 4));
    for (int c = 0; c < 4; c++)
        CUI(cudaMemcpy(d_raw + c * words_per_vec, bits[c], (size_t)host_words64 * 8, cudaMemcpyHostToDevice));
    // (bits at positions >= n_nodes in the last word are zero in every sdsl-written file; memory_management.hpp:360-367)
    CUI(cudaMalloc(&d_counts, (size_t)(4 * n_blocks) * 8));
    CUI(cudaMalloc(&d_partials, (size_t)scan_partials_needed(n_blocks) * 8));
    CUI(cudaMalloc(&d_total, 8 * 4));
    CUI(cudaMalloc(&d_flag, 4));
    CUI(cudaMemset(d_flag, 0, 4));
    k0_block_popcount_kernel<<<grid_for(4 * n_blocks, 256), 256>>>(d_raw, words_per_vec, n_blocks, d_counts); LAUNCHED();
    for (int c = 0; c < 4; c++)
        if (exclusive_scan_inplace(d_counts + c * n_blocks, n_blocks, d_partials, d_total + c, 0)) { cleanup_tmp(); return fail(1); }
    int64_t totals[4];
    CUI(cudaMemcpy(totals, d_total, 32, cudaMemcpyDeviceToHost));
    for (int c = 0; c < 3; c++)
        if (C[c + 1] - C[c] != totals[c]) {
            cleanup_tmp();
            return fail(set_error("Error: Corrupt index file (C array does not match the bit vectors: C[%d+1]-C[%d]=%lld, ones=%lld)",
                                  c, c, (long long)(C[c + 1] - C[c]), (long long)totals[c]));
        }
    if (C[3] + totals[3] > n_nodes || C[0] < 0) { cleanup_tmp(); return fail(set_error("Error: Corrupt index file (C array out of range)")); }

    if (dmalloc(&ix->d_sectors, (size_t)(4 * n_blocks) * sizeof(Sector), &ix->device_bytes)) { cleanup_tmp(); return fail(1); }
    if (wide && dmalloc(&ix->d_sbbase, (size_t)(4 * n_sb) * 8, &ix->device_bytes)) { cleanup_tmp(); return fail(1); }
    k0_emit_kernel<<<grid_for(4 * n_blocks, 256), 256>>>(d_raw, words_per_vec, n_blocks, d_counts, C[0], C[1], C[2], C[3], wide,
                                                         sb_shift, n_sb, (Sector*)ix->d_sectors, (int64_t*)ix->d_sbbase); LAUNCHED();
    CUI(cudaGetLastError());

    // one-hot layouts: narrow indexes in which (nearly) every column has exactly one outgoing edge.
    // SBWT_B200_COMPACT = 0 never, 1 (default) when at most 5 % of its blocks need the classic sectors, 2 always (tests);
    // SBWT_B200_LAYOUT = c96 | c64 picks the csector format (device_index.cuh). Read here, once per index.
    int compact_mode = 1;
    if (const char* e = getenv("SBWT_B200_COMPACT")) compact_mode = atoi(e);
    // csector format: csector64 (one request per rank) unless only the denser csector96 keeps the structure on chip
    // (measured, profiles/r02m_c2m_csector_format.txt: 150 M columns = 75 MB / 50 MB: 11.04 ms vs 9.72 ms per 1.2e9 lookups;
    // 100 M columns = 50 MB / 33 MB: 7.93 ms vs 8.27 ms)
    int layout = kDefaultCompactLayout;
    if (n_nodes / 2 > ((int64_t)56 << 20) && n_nodes / 3 <= ((int64_t)60 << 20)) layout = LAY_C96;
    if (const char* e = getenv("SBWT_B200_LAYOUT")) layout = strcmp(e, "c64") == 0 ? LAY_C64 : (strcmp(e, "c96") == 0 ? LAY_C96 : layout);
    if (const char* e = getenv("SBWT_B200_COMPACT_SEARCH")) ix->compact_in_search = atoi(e) > 0;
    if (compact_mode >= 2) ix->compact_in_search = true;
    const int ccols = layout == LAY_C64 ? kC64Cols : kCBlockCols;
    const int64_t n_cblocks = n_nodes / ccols + 1;
    int built_layout = LAY_CLASSIC;
    if (!wide && compact_mode > 0) {
        const int64_t n_csb = ((n_cblocks - 1) >> kCSbShift) + 1;
        unsigned long long* d_nflag = nullptr;
        CUI(cudaMalloc(&ix->d_compact, (size_t)n_cblocks * sizeof(Sector)));
        if (layout == LAY_C96) CUI(cudaMalloc(&ix->d_cbase, (size_t)n_csb * 16));
        CUI(cudaMalloc(&d_nflag, 8));
        CUI(cudaMemset(d_nflag, 0, 8));
        if (layout == LAY_C64)
            k0_compact64_kernel<<<grid_for(n_cblocks, 256), 256>>>(d_raw, words_per_vec, n_blocks, d_counts, C[0], C[1], C[2], C[3], n_nodes,
                                                                   n_cblocks, (Sector*)ix->d_compact, d_nflag);
        else
            k0_compact_kernel<<<grid_for(n_cblocks, 256), 256>>>(d_raw, words_per_vec, n_blocks, d_counts, C[0], C[1], C[2], C[3], n_nodes,
                                                                 n_cblocks, (Sector*)ix->d_compact, (uint32_t*)ix->d_cbase, d_nflag);
        LAUNCHED();
        unsigned long long n_flagged = 0;
        cudaError_t e1 = cudaMemcpy(&n_flagged, d_nflag, 8, cudaMemcpyDeviceToHost);
        cudaFree(d_nflag);
        CUI(e1);
        ix->flagged_fraction = (double)n_flagged / (double)n_cblocks;
        if (compact_mode == 1 && ix->flagged_fraction > 0.05) {
            cudaFree(ix->d_compact); cudaFree(ix->d_cbase);
            ix->d_compact = ix->d_cbase = nullptr;
        } else {
            built_layout = layout;
            ix->device_bytes += n_cblocks * (int64_t)sizeof(Sector) + (layout == LAY_C96 ? n_csb * 16 : 0);
        }
    }

    int edges_at_starts = 0;
    if (sgs) {
        const int64_t sgs_words = words_per_vec + 8;
        if (dmalloc(&ix->d_sgs, (size_t)sgs_words * 4, &ix->device_bytes)) { cleanup_tmp(); return fail(1); }
        CUI(cudaMemset(ix->d_sgs, 0, (size_t)sgs_words * 4));
        CUI(cudaMemcpy(ix->d_sgs, sgs, (size_t)host_words64 * 8, cudaMemcpyHostToDevice));
        k0_check_edges_kernel<<<grid_for(words_per_vec, 256), 256>>>(d_raw, words_per_vec, (const uint32_t*)ix->d_sgs, words_per_vec, d_flag); LAUNCHED();
        int flag = 0;
        CUI(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost));
        edges_at_starts = flag ? 0 : 1;
    }
    DeviceIndexView& v = ix->view;
    v.sectors = (const Sector*)ix->d_sectors;
    v.sbbase = (const int64_t*)ix->d_sbbase;
    v.sgs = (const uint32_t*)ix->d_sgs;
    v.n_nodes = n_nodes; v.n_blocks = n_blocks; v.n_sb = n_sb;
    v.k = (int)k; v.p = (int)p; v.sb_shift = sb_shift; v.wide = wide; v.edges_at_starts = edges_at_starts;
    v.compact = (const Sector*)ix->d_compact; v.cbase = (const uint32_t*)ix->d_cbase; v.n_cblocks = n_cblocks; v.layout = built_layout;
    if (p > 0) {
        const size_t bytes = (size_t)16 << (2 * p);
        const int64_t np = 1ll << (2 * p);
        if (dmalloc(&ix->d_precalc, bytes + 32, &ix->device_bytes)) { cleanup_tmp(); return fail(1); }
        CUI(cudaMemset((char*)ix->d_precalc + bytes, 0xFF, 32));
        if (precalc) {
            // keep the file's table verbatim, and check it against SBWT::do_kmer_prefix_precalc on the device
            CUI(cudaMemcpy(ix->d_precalc, precalc, bytes, cudaMemcpyHostToDevice));
            int64_t* d_chk = nullptr;
            CUI(cudaMalloc(&d_chk, bytes));
            if (wide) precalc_kernel<true, false><<<grid_for(np, 256), 256>>>(v, (int)p, d_chk);
            else precalc_kernel<false, false><<<grid_for(np, 256), 256>>>(v, (int)p, d_chk);
            LAUNCHED();
            CUI(cudaMemset(d_flag, 0, 4));
            table_compare_kernel<<<grid_for(2 * np, 256), 256>>>((const int64_t*)ix->d_precalc, d_chk, 2 * np, d_flag); LAUNCHED();
            int flag = 0;
            cudaError_t e1 = cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
            cudaFree(d_chk);
            CUI(e1);
            ix->table_from_bits = flag == 0;
        } else { // SBWT::do_kmer_prefix_precalc on the device
            if (wide) precalc_kernel<true, false><<<grid_for(np, 256), 256>>>(v, (int)p, ix->d_precalc);
            else precalc_kernel<false, false><<<grid_for(np, 256), 256>>>(v, (int)p, ix->d_precalc);
            LAUNCHED();
            CUI(cudaGetLastError());
        }
    }
    v.precalc = (const int64_t*)ix->d_precalc;
    CUI(cudaDeviceSynchronize());
    cleanup_tmp();
#undef CUI
    // search table: by default a few characters longer than the file's (same results, fewer dependent steps)
    int tp = (int)p;
    if (ix->table_from_bits) {
        const char* e = getenv("SBWT_B200_TABLE_P");
        if (e) tp = atoi(e);
        else {
            // longest table (<= 16 characters) with at most 64 rows per column of the index that takes at most a fifth of the
            // device memory free right now. A row is read once per from-scratch search and replaces one dependent
            // interval step per extra character -- the steps right after the table jump are the expensive ones (wide
            // intervals, two sectors, lanes dying at different depths) -- and every workload measured faster with each
            // character although the table then dwarfs the rank structure and lives in HBM: c2 10.63 / 9.31 / 7.94 / 7.22 /
            // 6.93 ms for 10 / 12 / 14 / 15 / 16 characters (2.1 / 8.6 / 34 GB of 8-byte rows), c4s 21.8 / 19.7 / 17.6 /
            // 16.4 / 15.7 ms (profiles/r02q_table_length.txt). HBM is what a B200 has plenty of.
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
            const int64_t row_bytes = wide ? 16 : 8;
            tp = (int)std::min<int64_t>(std::max<int64_t>(p, kMaxTableLength), k);
            while (tp > p && ((1ll << (2 * tp)) > std::max<int64_t>(64 * n_nodes, 1 << 16) || (row_bytes << (2 * tp)) > (int64_t)(free_b / 5))) tp--;
        }
    }
    if (int rc = sbwt_gpu_index_set_table_length(ix, tp)) { sbwt_gpu_index_destroy(ix); return rc; }
    // L2 set-aside for a one-hot index that fits on chip: its csectors are read with an evict_last (= persisting) policy, and
    // a set-aside of the structure's size keeps the result / read / table streams from displacing them. Measured on c2
    // (profiles/r02h_l2_set_aside.txt: csector64, 50 MB: 8.62 ms without, 7.94 ms with 52 MB; larger set-asides starve the
    // streams: 60 MB 8.35 ms, and 72 MB 2.2x slower in r02d). Device-wide limit, so only raised, and only for indexes it pays for.
    if (ix->d_compact) {
        const int64_t cbytes = ix->view.n_cblocks * (int64_t)sizeof(Sector);
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        const char* e = getenv("SBWT_B200_L2_SET_ASIDE_MB"); // 0 = never
        int64_t want = e ? (int64_t)atoi(e) << 20 : (cbytes <= ((int64_t)56 << 20) ? cbytes + ((int64_t)2 << 20) : 0);
        ix->l2_set_aside = std::min<int64_t>(want, max_persist);
        l2_set_aside_update(device);
    }
    if (ix->table_from_bits) { // (see launch_walk)
        int log4n = 0;
        while (log4n < 32 && (1ll << (2 * log4n)) < ix->n_nodes) log4n++;
        const int64_t d = ix->k - log4n - 2;
        ix->probe_stride = d >= 8 ? (uint32_t)d : 0u;
        if (const char* pe = getenv("SBWT_B200_PROBE")) ix->probe_stride = (uint32_t)std::max(0, atoi(pe));
    }
    *out = ix;
    return 0;
}

extern "C" int sbwt_gpu_index_set_table_length(sbwt_gpu_index* ix, int tp) {
    if (!ix) return set_error("null index");
    if (tp < 0) return set_error("negative table length");
    if (tp > ix->k) tp = (int)ix->k;
    if (tp > kMaxTableLength) tp = kMaxTableLength;
    if (!ix->table_from_bits && tp != ix->precalc_k)
        return set_error("the file's precalc table does not follow from its bit vectors; only its own length (%lld) can be used", (long long)ix->precalc_k);
    DeviceGuard guard(ix->device);
    CU(cudaDeviceSynchronize());
    if (ix->d_table) { cudaFree(ix->d_table); ix->device_bytes -= ix->table_bytes; ix->d_table = nullptr; ix->table_bytes = 0; }
    DeviceIndexView& v = ix->view;
    v.table = nullptr;
    v.tp = 0;
    if (tp == 0) return 0;
    const int64_t np = 1ll << (2 * tp);
    const bool wide = v.wide;
    ix->table_bytes = np * (wide ? 16 : 8) + 32;
    CU(cudaMalloc(&ix->d_table, ix->table_bytes));
    ix->device_bytes += ix->table_bytes;
    CU(cudaMemset((char*)ix->d_table + ix->table_bytes - 32, 0xFF, 32));
    if (tp == ix->precalc_k) { // the file's own table, re-encoded
        if (wide) CU(cudaMemcpy(ix->d_table, ix->d_precalc, np * 16, cudaMemcpyDeviceToDevice));
        else { table_compact_kernel<<<grid_for(np, 256), 256>>>((const int64_t*)ix->d_precalc, np, (uint2*)ix->d_table); LAUNCHED(); }
    } else {
        // level by level from a short table walked from scratch (table_extend_kernel); the two previous levels live in scratch
        // buffers of 1/4 + 1/16 of the table's size -- without room for them every row is walked from scratch as before
        const int row_bytes = wide ? 16 : 8, p0 = std::min(tp, 8);
        void* tmp[2] = {nullptr, nullptr};
        bool levels = tp >= 15; // (shorter tables are quicker from scratch than the scratch buffers take to allocate: profiles/r03j)
        if (levels) {
            if (cudaMalloc(&tmp[0], (size_t)(np >> 2) * row_bytes) != cudaSuccess) { cudaGetLastError(); levels = false; }
            else if (tp - 2 >= p0 && cudaMalloc(&tmp[1], (size_t)(np >> 4) * row_bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(tmp[0]); tmp[0] = nullptr; levels = false; }
        }
        if (!levels) {
            if (wide) precalc_kernel<true, false><<<grid_for(np, 256), 256>>>(v, tp, ix->d_table);
            else precalc_kernel<false, true><<<grid_for(np, 256), 256>>>(v, tp, ix->d_table);
            LAUNCHED();
        } else {
            // level q lands in: the table itself for q == tp, else the scratch buffer of its parity counted from the top
            auto buf = [&](int q) -> void* { return q == tp ? ix->d_table : tmp[(tp - 1 - q) & 1]; };
            const int64_t n0 = 1ll << (2 * p0);
            if (wide) precalc_kernel<true, false><<<grid_for(n0, 256), 256>>>(v, p0, buf(p0));
            else precalc_kernel<false, true><<<grid_for(n0, 256), 256>>>(v, p0, buf(p0));
            LAUNCHED();
            for (int q = p0 + 1; q <= tp; q++) {
                const int64_t nq = 1ll << (2 * q);
                if (wide) table_extend_kernel<true, false><<<grid_for(nq, 256), 256>>>(v, q, buf(q - 1), buf(q));
                else table_extend_kernel<false, true><<<grid_for(nq, 256), 256>>>(v, q, buf(q - 1), buf(q));
                LAUNCHED();
            }
        }
        const cudaError_t e = cudaDeviceSynchronize();
        cudaFree(tmp[0]);
        cudaFree(tmp[1]);
        CU(e);
    }
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    v.table = ix->d_table;
    v.tp = tp;
    return 0;
}

extern "C" int sbwt_gpu_index_table_length(const sbwt_gpu_index* ix) { return ix ? ix->view.tp : 0; }

extern "C" int sbwt_gpu_index_create(const uint64_t* const bits[4], const uint64_t* sgs, int64_t n_nodes, int64_t n_kmers,
                                     int64_t k, const int64_t C[4], const int64_t* precalc, int64_t p, int device,
                                     sbwt_gpu_index** out) {
    if (!out || !bits || !C) return set_error("null argument");
    return index_create_impl(bits, sgs, n_nodes, n_kmers, k, C, precalc, p, device, out);
}

extern "C" int sbwt_gpu_index_load(const char* path, int device, sbwt_gpu_index** out) {
    if (!out || !path) return set_error("null argument");
    *out = nullptr;
    try {
        PlainMatrixFile F = load_plain_matrix_file(path);
        const uint64_t* bits[4] = {F.bits[0].data(), F.bits[1].data(), F.bits[2].data(), F.bits[3].data()};
        return index_create_impl(bits, F.suffix_group_starts.empty() ? nullptr : F.suffix_group_starts.data(), F.n_nodes,
                                 F.n_kmers, F.k, F.C, F.precalc.empty() ? nullptr : F.precalc.data(), F.precalc_k, device, out);
    } catch (const std::exception& e) {
        return set_error("%s", e.what());
    }
}

extern "C" void sbwt_gpu_index_destroy(sbwt_gpu_index* ix) {
    if (!ix) return;
    DeviceGuard guard(ix->device);
    cudaFree(ix->d_sectors); cudaFree(ix->d_sbbase); cudaFree(ix->d_precalc); cudaFree(ix->d_sgs); cudaFree(ix->d_table);
    cudaFree(ix->d_compact); cudaFree(ix->d_cbase); cudaFree(ix->d_scratch);
    {
        std::lock_guard<std::mutex> lock(g_l2_mutex);
        g_l2_indexes.erase(std::remove(g_l2_indexes.begin(), g_l2_indexes.end(), ix), g_l2_indexes.end());
    }
    if (ix->l2_set_aside) l2_set_aside_update(ix->device);
    delete ix;
}

extern "C" int64_t sbwt_gpu_index_k(const sbwt_gpu_index* ix) { return ix->k; }
extern "C" int64_t sbwt_gpu_index_n_nodes(const sbwt_gpu_index* ix) { return ix->n_nodes; }
extern "C" int64_t sbwt_gpu_index_n_kmers(const sbwt_gpu_index* ix) { return ix->n_kmers; }
extern "C" int64_t sbwt_gpu_index_precalc_k(const sbwt_gpu_index* ix) { return ix->precalc_k; }
extern "C" int sbwt_gpu_index_has_streaming_support(const sbwt_gpu_index* ix) { return ix->has_sgs ? 1 : 0; }
extern "C" int sbwt_gpu_index_device(const sbwt_gpu_index* ix) { return ix->device; }
extern "C" void sbwt_gpu_index_C(const sbwt_gpu_index* ix, int64_t C[4]) { for (int c = 0; c < 4; c++) C[c] = ix->C[c]; }
extern "C" int64_t sbwt_gpu_index_device_bytes(const sbwt_gpu_index* ix) { return ix->device_bytes; }
extern "C" int sbwt_gpu_index_compact_layout(const sbwt_gpu_index* ix, double* flagged_fraction) {
    if (flagged_fraction) *flagged_fraction = ix->flagged_fraction;
    return ix->d_compact ? ix->view.layout : 0; // LAY_C96 = 1, LAY_C64 = 2
}
extern "C" int64_t sbwt_gpu_index_l2_set_aside(const sbwt_gpu_index* ix) { return ix ? ix->l2_set_aside : 0; }
extern "C" int sbwt_gpu_index_edges_only_at_group_starts(const sbwt_gpu_index* ix) { return ix->view.edges_at_starts; }

extern "C" int sbwt_gpu_index_get_precalc(const sbwt_gpu_index* ix, int64_t* out_lr) {
    if (!ix || !out_lr) return set_error("null argument");
    if (ix->precalc_k == 0) return 0;
    DeviceGuard guard(ix->device);
    CU(cudaMemcpy(out_lr, ix->d_precalc, (size_t)16 << (2 * ix->precalc_k), cudaMemcpyDeviceToHost));
    return 0;
}

// scratch of at least `bytes` (256-byte aligned pieces are carved out of it by the caller); call with scratch_mutex held
static int index_scratch(sbwt_gpu_index* ix, size_t bytes, char** out) {
    if (bytes > ix->scratch_bytes) {
        cudaFree(ix->d_scratch);
        ix->d_scratch = nullptr;
        ix->scratch_bytes = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 2, (size_t)1 << 20);
        CU(cudaMalloc(&ix->d_scratch, want));
        ix->scratch_bytes = want;
    }
    *out = (char*)ix->d_scratch;
    return 0;
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int sbwt_gpu_rank(sbwt_gpu_index* ix, const int64_t* pos, const char* chars, int64_t n, int64_t* out) {
    if (!ix) return set_error("null index");
    if (n <= 0) return 0;
    if (!pos || !chars || !out) return set_error("null buffer");
    for (int64_t i = 0; i < n; i++)
        if (pos[i] < 0 || pos[i] > ix->n_nodes) return set_error("rank position %lld out of range [0, %lld]", (long long)pos[i], (long long)ix->n_nodes);
    DeviceGuard guard(ix->device);
    std::lock_guard<std::mutex> lock(ix->scratch_mutex);
    char* d = nullptr;
    if (index_scratch(ix, 2 * al256((size_t)n * 8) + al256((size_t)n), &d)) return 1;
    int64_t* d_pos = (int64_t*)d;
    int64_t* d_out = (int64_t*)(d + al256((size_t)n * 8));
    char* d_ch = d + 2 * al256((size_t)n * 8);
    CU(cudaMemcpy(d_pos, pos, n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_ch, chars, n, cudaMemcpyHostToDevice));
    if (ix->view.wide) rank_kernel<true><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, ix->C[0], ix->C[1], ix->C[2], ix->C[3], d_out);
    else rank_kernel<false><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, ix->C[0], ix->C[1], ix->C[2], ix->C[3], d_out);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d_out, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------ the other read-only queries, batched (query_kernels.cuh)

static int extend_batch(sbwt_gpu_index* ix, const char* ascii, const int64_t* off, int64_t n, int mode, int64_t* l, int64_t* r, int64_t* matched) {
    if (!ix) return set_error("null index");
    if (n < 0) return set_error("negative batch size");
    if (n == 0) return 0;
    if (!off || !l || !r || (mode == kExtendPartial && !matched)) return set_error("null buffer");
    const int64_t bytes = off[n] - off[0];
    if (bytes < 0 || (bytes > 0 && !ascii)) return set_error("invalid string offsets");
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return set_error("string offsets must be non-decreasing");
    if (mode == kExtendUpdate)
        for (int64_t i = 0; i < n; i++)
            if (l[i] != -1 && (l[i] < 0 || r[i] < l[i] - 1 || r[i] >= ix->n_nodes))
                return set_error("interval %lld: [%lld, %lld] is not inside [0, %lld)", (long long)i, (long long)l[i], (long long)r[i], (long long)ix->n_nodes);
    DeviceGuard guard(ix->device);
    std::lock_guard<std::mutex> lock(ix->scratch_mutex);
    char* d = nullptr;
    const size_t s_off = al256((size_t)(n + 1) * 8), s_v = al256((size_t)n * 8), s_a = al256((size_t)bytes + 1);
    if (index_scratch(ix, s_off + 3 * s_v + s_a, &d)) return 1;
    int64_t* d_off = (int64_t*)d;
    int64_t *d_l = (int64_t*)(d + s_off), *d_r = (int64_t*)(d + s_off + s_v), *d_m = (int64_t*)(d + s_off + 2 * s_v);
    uint8_t* d_a = (uint8_t*)(d + s_off + 3 * s_v);
    CU(cudaMemcpy(d_off, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice));
    if (bytes) CU(cudaMemcpy(d_a, ascii + off[0], (size_t)bytes, cudaMemcpyHostToDevice));
    if (mode == kExtendUpdate) {
        CU(cudaMemcpy(d_l, l, (size_t)n * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_r, r, (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    if (ix->view.wide) extend_kernel<true><<<grid_for(n, 256), 256>>>(ix->view, d_a, d_off, n, mode, d_l, d_r, d_m);
    else extend_kernel<false><<<grid_for(n, 256), 256>>>(ix->view, d_a, d_off, n, mode, d_l, d_r, d_m);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpy(l, d_l, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r, d_r, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (matched) CU(cudaMemcpy(matched, d_m, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sbwt_gpu_update_interval_batch(sbwt_gpu_index* ix, const char* ascii, const int64_t* offsets, int64_t n, int64_t* l, int64_t* r) {
    return extend_batch(ix, ascii, offsets, n, kExtendUpdate, l, r, nullptr);
}
extern "C" int sbwt_gpu_partial_search_batch(sbwt_gpu_index* ix, const char* ascii, const int64_t* offsets, int64_t n, int64_t* l, int64_t* r,
                                             int64_t* matched) {
    return extend_batch(ix, ascii, offsets, n, kExtendPartial, l, r, matched);
}

// positions + characters in, one value each out
template <typename OutT, typename Launch>
static int pos_char_batch(sbwt_gpu_index* ix, const int64_t* pos, const char* chars, int64_t n, int64_t max_pos, OutT* out, Launch launch) {
    if (!ix) return set_error("null index");
    if (n < 0) return set_error("negative batch size");
    if (n == 0) return 0;
    if (!pos || !chars || !out) return set_error("null buffer");
    for (int64_t i = 0; i < n; i++)
        if (pos[i] < 0 || pos[i] > max_pos) return set_error("position %lld out of range [0, %lld]", (long long)pos[i], (long long)max_pos);
    DeviceGuard guard(ix->device);
    std::lock_guard<std::mutex> lock(ix->scratch_mutex);
    char* d = nullptr;
    if (index_scratch(ix, 2 * al256((size_t)n * 8) + al256((size_t)n), &d)) return 1;
    int64_t* d_pos = (int64_t*)d;
    OutT* d_out = (OutT*)(d + al256((size_t)n * 8));
    char* d_ch = d + 2 * al256((size_t)n * 8);
    CU(cudaMemcpy(d_pos, pos, (size_t)n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_ch, chars, (size_t)n, cudaMemcpyHostToDevice));
    launch(d_pos, d_ch, d_out);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d_out, (size_t)n * sizeof(OutT), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sbwt_gpu_forward_batch(sbwt_gpu_index* ix, const int64_t* nodes, const char* chars, int64_t n, int64_t* out) {
    if (ix && !ix->has_sgs) return set_error("Error: Streaming support required for SBWT::forward"); // SBWT.hh:370-371
    return pos_char_batch<int64_t>(ix, nodes, chars, n, ix ? ix->n_nodes - 1 : 0, out, [&](const int64_t* d_pos, const char* d_ch, int64_t* d_out) {
        if (ix->view.wide) forward_kernel<true><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, d_out);
        else forward_kernel<false><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, d_out);
    });
}

extern "C" int sbwt_gpu_contains_batch(sbwt_gpu_index* ix, const int64_t* pos, const char* chars, int64_t n, uint8_t* out) {
    return pos_char_batch<uint8_t>(ix, pos, chars, n, ix ? ix->n_nodes - 1 : 0, out, [&](const int64_t* d_pos, const char* d_ch, uint8_t* d_out) {
        if (ix->view.wide) contains_kernel<true><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, d_out);
        else contains_kernel<false><<<grid_for(n, 256), 256>>>(ix->view, d_pos, d_ch, n, d_out);
    });
}

extern "C" int sbwt_gpu_get_kmer_batch(sbwt_gpu_index* ix, const int64_t* colex_ranks, int64_t n, char* out) {
    if (!ix) return set_error("null index");
    if (n < 0) return set_error("negative batch size");
    if (n == 0) return 0;
    if (!colex_ranks || !out) return set_error("null buffer");
    for (int64_t i = 0; i < n; i++)
        if (colex_ranks[i] < 0 || colex_ranks[i] >= ix->n_nodes) return set_error("colex rank %lld out of range [0, %lld)", (long long)colex_ranks[i], (long long)ix->n_nodes);
    DeviceGuard guard(ix->device);
    std::lock_guard<std::mutex> lock(ix->scratch_mutex);
    char* d = nullptr;
    if (index_scratch(ix, al256((size_t)n * 8) + al256((size_t)n * (size_t)ix->k), &d)) return 1;
    int64_t* d_rk = (int64_t*)d;
    char* d_out = d + al256((size_t)n * 8);
    CU(cudaMemcpy(d_rk, colex_ranks, (size_t)n * 8, cudaMemcpyHostToDevice));
    if (ix->view.wide) get_kmer_kernel<true><<<grid_for(n, 256), 256>>>(ix->view, d_rk, n, ix->C[0], ix->C[1], ix->C[2], ix->C[3], d_out);
    else get_kmer_kernel<false><<<grid_for(n, 256), 256>>>(ix->view, d_rk, n, ix->C[0], ix->C[1], ix->C[2], ix->C[3], d_out);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d_out, (size_t)n * (size_t)ix->k, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sbwt_gpu_ascii_export_sets(sbwt_gpu_index* ix, char* out, int64_t capacity, int64_t* n_bytes) {
    if (!ix || !n_bytes) return set_error("null argument");
    *n_bytes = 0;
    DeviceGuard guard(ix->device);
    std::lock_guard<std::mutex> lock(ix->scratch_mutex);
    const int64_t n_words = (ix->n_nodes + 31) / 32;
    char* d = nullptr;
    const size_t s_len = al256((size_t)(n_words + 1) * 8), s_part = al256((size_t)scan_partials_needed(n_words) * 8);
    if (index_scratch(ix, s_len + s_part + 256, &d)) return 1;
    int64_t *d_len = (int64_t*)d, *d_part = (int64_t*)(d + s_len), *d_total = (int64_t*)(d + s_len + s_part);
    export_len_kernel<<<grid_for(n_words, 256), 256>>>(ix->view, n_words, d_len); LAUNCHED();
    if (exclusive_scan_inplace(d_len, n_words, d_part, d_total, 0)) return 1;
    int64_t total = 0;
    CU(cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost));
    *n_bytes = total + 1; // the text ends with one newline (SBWT.hh:772)
    if (!out || capacity < total + 1) return set_error("ascii_export_sets needs a buffer of %lld bytes", (long long)(total + 1));
    char* d_text = nullptr;
    CU(cudaMalloc(&d_text, (size_t)total + 1));
    export_emit_kernel<<<grid_for(n_words, 256), 256>>>(ix->view, n_words, d_len, d_text); LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_text, (size_t)total, cudaMemcpyDeviceToHost);
    cudaFree(d_text);
    CU(e);
    out[total] = '\n';
    return 0;
}

// ------------------------------------------------------------------ sessions

static void session_count(int delta); // sessions alive in this process (see widen_thread_count)

static void scratch_free(Scratch& sc) {
    cudaFree(sc.codes); cudaFree(sc.invalid); cudaFree(sc.lower); cudaFree(sc.n_out); cudaFree(sc.n_win); cudaFree(sc.partials);
    cudaFree(sc.totals); cudaFree(sc.items); cudaFree(sc.stats);
    sc = Scratch();
}

static int scratch_alloc(Scratch& sc, int64_t max_bases, int64_t max_reads, int window) {
    sc.max_bases = max_bases; sc.max_reads = max_reads;
    sc.max_items = max_reads + max_bases / std::min(window, 32) + 1; // search mode plans 32-k-mer chunks
    const int64_t words = (max_bases / 32 + 8 + 1) & ~1ll; // even: the walk kernel reads whole 64-base chunks
    sc.n_words = words;
    CU(cudaMalloc(&sc.codes, words * 8));
    CU(cudaMalloc(&sc.invalid, words * 4));
    CU(cudaMemset(sc.codes, 0, words * 8));
    CU(cudaMemset(sc.invalid, 0xFF, words * 4));
    CU(cudaMalloc(&sc.n_out, (max_reads + 1) * 8));
    CU(cudaMalloc(&sc.n_win, (max_reads + 1) * 8));
    CU(cudaMalloc(&sc.partials, 2 * scan_partials_needed(max_reads + 1) * 8)); // (two channels: plan_reduce_kernel)
    CU(cudaMalloc(&sc.totals, 4 * 8));
    CU(cudaMalloc(&sc.items, sc.max_items * sizeof(WalkItem)));
    CU(cudaMalloc(&sc.stats, 8 * 8));
    return 0;
}

extern "C" int sbwt_gpu_session_create(sbwt_gpu_index* ix, int64_t max_bases, int64_t max_reads, sbwt_gpu_session** out) {
    if (!ix || !out) return set_error("null argument");
    *out = nullptr;
    if (max_bases < 1 || max_reads < 1) return set_error("session capacity must be positive");
    if (max_bases > (int64_t)0xFFFF0000ll || max_reads > (int64_t)0x7FFF0000ll)
        return set_error("session capacity too large: at most 2^32 - 65536 bases and 2^31 - 65536 reads per device-side batch (host-side calls are chunked)");
    DeviceGuard guard(ix->device);
    sbwt_gpu_session* s = new sbwt_gpu_session();
    s->idx = ix; s->max_bases = max_bases; s->max_reads = max_reads;
    if (const char* w = getenv("SBWT_B200_WINDOW")) s->window = std::max(1, atoi(w));
    session_count(+1);
    if (scratch_alloc(s->sc, max_bases, max_reads, s->window)) { sbwt_gpu_session_destroy(s); return 1; }
    *out = s;
    return 0;
}

extern "C" int sbwt_gpu_session_set_timing(sbwt_gpu_session* s, int enable) {
    if (!s) return set_error("null session");
    DeviceGuard guard(s->idx->device);
    if (enable && !s->ev_start) {
        CU(cudaEventCreate(&s->ev_start)); CU(cudaEventCreate(&s->ev_walk0)); CU(cudaEventCreate(&s->ev_walk1));
    }
    s->timing = enable != 0;
    return 0;
}

extern "C" int sbwt_gpu_session_last_timing(sbwt_gpu_session* s, double* prep_ms, double* walk_ms) {
    if (!s || !s->ev_start) return set_error("timing was not enabled on this session");
    DeviceGuard guard(s->idx->device);
    CU(cudaEventSynchronize(s->ev_walk1));
    float a = 0, b = 0;
    CU(cudaEventElapsedTime(&a, s->ev_start, s->ev_walk0));
    CU(cudaEventElapsedTime(&b, s->ev_walk0, s->ev_walk1));
    if (prep_ms) *prep_ms = a;
    if (walk_ms) *walk_ms = b;
    return 0;
}

extern "C" void sbwt_gpu_session_destroy(sbwt_gpu_session* s) {
    if (!s) return;
    session_count(-1);
    DeviceGuard guard(s->idx->device);
    scratch_free(s->sc);
    if (s->ev_start) { cudaEventDestroy(s->ev_start); cudaEventDestroy(s->ev_walk0); cudaEventDestroy(s->ev_walk1); }
    for (HostSlot& h : s->slots) {
        scratch_free(h.sc);
        cudaFree(h.d_ascii); cudaFree(h.d_offsets); cudaFree(h.d_out); cudaFree(h.d_text);
        cudaFreeHost(h.h_ascii); cudaFreeHost(h.h_offsets); cudaFreeHost(h.h_out); cudaFreeHost(h.h_totals); cudaFreeHost(h.h_out32);
        cudaFree(h.d_masks); cudaFree(h.d_bbase); cudaFree(h.d_total); cudaFree(h.d_bcount); cudaFree(h.d_bpart); cudaFree(h.d_btotal);
        cudaFreeHost(h.h_btotal);
        cudaFreeHost(h.h_masks); cudaFreeHost(h.h_bbase); cudaFreeHost(h.h_total);
        if (h.ev_small) cudaEventDestroy(h.ev_small);
        if (h.stream) cudaStreamDestroy(h.stream);
        if (h.done) cudaEventDestroy(h.done);
    }
    for (int b = 0; b < 2; b++) {
        cudaFreeHost(s->h_text[b]);
        if (s->text_ev[b]) cudaEventDestroy(s->text_ev[b]);
    }
    delete s->widen_pool;
    delete s;
}

// ------------------------------------------------------------------ launches

static int launch_pack(const char* d_ascii, int64_t n_bases, int case_mode, uint64_t* codes, uint32_t* invalid, cudaStream_t st) {
    const int64_t n_units = 2 * (n_bases / 32 + 4); // 16-base units, padding words included
    const uint32_t fold = case_mode == SBWT_GPU_CASE_EXACT ? 0xFFFFFFFFu : 0xDFDFDFDFu; // (CASE_API arrives here as EXACT or UPPER)
    const bool vec = ((uintptr_t)d_ascii & 15) == 0;
    const unsigned grid = grid_for(n_units, 256 * kPackUnits);
    if (vec) pack_kernel<true><<<grid, 256, 0, st>>>((const uint8_t*)d_ascii, n_bases, fold, reinterpret_cast<uint32_t*>(codes), invalid, n_units);
    else pack_kernel<false><<<grid, 256, 0, st>>>((const uint8_t*)d_ascii, n_bases, fold, reinterpret_cast<uint32_t*>(codes), invalid, n_units);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32, int KW, bool LITERAL, int LAY>
static cudaError_t launch_walk_ttt(const WalkParams& P, int sm_count, cudaStream_t st) {
    constexpr size_t smem = sizeof(WalkShared<STREAMING, WIDE, OUT32, LITERAL>); // queues + stage: dynamic (more than 48 KB)
    static int occ[64] = {0}; // resident blocks per SM of this instantiation, per device (the attribute below is per device too)
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (occ[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(walk_kernel<STREAMING, WIDE, COUNT, OUT32, KW, LITERAL, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, walk_kernel<STREAMING, WIDE, COUNT, OUT32, KW, LITERAL, LAY>, kWalkThreads, smem);
        if (e != cudaSuccess) return e;
        occ[dev] = o < 1 ? 1 : o;
    }
    walk_kernel<STREAMING, WIDE, COUNT, OUT32, KW, LITERAL, LAY><<<(unsigned)(sm_count * occ[dev]), kWalkThreads, smem, st>>>(P);
    return cudaGetLastError();
}

template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32, int KW>
static cudaError_t launch_walk_tt(const WalkParams& P, int sm_count, cudaStream_t st) {
    if (STREAMING && !P.ix.edges_at_starts) // the reference's control flow to the letter (hand-made index files)
        return launch_walk_ttt<STREAMING, WIDE, COUNT, OUT32, KW, STREAMING, LAY_CLASSIC>(P, sm_count, st);
    if (!WIDE && P.ix.compact) { // one-hot layouts (device_index.cuh)
        if (P.ix.layout == LAY_C64) return launch_walk_ttt<STREAMING, WIDE, COUNT, OUT32, KW, false, WIDE ? LAY_CLASSIC : LAY_C64>(P, sm_count, st);
        return launch_walk_ttt<STREAMING, WIDE, COUNT, OUT32, KW, false, WIDE ? LAY_CLASSIC : LAY_C96>(P, sm_count, st);
    }
    return launch_walk_ttt<STREAMING, WIDE, COUNT, OUT32, KW, false, LAY_CLASSIC>(P, sm_count, st);
}

template <bool STREAMING, bool WIDE, int KW>
static cudaError_t launch_walk_t(const WalkParams& P, bool count, int sm_count, cudaStream_t st) {
    if (P.out32) {
        if (WIDE || count) return cudaErrorInvalidValue; // int32 results exist only for narrow, uncounted batches
        return launch_walk_tt<STREAMING, false, false, true, KW>(P, sm_count, st);
    }
    return count ? launch_walk_tt<STREAMING, WIDE, true, false, KW>(P, sm_count, st)
                 : launch_walk_tt<STREAMING, WIDE, false, false, KW>(P, sm_count, st);
}

// One persistent wave: grid = SM count x resident blocks per SM (occupancy query); work is handed out through a global cursor.
static int launch_walk(const sbwt_gpu_index* ix, WalkParams& P, bool streaming, bool count, cudaStream_t st) {
    // The compact layouts pay in streaming mode, whose 8 B-per-k-mer result stream competes with the index for L2.
    // The per-k-mer search path re-reads the wide top of the tree, is bound by L2 sector throughput and issue slots,
    // and is faster on the 224-column classic sectors (fewer two-sector steps): profiles/r01h_compact_ab.txt.
    if (!streaming && !ix->compact_in_search) { P.ix.compact = nullptr; P.ix.layout = LAY_CLASSIC; }
    P.table_streams = ix->table_bytes > ((int64_t)16 << 20) ? 1 : 0;
    // probe stride of the streaming walk: a from-scratch walk on this index dies after about log4(n) characters, and a
    // probe that dies at character j proves k - j k-mers absent. Any stride >= 1 gives the same results; it needs the
    // search table to follow from the bit vectors (monotonicity of the interval step), else every k-mer is searched.
    P.probe_stride = streaming ? ix->probe_stride : 0u;
    const bool wide = ix->view.wide;
    cudaError_t e;
    if (ix->k > 64) { // no register-held k-mer window: the literal kernels of long_kmer_kernels.cuh
        const bool o32 = P.out32 != nullptr;
        const unsigned grid = (unsigned)ix->sm_count * 8u;
        const uint32_t k = (uint32_t)ix->k;
#define LONGK(W_, O_) do { if (streaming) long_streaming_kernel<W_, O_><<<grid * 2u, 128, 0, st>>>(P, k); else long_search_kernel<W_, O_><<<grid, 256, 0, st>>>(P, k); } while (0)
        if (wide) { if (o32) LONGK(true, true); else LONGK(true, false); }
        else { if (o32) LONGK(false, true); else LONGK(false, false); }
#undef LONGK
        LAUNCHED();
        CU(cudaGetLastError());
        return 0;
    }
    CU(cudaMemsetAsync(P.cursor, 0, 8, st));
    const bool k64 = ix->k > 32;
#define WALK(S_, W_) e = k64 ? launch_walk_t<S_, W_, 2>(P, count, ix->sm_count, st) : launch_walk_t<S_, W_, 1>(P, count, ix->sm_count, st)
    if (streaming) { if (wide) WALK(true, true); else WALK(true, false); }
    else { if (wide) WALK(false, true); else WALK(false, false); }
#undef WALK
    LAUNCHED();
    CU(e);
    return 0;
}

// pack -> plan -> walk for one device-resident batch, all on `st`
static int run_device_batch(sbwt_gpu_session* s, Scratch& sc, const char* d_ascii, const int64_t* d_offsets, int64_t n_reads,
                            int64_t n_bases, int mode, int case_mode, void* d_out, bool out32, bool count, cudaStream_t st) {
    sbwt_gpu_index* ix = s->idx;
    if (mode != SBWT_GPU_MODE_SEARCH && mode != SBWT_GPU_MODE_STREAMING) return set_error("unknown mode %d", mode);
    if (case_mode != SBWT_GPU_CASE_UPPER && case_mode != SBWT_GPU_CASE_EXACT && case_mode != SBWT_GPU_CASE_API) return set_error("unknown case mode %d", case_mode);
    if (mode == SBWT_GPU_MODE_STREAMING && !ix->has_sgs) return set_error("Error: streaming search support not built"); // SBWT.hh:546-547
    if (n_reads < 0 || n_bases < 0) return set_error("negative batch size");
    if (out32 && (ix->n_nodes >= (1ll << 31) || ix->view.wide))
        return set_error("int32 results need an index with fewer than 2^31 columns (this one has %lld)", (long long)ix->n_nodes);
    if (n_bases > sc.max_bases || n_reads > sc.max_reads)
        return set_error("batch of %lld reads / %lld bases exceeds the session capacity (%lld / %lld)", (long long)n_reads,
                         (long long)n_bases, (long long)sc.max_reads, (long long)sc.max_bases);
    if (n_reads == 0) return 0;
    if (s->timing && &sc == &s->sc) CU(cudaEventRecord(s->ev_start, st));
    // CASE_API: per-k-mer search is CASE_EXACT; streaming is walked upper-cased and corrected afterwards (case_fixup_kernel)
    // (an index that violates the edge invariant -- hand-made files only -- is walked with CASE_EXACT instead: there a streaming
    // step and a from-scratch search may disagree, which the correction pass relies on not happening)
    const bool api_fixup = case_mode == SBWT_GPU_CASE_API && mode == SBWT_GPU_MODE_STREAMING && ix->view.edges_at_starts && ix->table_from_bits;
    const int pack_case = case_mode == SBWT_GPU_CASE_API ? (api_fixup ? SBWT_GPU_CASE_UPPER : SBWT_GPU_CASE_EXACT) : case_mode;
    if (launch_pack(d_ascii, n_bases, pack_case, sc.codes, sc.invalid, st)) return 1;
    // search mode is planned as chunks of 32 k-mers (one lane per k-mer), streaming mode as windows of a read. The first
    // k-mer of a window is searched from scratch, which gives streaming_search's answers only where those equal search()'s:
    // an index that violates the edge invariant (LITERAL kernel) or whose table does not follow from its bit vectors
    // is walked read by read, as SBWT.hh:556-576 does
    const bool windows_ok = ix->view.edges_at_starts && ix->table_from_bits;
    const int window = mode == SBWT_GPU_MODE_SEARCH ? 32 : (windows_ok ? std::min(s->window, 1 << 23) : (1 << 23));
    // (sc.n_out becomes the reads' result offsets, sc.totals[0..1] the number of results and of work items)
    const unsigned n_tiles = grid_for(n_reads, kScanTile);
    int64_t* part_win = sc.partials + scan_partials_needed(sc.max_reads + 1);
    plan_reduce_kernel<<<n_tiles, kScanThreads, 0, st>>>(d_offsets, n_reads, (int)ix->k, window, sc.partials, part_win); LAUNCHED();
    plan_partials_kernel<<<1, kScanThreads, 0, st>>>(sc.partials, part_win, n_tiles, sc.totals); LAUNCHED();
    plan_fused_emit_kernel<<<n_tiles, kScanThreads, 0, st>>>(d_offsets, n_reads, (int)ix->k, window, sc.partials, part_win, sc.invalid, sc.n_out, sc.items);
    LAUNCHED();
    CU(cudaGetLastError());
    WalkParams P;
    P.ix = ix->view;
    P.codes = reinterpret_cast<const uint32_t*>(sc.codes); P.invalid = sc.invalid;
    P.items = sc.items;
    P.out32 = out32 ? (int32_t*)d_out : nullptr;
    P.n_items = sc.totals + 1;
    P.out = out32 ? nullptr : (int64_t*)d_out;
    P.stats = sc.stats;
    P.cursor = sc.stats + 4;
    if (count) CU(cudaMemsetAsync(sc.stats, 0, 64, st));
    const bool timed = s->timing && &sc == &s->sc;
    if (timed) CU(cudaEventRecord(s->ev_walk0, st));
    if (launch_walk(ix, P, mode == SBWT_GPU_MODE_STREAMING, count, st)) return 1;
    if (timed) CU(cudaEventRecord(s->ev_walk1, st));
    if (api_fixup) {
        const int64_t n_words = n_bases / 32 + 1;
        if (!sc.lower) CU(cudaMalloc(&sc.lower, (size_t)(sc.max_bases / 32 + 2) * 4));
        lower_mask_kernel<<<grid_for(n_words, 256), 256, 0, st>>>((const uint8_t*)d_ascii, n_bases, sc.lower, n_words); LAUNCHED();
        if (out32) case_fixup_kernel<int32_t><<<grid_for(n_reads, 256), 256, 0, st>>>(d_offsets, n_reads, (int)ix->k, sc.n_out, sc.lower, (int32_t*)d_out);
        else case_fixup_kernel<int64_t><<<grid_for(n_reads, 256), 256, 0, st>>>(d_offsets, n_reads, (int)ix->k, sc.n_out, sc.lower, (int64_t*)d_out);
        LAUNCHED();
        CU(cudaGetLastError());
    }
    return 0;
}

extern "C" int sbwt_gpu_pack_device(const char* d_ascii, int64_t n_bases, int case_mode, uint64_t* d_codes, uint32_t* d_invalid, void* stream) {
    if (n_bases < 0) return set_error("negative size");
    return launch_pack(d_ascii, n_bases, case_mode, d_codes, d_invalid, (cudaStream_t)stream);
}

extern "C" int sbwt_gpu_query_device(sbwt_gpu_session* s, const char* d_ascii, const int64_t* d_offsets, int64_t n_reads,
                                     int64_t n_bases, int mode, int case_mode, int64_t* d_out, int64_t n_out, void* stream) {
    if (!s) return set_error("null session");
    (void)n_out;
    DeviceGuard guard(s->idx->device);
    return run_device_batch(s, s->sc, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, false, false, (cudaStream_t)stream);
}

extern "C" int sbwt_gpu_query_device_i32(sbwt_gpu_session* s, const char* d_ascii, const int64_t* d_offsets, int64_t n_reads,
                                         int64_t n_bases, int mode, int case_mode, int32_t* d_out, int64_t n_out, void* stream) {
    if (!s) return set_error("null session");
    (void)n_out;
    DeviceGuard guard(s->idx->device);
    return run_device_batch(s, s->sc, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, true, false, (cudaStream_t)stream);
}

extern "C" int sbwt_gpu_query_device_counted(sbwt_gpu_session* s, const char* d_ascii, const int64_t* d_offsets, int64_t n_reads,
                                             int64_t n_bases, int mode, int case_mode, int64_t* d_out, int64_t n_out, void* stream,
                                             sbwt_gpu_stats* stats) {
    if (!s || !stats) return set_error("null argument");
    (void)n_out;
    DeviceGuard guard(s->idx->device);
    const int64_t before = g_launches;
    if (run_device_batch(s, s->sc, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, false, true, (cudaStream_t)stream)) return 1;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned long long h[4] = {0, 0, 0, 0};
    if (n_reads > 0) CU(cudaMemcpy(h, s->sc.stats, 32, cudaMemcpyDeviceToHost));
    stats->lookups = (int64_t)h[0]; stats->hits = (int64_t)h[1]; stats->rank_ops = (int64_t)h[2]; stats->index_sectors = (int64_t)h[3];
    stats->kernel_launches = g_launches - before;
    return 0;
}

// ------------------------------------------------------------------ host pipeline

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int host_slots_init(sbwt_gpu_session* s) {
    if (s->host_ready) return 0;
    for (HostSlot& h : s->slots) {
        if (scratch_alloc(h.sc, s->max_bases, s->max_reads, s->window)) return 1;
        CU(cudaMalloc(&h.d_ascii, s->max_bases + 64));
        CU(cudaMalloc(&h.d_offsets, (s->max_reads + 1) * 8));
        CU(cudaMalloc(&h.d_out, std::max<int64_t>(s->max_bases, 1) * 8));
        CU(cudaMallocHost(&h.h_totals, 4 * 8));
        CU(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
    }
    s->host_ready = true;
    return 0;
}

static void CUDART_CB sparse_callback(void* p) { // stream callback: no CUDA calls in here
    HostSlot::SparseJob* j = static_cast<HostSlot::SparseJob*>(p);
    j->pool->submit_sparse(j->masks, j->bbase, j->packed, j->n, j->dst, j->dst64, j->ticket);
}

// second half of a sparse-format chunk: the hit count has arrived, fetch exactly that many hits and hand the three
// pieces to the pool. Called one chunk late, so that the device already has the next chunk's work queued.
static int slot_back(HostSlot& h) {
    if (!h.back_pending) return 0;
    h.back_pending = false;
    CU(cudaEventSynchronize(h.ev_small));
    const unsigned long long total = *h.h_total;
    if (total > h.sparse_job.n) return set_error("sparse result format: %llu hits reported for %zu results", total, h.sparse_job.n);
    const int32_t* d_packed = reinterpret_cast<const int32_t*>(h.d_out) + h.sc.max_bases;
    if (total) CU(cudaMemcpyAsync(h.h_out32, d_packed, (size_t)total * 4, cudaMemcpyDeviceToHost, h.stream));
    h.widen_pool = h.sparse_job.pool;
    CU(cudaLaunchHostFunc(h.stream, sparse_callback, &h.sparse_job));
    CU(cudaEventRecord(h.done, h.stream));
    return 0;
}

static int slot_finish(HostSlot& h) {
    if (!h.busy) return 0;
    if (slot_back(h)) return 1;
    CU(cudaEventSynchronize(h.done)); // the widening job (if any) was submitted by a callback that precedes this event
    if (h.widen_pool) {
        h.widen_pool->wait(&h.widen_ticket);
        h.widen_pool = nullptr;
    }
    if (h.out_staged && h.out_bytes) memcpy(h.out_dst, h.h_out, (size_t)h.out_bytes);
    h.busy = false;
    return 0;
}

static void CUDART_CB widen_callback(void* p) { // stream callback: no CUDA calls in here
    HostSlot::WidenJob* j = static_cast<HostSlot::WidenJob*>(p);
    j->pool->submit(j->src, j->dst, j->n, j->ticket);
}

// How many host threads sign-extend int32 results into the caller's int64 array (0 = none: int64 values cross
// PCIe). Default: the host's hardware threads divided by the GPUs THIS JOB drives -- LOCAL_WORLD_SIZE when a launcher
// set it (one process per GPU), else the sessions alive in this process (one per device in sbwt_gpu_query_host_sharded)
// -- at most 10; fewer than 4 cannot keep up with a PCIe 5 x16 link, so the plain int64 copy is used instead.
// (Round 1 divided by the VISIBLE devices: a single-GPU job on an 8-GPU host got 4 threads and lost 20 %.)
static std::atomic<int> g_live_sessions{0};
static void session_count(int delta) { g_live_sessions.fetch_add(delta); }
static int widen_thread_count() {
    if (const char* e = getenv("SBWT_B200_WIDEN_THREADS")) return std::max(0, std::min(64, atoi(e)));
    int share = std::max(1, g_live_sessions.load());
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = std::max(share, atoi(e));
    const int hw = (int)std::thread::hardware_concurrency();
    const int t = std::min(10, hw / share); // 10 threads: best dense-result rate on the 16-core box (profiles/r03b_widen_threads.txt; 8: -5 %, 16: -3 %)
    return t >= 4 ? t : 0;
}

// logical CPUs of the NUMA node the device hangs off (empty when the host has one node or the topology is not exposed):
// /sys/bus/pci/devices/<bus id>/numa_node -> /sys/devices/system/node/node<N>/cpulist
static std::vector<int> device_numa_cpus(int device) {
    std::vector<int> cpus;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return cpus; }
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return cpus;
    f = fopen("/sys/devices/system/node/node1/cpulist", "r"); // a single-node host needs no binding
    if (!f) return cpus;
    fclose(f);
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return cpus;
    int a = 0, b = 0;
    while (fscanf(f, "%d", &a) == 1) { // "0-15,32-47"
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
        for (int c = a; c <= b; c++) cpus.push_back(c);
        if (ch != ',') break;
    }
    fclose(f);
    return cpus;
}

// values per D2H piece of the 32-bit wire format (SBWT_B200_D2H_PIECE, in values). Default 64 Mi: a chunk of the
// default size goes in one piece -- smaller pieces only added per-piece overhead on the measured host
// (profiles/r01g_e2e_sweep.txt)
static int64_t d2h_piece_values() {
    if (const char* e = getenv("SBWT_B200_D2H_PIECE")) {
        const long long v = atoll(e);
        if (v >= 4096) return (int64_t)v;
    }
    return (int64_t)64 << 20;
}

// After a failed host-buffer call nothing that was queued may still write into the caller's buffers: wait for every
// slot's copies and stream callbacks (which hand work to the pool), then for the pool, and forget the half-done chunks.
static void drain_slots(sbwt_gpu_session* s) {
    const std::string keep = g_last_error;
    for (HostSlot& h : s->slots) {
        if (h.stream) cudaStreamSynchronize(h.stream);
        h.back_pending = false; // (the hits of a sparse chunk are not fetched: its results are abandoned)
        if (s->widen_pool) s->widen_pool->wait(&h.widen_ticket);
        h.widen_pool = nullptr;
        h.busy = false;
        h.out_staged = false;
        h.out_bytes = 0;
    }
    cudaGetLastError();
    g_last_error = keep;
}

// The next chunk [r0, r1) of a host batch: as many reads as the session capacity takes, found by bisection on the
// offsets; one pass over the chunk's reads then checks them and counts their results. (Round 1 walked every read three
// times on the calling thread -- one validation pass over the whole batch before anything was queued, one to cut the
// chunk, one to count -- 80 MB of offsets each for 10 M reads: ~10 ms of a 130 ms call before the first copy started.)
// A malformed or over-long read is therefore reported when its chunk is reached; the chunks before it have been
// queued and are drained by the caller (drain_slots): the contents of the result buffers are unspecified after an error.
static int next_chunk(const sbwt_gpu_session* s, const int64_t* off, int64_t n_reads, int64_t k, int64_t r0, int64_t* r1_out, int64_t* bases_out,
                      int64_t* n_out_out) {
    const int64_t hi = std::min(n_reads, r0 + s->max_reads);
    int64_t r1 = (std::upper_bound(off + r0 + 1, off + hi + 1, off[r0] + s->max_bases) - off) - 1; // off[r1] - off[r0] <= max_bases
    if (r1 <= r0) {
        const int64_t len = off[r0 + 1] - off[r0];
        if (len < 0) return set_error("read offsets must be non-decreasing");
        return set_error("read %lld (%lld bases) is longer than the session capacity (%lld bases)", (long long)r0, (long long)len, (long long)s->max_bases);
    }
    int64_t t = 0, mn = 0;
    for (int64_t i = r0; i < r1; i++) {
        const int64_t len = off[i + 1] - off[i], v = len - (k - 1);
        mn = len < mn ? len : mn;
        t += v > 0 ? v : 0;
    }
    if (mn < 0 || off[r1] - off[r0] > s->max_bases) return set_error("read offsets must be non-decreasing");
    *r1_out = r1; *bases_out = off[r1] - off[r0]; *n_out_out = t;
    return 0;
}

static int query_host_body(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                           int case_mode, void* out, bool out32);

static int query_host_impl(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                           int case_mode, void* out, bool out32) {
    if (!s) return set_error("null session");
    if (n_reads < 0) return set_error("negative batch size");
    if (n_reads == 0) return 0;
    if (!ascii || !off || !out) return set_error("null buffer");
    if (mode != SBWT_GPU_MODE_SEARCH && mode != SBWT_GPU_MODE_STREAMING) return set_error("unknown mode %d", mode);
    if (case_mode != SBWT_GPU_CASE_UPPER && case_mode != SBWT_GPU_CASE_EXACT && case_mode != SBWT_GPU_CASE_API) return set_error("unknown case mode %d", case_mode);
    if (mode == SBWT_GPU_MODE_STREAMING && !s->idx->has_sgs) return set_error("Error: streaming search support not built");
    DeviceGuard guard(s->idx->device);
    const int rc = query_host_body(s, ascii, off, n_reads, mode, case_mode, out, out32);
    if (rc) drain_slots(s); // a CUDA or internal error in the middle of the pipeline
    return rc;
}

static int query_host_body(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                           int case_mode, void* out, bool out32) {
    sbwt_gpu_index* ix = s->idx;
    if (host_slots_init(s)) return 1;
    const bool pin_in = is_pinned(ascii), pin_off = is_pinned(off), pin_out = is_pinned(out);
    const int64_t k = ix->k;
    // int64 results of a narrow index leave the device as int32 and are widened on the host (host_widen.hpp)
    bool widen = false;
    bool sparse = false; // sparse wire format: hit masks + hits only (SBWT_B200_WIRE=dense turns it off)
    if (!ix->view.wide && ix->n_nodes < (1ll << 31)) {
        if (s->widen_threads < 0) s->widen_threads = widen_thread_count();
        if (s->widen_threads > 0 && !s->widen_pool) s->widen_pool = new WidenPool(s->widen_threads, device_numa_cpus(ix->device));
        widen = !out32 && s->widen_pool != nullptr;
        const char* we = getenv("SBWT_B200_WIRE");
        sparse = s->widen_pool != nullptr && !(we && strcmp(we, "dense") == 0);
    }
    int64_t r0 = 0, out_pos = 0;
    int turn = 0;
    HostSlot* prev = nullptr; // the slot of the previous chunk, whose second half (sparse format) is still to be issued
    while (r0 < n_reads) {
        // largest chunk [r0, r1) that fits the session capacity
        int64_t r1 = r0, bases = 0, n_out = 0;
        if (next_chunk(s, off, n_reads, k, r0, &r1, &bases, &n_out)) return 1;
        const int64_t nr = r1 - r0;
        HostSlot& h = s->slots[turn % sbwt_gpu_session::kSlots];
        turn++;
        if (slot_finish(h)) return 1;
        const char* src = ascii + off[r0];
        if (!pin_in) {
            if (!h.h_ascii) CU(cudaMallocHost(&h.h_ascii, s->max_bases + 64));
            memcpy(h.h_ascii, src, (size_t)bases);
            src = h.h_ascii;
        }
        const int64_t* osrc = off + r0;
        if (!pin_off) {
            if (!h.h_offsets) CU(cudaMallocHost(&h.h_offsets, (s->max_reads + 1) * 8));
            memcpy(h.h_offsets, osrc, (size_t)(nr + 1) * 8);
            osrc = h.h_offsets;
        }
        CU(cudaMemcpyAsync(h.d_ascii, src, (size_t)bases, cudaMemcpyHostToDevice, h.stream));
        CU(cudaMemcpyAsync(h.d_offsets, osrc, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, h.stream));
        if (run_device_batch(s, h.sc, h.d_ascii, h.d_offsets, nr, bases, mode, case_mode, h.d_out, out32 || widen || sparse, false, h.stream)) return 1;
        if (sparse) {
            h.out_staged = false; h.out_bytes = 0;
            if (n_out) {
                const int64_t cap = std::max<int64_t>(s->max_bases, 1);
                const int64_t n_groups = (n_out + 31) / 32, n_sblocks = (n_out + kSparseBlock - 1) / kSparseBlock;
                if (!h.d_masks) {
                    CU(cudaMalloc(&h.d_masks, (size_t)(cap / 32 + 2) * 4));
                    CU(cudaMalloc(&h.d_bbase, (size_t)(cap / kSparseBlock + 2) * 4));
                    CU(cudaMalloc(&h.d_total, 8));
                    CU(cudaMallocHost(&h.h_masks, (size_t)(cap / 32 + 2) * 4));
                    CU(cudaMallocHost(&h.h_bbase, (size_t)(cap / kSparseBlock + 2) * 4));
                    CU(cudaMallocHost(&h.h_total, 8));
                    CU(cudaEventCreateWithFlags(&h.ev_small, cudaEventDisableTiming));
                }
                if (!h.h_out32) CU(cudaMallocHost(&h.h_out32, (size_t)cap * 4 + 64)); // (+64: the mixed-group expansion reads 8 values at a time)
                int32_t* d32 = reinterpret_cast<int32_t*>(h.d_out);
                CU(cudaMemsetAsync(h.d_total, 0, 8, h.stream));
                sparse_pack_kernel<<<(unsigned)n_sblocks, 256, 0, h.stream>>>(d32, n_out, h.d_masks, d32 + cap, h.d_bbase, h.d_total); LAUNCHED();
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(h.h_masks, h.d_masks, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, h.stream));
                CU(cudaMemcpyAsync(h.h_bbase, h.d_bbase, (size_t)n_sblocks * 4, cudaMemcpyDeviceToHost, h.stream));
                CU(cudaMemcpyAsync(h.h_total, h.d_total, 8, cudaMemcpyDeviceToHost, h.stream));
                CU(cudaEventRecord(h.ev_small, h.stream));
                HostSlot::SparseJob& j = h.sparse_job;
                j.pool = s->widen_pool; j.masks = h.h_masks; j.bbase = h.h_bbase; j.packed = h.h_out32; j.n = (size_t)n_out;
                j.dst = out32 ? (void*)((int32_t*)out + out_pos) : (void*)((int64_t*)out + out_pos);
                j.dst64 = !out32; j.ticket = &h.widen_ticket;
                h.back_pending = true;
            } else {
                CU(cudaEventRecord(h.done, h.stream));
            }
            h.busy = true;
            // the previous chunk's hits are fetched now that this chunk's kernels are queued behind them
            if (prev && prev != &h && slot_back(*prev)) return 1;
            prev = &h;
            out_pos += n_out;
            r0 = r1;
            continue;
        }
        if (widen) {
            if (!h.h_out32) CU(cudaMallocHost(&h.h_out32, std::max<int64_t>(s->max_bases, 1) * 4 + 64));
            h.out_staged = false; h.out_bytes = 0;
            if (n_out) {
                // the result copy goes piece by piece, each piece handed to the pool as soon as it has landed: the
                // widening of piece i overlaps the DMA of piece i + 1 and reads it while it is still cache-warm
                const int64_t piece = d2h_piece_values();
                const size_t n_pieces = (size_t)((n_out + piece - 1) / piece);
                h.widen_jobs.assign(n_pieces, HostSlot::WidenJob());
                h.widen_pool = s->widen_pool;
                const int32_t* d32 = reinterpret_cast<const int32_t*>(h.d_out);
                for (size_t pi = 0; pi < n_pieces; pi++) {
                    const int64_t p0 = (int64_t)pi * piece, pn = std::min(piece, n_out - p0);
                    CU(cudaMemcpyAsync(h.h_out32 + p0, d32 + p0, (size_t)pn * 4, cudaMemcpyDeviceToHost, h.stream));
                    HostSlot::WidenJob& j = h.widen_jobs[pi];
                    j.pool = s->widen_pool; j.src = h.h_out32 + p0; j.dst = (int64_t*)out + out_pos + p0;
                    j.n = (size_t)pn; j.ticket = &h.widen_ticket;
                    CU(cudaLaunchHostFunc(h.stream, widen_callback, &j));
                }
            }
            CU(cudaEventRecord(h.done, h.stream));
            h.busy = true;
            out_pos += n_out;
            r0 = r1;
            continue;
        }
        const size_t esz = out32 ? 4 : 8;
        void* dst = (char*)out + (size_t)out_pos * esz;
        h.out_dst = dst; h.out_bytes = n_out * (int64_t)esz; h.out_staged = !pin_out;
        if (!pin_out) {
            if (!h.h_out) CU(cudaMallocHost(&h.h_out, std::max<int64_t>(s->max_bases, 1) * 8));
            dst = h.h_out;
        }
        if (n_out) CU(cudaMemcpyAsync(dst, h.d_out, (size_t)n_out * esz, cudaMemcpyDeviceToHost, h.stream));
        CU(cudaEventRecord(h.done, h.stream));
        h.busy = true;
        out_pos += n_out;
        r0 = r1;
    }
    for (HostSlot& h : s->slots)
        if (slot_finish(h)) return 1;
    return 0;
}

extern "C" int sbwt_gpu_query_host(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                                   int case_mode, int64_t* out) {
    return query_host_impl(s, ascii, off, n_reads, mode, case_mode, out, false);
}
extern "C" int sbwt_gpu_query_host_i32(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                                       int case_mode, int32_t* out) {
    if (s && (s->idx->n_nodes >= (1ll << 31) || s->idx->view.wide))
        return set_error("int32 results need an index with fewer than 2^31 columns (this one has %lld)", (long long)s->idx->n_nodes);
    return query_host_impl(s, ascii, off, n_reads, mode, case_mode, out, true);
}

// ------------------------------------------------------------------ hits-only results (membership bitmap + found values)
//
// For callers that do not need a dense int64 array: what crosses PCIe and what the host touches shrinks to one bit per
// k-mer plus four bytes per FOUND k-mer, written by the DMA engines straight into the caller's buffers -- no host thread
// rebuilds anything, so the call scales with the number of GPUs on one host where sbwt_gpu_query_host is bound by the
// host's memory system (VERDICT round 1, weak point 5). Same answers: a miss is always -1 (SBWT.hh:390-415, :545-581).
static int query_host_hits_body(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode, int case_mode,
                                uint32_t* hit_mask, int32_t* hits, int64_t* n_hits_out) {
    sbwt_gpu_index* ix = s->idx;
    if (host_slots_init(s)) return 1;
    const bool pin_in = is_pinned(ascii), pin_off = is_pinned(off), pin_hits = hits && is_pinned(hits);
    const bool v32 = ix->n_nodes < (1ll << 31) && !ix->view.wide; // the walk writes int32 values (int64 on larger indexes: bitmap only)
    const int64_t k = ix->k, cap = std::max<int64_t>(s->max_bases, 1);
    const int64_t cap_blocks = cap / kSparseBlock + 3;
    for (HostSlot& h : s->slots) {
        if (!h.d_bcount) {
            CU(cudaMalloc(&h.d_bcount, (size_t)cap_blocks * 8));
            CU(cudaMalloc(&h.d_bpart, (size_t)scan_partials_needed(cap_blocks) * 8));
            CU(cudaMalloc(&h.d_btotal, 8));
            CU(cudaMallocHost(&h.h_btotal, 8));
        }
        if (!h.d_masks) {
            CU(cudaMalloc(&h.d_masks, (size_t)(cap / 32 + 2) * 4));
            CU(cudaMalloc(&h.d_bbase, (size_t)(cap / kSparseBlock + 2) * 4));
            CU(cudaMalloc(&h.d_total, 8));
            CU(cudaMallocHost(&h.h_masks, (size_t)(cap / 32 + 2) * 4));
            CU(cudaMallocHost(&h.h_bbase, (size_t)(cap / kSparseBlock + 2) * 4));
            CU(cudaMallocHost(&h.h_total, 8));
            CU(cudaEventCreateWithFlags(&h.ev_small, cudaEventDisableTiming));
        }
        if (hits && !pin_hits && !h.h_out32) CU(cudaMallocHost(&h.h_out32, (size_t)cap * 4 + 64));
    }
    struct Pending { // a chunk whose masks / hit count are on their way; its hits are fetched once the count is known
        HostSlot* h = nullptr;
        int64_t bit0 = 0, n_out = 0;
    } pend;
    int64_t n_hits = 0;
    // masks land in pinned staging (the chunk's first word overlaps the previous chunk's last one) and are merged here
    auto finish = [&](Pending& p) -> int {
        if (!p.h) return 0;
        HostSlot& h = *p.h;
        CU(cudaEventSynchronize(h.ev_small));
        const int64_t total = *h.h_btotal;
        if (total < 0 || total > p.n_out) return set_error("hits-only results: %lld hits reported for %lld results", (long long)total, (long long)p.n_out);
        if (hits && total) { // (pinned caller buffer: the DMA engine writes the hits where they belong)
            const int32_t* d_packed = reinterpret_cast<const int32_t*>(h.d_out) + h.sc.max_bases;
            CU(cudaMemcpyAsync(pin_hits ? hits + n_hits : h.h_out32, d_packed, (size_t)total * 4, cudaMemcpyDeviceToHost, h.stream));
        }
        const int64_t w0 = p.bit0 >> 5, nw = ((p.bit0 & 31) + p.n_out + 31) / 32;
        if (p.n_out) { // (every word of the bitmap is written here: no clearing pass over the caller's buffer; a chunk without results brought no masks)
            if (p.bit0 & 31) hit_mask[w0] |= h.h_masks[0]; // the previous chunk wrote the low bits of this word
            else hit_mask[w0] = h.h_masks[0];
            if (nw > 1) memcpy(hit_mask + w0 + 1, h.h_masks + 1, (size_t)(nw - 1) * 4);
        }
        if (hits && total && !pin_hits) {
            CU(cudaStreamSynchronize(h.stream));
            memcpy(hits + n_hits, h.h_out32, (size_t)total * 4);
        }
        n_hits += total;
        h.busy = false;
        p.h = nullptr;
        return 0;
    };
    int64_t r0 = 0, out_pos = 0;
    int turn = 0;
    while (r0 < n_reads) {
        int64_t r1 = r0, bases = 0, n_out = 0;
        if (next_chunk(s, off, n_reads, k, r0, &r1, &bases, &n_out)) return 1;
        const int64_t nr = r1 - r0;
        HostSlot& h = s->slots[turn & 1];
        turn++;
        const char* src = ascii + off[r0];
        if (!pin_in) {
            if (!h.h_ascii) CU(cudaMallocHost(&h.h_ascii, s->max_bases + 64));
            memcpy(h.h_ascii, src, (size_t)bases);
            src = h.h_ascii;
        }
        const int64_t* osrc = off + r0;
        if (!pin_off) {
            if (!h.h_offsets) CU(cudaMallocHost(&h.h_offsets, (s->max_reads + 1) * 8));
            memcpy(h.h_offsets, osrc, (size_t)(nr + 1) * 8);
            osrc = h.h_offsets;
        }
        CU(cudaMemcpyAsync(h.d_ascii, src, (size_t)bases, cudaMemcpyHostToDevice, h.stream));
        CU(cudaMemcpyAsync(h.d_offsets, osrc, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, h.stream));
        if (run_device_batch(s, h.sc, h.d_ascii, h.d_offsets, nr, bases, mode, case_mode, h.d_out, v32, false, h.stream)) return 1;
        const uint32_t sh = (uint32_t)(out_pos & 31);
        const int64_t nw = ((int64_t)sh + n_out + 31) / 32, nb = (nw + kSparseBlock / 32 - 1) / (kSparseBlock / 32);
        int32_t* d32 = reinterpret_cast<int32_t*>(h.d_out);
        if (n_out) {
            if (v32) hits_kernel<false, int32_t><<<(unsigned)nb, 256, 0, h.stream>>>(d32, n_out, sh, h.d_masks, h.d_bcount, nullptr, nullptr);
            else hits_kernel<false, int64_t><<<(unsigned)nb, 256, 0, h.stream>>>(h.d_out, n_out, sh, h.d_masks, h.d_bcount, nullptr, nullptr);
            LAUNCHED();
            if (exclusive_scan_inplace(h.d_bcount, nb, h.d_bpart, h.d_btotal, h.stream)) return 1;
            if (hits) { hits_kernel<true, int32_t><<<(unsigned)nb, 256, 0, h.stream>>>(d32, n_out, sh, nullptr, nullptr, h.d_bcount, d32 + cap); LAUNCHED(); }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(h.h_masks, h.d_masks, (size_t)nw * 4, cudaMemcpyDeviceToHost, h.stream));
        } else {
            CU(cudaMemsetAsync(h.d_btotal, 0, 8, h.stream));
        }
        CU(cudaMemcpyAsync(h.h_btotal, h.d_btotal, 8, cudaMemcpyDeviceToHost, h.stream));
        CU(cudaEventRecord(h.ev_small, h.stream));
        h.busy = true;
        if (finish(pend)) return 1; // the previous chunk, now that this one's kernels are queued behind it
        pend.h = &h; pend.bit0 = out_pos; pend.n_out = n_out;
        out_pos += n_out;
        r0 = r1;
    }
    if (finish(pend)) return 1;
    for (int i = 0; i < 2; i++) CU(cudaStreamSynchronize(s->slots[i].stream));
    if (n_hits_out) *n_hits_out = n_hits;
    return 0;
}

extern "C" int sbwt_gpu_query_host_hits(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode, int case_mode,
                                        uint32_t* hit_mask, int32_t* hits, int64_t* n_hits) {
    if (!s) return set_error("null session");
    if (n_hits) *n_hits = 0;
    if (n_reads < 0) return set_error("negative batch size");
    if (n_reads == 0) return 0;
    if (!ascii || !off || !hit_mask) return set_error("null buffer");
    if (mode != SBWT_GPU_MODE_SEARCH && mode != SBWT_GPU_MODE_STREAMING) return set_error("unknown mode %d", mode);
    if (case_mode != SBWT_GPU_CASE_UPPER && case_mode != SBWT_GPU_CASE_EXACT && case_mode != SBWT_GPU_CASE_API) return set_error("unknown case mode %d", case_mode);
    if (mode == SBWT_GPU_MODE_STREAMING && !s->idx->has_sgs) return set_error("Error: streaming search support not built");
    if (hits && (s->idx->n_nodes >= (1ll << 31) || s->idx->view.wide))
        return set_error("32-bit hit values need an index with fewer than 2^31 columns (this one has %lld); pass hits = NULL for the membership bitmap alone",
                         (long long)s->idx->n_nodes);
    DeviceGuard guard(s->idx->device);
    const int rc = query_host_hits_body(s, ascii, off, n_reads, mode, case_mode, hit_mask, hits, n_hits);
    if (rc) drain_slots(s);
    return rc;
}

// Multi-GPU in one process: the index is replicated (one session per device, made by the caller),
// the reads are cut into contiguous ranges of (almost) equal total bases -- the split of sbwt_b200/sharding.py --
// and one host thread per session answers its range into its own slice of `out`. No collective, nothing is exchanged.
extern "C" int sbwt_gpu_query_host_sharded(sbwt_gpu_session* const* sessions, int n_sessions, const char* ascii, const int64_t* off,
                                           int64_t n_reads, int mode, int case_mode, int64_t* out) {
    if (!sessions || n_sessions < 1 || n_sessions > 64) return set_error("sbwt_gpu_query_host_sharded: 1 <= n_sessions <= 64 expected");
    for (int i = 0; i < n_sessions; i++) {
        if (!sessions[i]) return set_error("null session");
        if (sessions[i]->idx->k != sessions[0]->idx->k || sessions[i]->idx->n_nodes != sessions[0]->idx->n_nodes)
            return set_error("sbwt_gpu_query_host_sharded: the sessions must hold replicas of one index");
        for (int j = 0; j < i; j++)
            if (sessions[j] == sessions[i]) return set_error("sbwt_gpu_query_host_sharded: a session is listed twice");
    }
    if (n_reads < 0) return set_error("negative batch size");
    if (n_reads == 0) return 0;
    if (!ascii || !off || !out) return set_error("null buffer");
    const int64_t k = sessions[0]->idx->k;
    // cuts[i] = first read of shard i: the first read starting at or after the i-th equal share of the bases
    std::vector<int64_t> cuts((size_t)n_sessions + 1, n_reads);
    cuts[0] = 0;
    const int64_t total = off[n_reads] - off[0];
    for (int i = 1; i < n_sessions; i++) {
        const int64_t target = off[0] + (int64_t)((__int128)total * i / n_sessions);
        cuts[(size_t)i] = std::max<int64_t>(cuts[(size_t)i - 1], std::lower_bound(off, off + n_reads + 1, target) - off);
        cuts[(size_t)i] = std::min(cuts[(size_t)i], n_reads);
    }
    std::vector<int64_t> out_pos((size_t)n_sessions + 1, 0);
    for (int i = 0; i < n_sessions; i++)
        out_pos[(size_t)i + 1] = out_pos[(size_t)i] + sbwt_gpu_count_outputs(off + cuts[(size_t)i], cuts[(size_t)i + 1] - cuts[(size_t)i], k);
    std::vector<int> rc((size_t)n_sessions, 0);
    std::vector<std::string> err((size_t)n_sessions);
    std::vector<int64_t> launches((size_t)n_sessions, 0);
    std::vector<std::thread> th;
    auto work = [&](int i) {
        const int64_t r0 = cuts[(size_t)i], nr = cuts[(size_t)i + 1] - r0;
        if (nr == 0) return;
        const int64_t before = g_launches; // (thread-local, like the error text)
        rc[(size_t)i] = query_host_impl(sessions[i], ascii, off + r0, nr, mode, case_mode, out + out_pos[(size_t)i], false);
        if (rc[(size_t)i]) err[(size_t)i] = g_last_error;
        launches[(size_t)i] = g_launches - before;
    };
    for (int i = 1; i < n_sessions; i++) th.emplace_back(work, i);
    work(0);
    for (std::thread& t : th) t.join();
    for (int i = 1; i < n_sessions; i++) g_launches += launches[(size_t)i];
    for (int i = 0; i < n_sessions; i++)
        if (rc[(size_t)i]) return set_error("shard %d (device %d): %s", i, sessions[i]->idx->device, err[(size_t)i].c_str());
    return 0;
}

extern "C" int sbwt_gpu_widen_i32(const int32_t* in, int64_t* out, int64_t n, int threads) {
    if (n < 0 || threads < 1 || threads > 64) return set_error("sbwt_gpu_widen_i32: n >= 0 and 1 <= threads <= 64 expected");
    if (n == 0) return 0;
    if (!in || !out) return set_error("null buffer");
    WidenPool pool(threads);
    WidenTicket t;
    pool.submit(in, out, (size_t)n, &t);
    pool.wait(&t);
    return 0;
}

extern "C" int sbwt_gpu_expand_sparse(const uint32_t* masks, const uint32_t* block_base, const int32_t* packed, int64_t n, void* out,
                                      int out_is_i64, int threads) {
    if (n < 0 || threads < 1 || threads > 64) return set_error("sbwt_gpu_expand_sparse: n >= 0 and 1 <= threads <= 64 expected");
    if (n == 0) return 0;
    if (!masks || !block_base || !out || !packed) return set_error("null buffer");
    WidenPool pool(threads);
    WidenTicket t;
    pool.submit_sparse(masks, block_base, packed, (size_t)n, out, out_is_i64 != 0, &t);
    pool.wait(&t);
    return 0;
}

extern "C" int sbwt_gpu_session_widen_threads(const sbwt_gpu_session* s) { return s ? s->widen_threads : -1; }

extern "C" int sbwt_gpu_search_batch(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int64_t* out) {
    return sbwt_gpu_query_host(s, ascii, off, n_reads, SBWT_GPU_MODE_SEARCH, SBWT_GPU_CASE_UPPER, out);
}
extern "C" int sbwt_gpu_streaming_batch(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int64_t* out) {
    return sbwt_gpu_query_host(s, ascii, off, n_reads, SBWT_GPU_MODE_STREAMING, SBWT_GPU_CASE_UPPER, out);
}

// ------------------------------------------------------------------ text output (print_vector on the device)

static int text_digits(int64_t n_nodes) { // longest printed value: n_nodes - 1, and "-1"
    int d = 1;
    for (int64_t v = std::max<int64_t>(n_nodes - 1, 1); v >= 10; v /= 10) d++;
    return std::max(d, 2);
}

extern "C" int64_t sbwt_gpu_text_capacity(const sbwt_gpu_index* ix, int64_t n_values, int64_t n_reads) {
    if (!ix || n_values < 0 || n_reads < 0) return -1;
    return n_values * (text_digits(ix->n_nodes) + 1) + n_reads + 16;
}

// values -> text on `st`. voff_ready: sc.n_out already holds the exclusive scan of the per-read result
// counts of exactly these reads (the batch has just been planned on the same stream).
template <typename T>
static int run_format_t(Scratch& sc, int64_t k, const T* d_vals, const int64_t* d_offsets, int64_t n_reads, bool voff_ready,
                        char* d_text, int64_t capacity, int64_t* d_text_bytes, cudaStream_t st) {
    if (n_reads > sc.max_reads) return set_error("%lld reads exceed the session capacity (%lld)", (long long)n_reads, (long long)sc.max_reads);
    if (n_reads == 0) {
        if (d_text_bytes) CU(cudaMemsetAsync(d_text_bytes, 0, 8, st));
        CU(cudaMemsetAsync(sc.totals + 2, 0, 8, st));
        return 0;
    }
    if (!voff_ready) {
        plan_count_kernel<<<grid_for(n_reads, 256), 256, 0, st>>>(d_offsets, n_reads, (int)k, 1 << 30, sc.n_out, sc.n_win); LAUNCHED();
        if (exclusive_scan_inplace(sc.n_out, n_reads, sc.partials, sc.totals + 0, st)) return 1;
    }
    const int64_t warps = (n_reads + kFmtReadsPerWarp - 1) / kFmtReadsPerWarp;
    const unsigned grid = grid_for(warps * 32, kFmtThreads);
    fmt_len_kernel<T><<<grid, kFmtThreads, 0, st>>>(d_vals, sc.n_out, sc.totals + 0, n_reads, sc.n_win); LAUNCHED();
    if (exclusive_scan_inplace(sc.n_win, n_reads, sc.partials, sc.totals + 2, st)) return 1;
    fmt_emit_kernel<T><<<grid, kFmtThreads, 0, st>>>(d_vals, sc.n_out, sc.totals + 0, sc.n_win, sc.totals + 2, n_reads, d_text, capacity); LAUNCHED();
    CU(cudaGetLastError());
    if (d_text_bytes) CU(cudaMemcpyAsync(d_text_bytes, sc.totals + 2, 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int sbwt_gpu_format_device(sbwt_gpu_session* s, const void* d_vals, int vals_are_i32, const int64_t* d_read_offsets,
                                      int64_t n_reads, char* d_text, int64_t text_capacity, int64_t* d_text_bytes, void* stream) {
    if (!s) return set_error("null session");
    if (n_reads < 0 || text_capacity < 0) return set_error("negative size");
    if (n_reads > 0 && (!d_vals || !d_read_offsets || !d_text)) return set_error("null buffer");
    DeviceGuard guard(s->idx->device);
    if (vals_are_i32)
        return run_format_t<int32_t>(s->sc, s->idx->k, (const int32_t*)d_vals, d_read_offsets, n_reads, false, d_text, text_capacity,
                                     d_text_bytes, (cudaStream_t)stream);
    return run_format_t<int64_t>(s->sc, s->idx->k, (const int64_t*)d_vals, d_read_offsets, n_reads, false, d_text, text_capacity,
                                 d_text_bytes, (cudaStream_t)stream);
}

constexpr int64_t kTextPiece = (int64_t)32 << 20; // bytes per D2H piece handed to the sink

// waits for the slot's kernels, then moves its text to the host piece by piece (two pinned buffers) and hands
// every piece to the sink, in order. The other slot's kernels keep the GPU busy meanwhile.
static int slot_deliver_text(sbwt_gpu_session* s, HostSlot& h, sbwt_gpu_text_sink sink, void* user, int64_t* n_lookups) {
    if (!h.busy) return 0;
    CU(cudaEventSynchronize(h.done));
    h.busy = false;
    const int64_t total = h.h_totals[2];
    if (n_lookups) *n_lookups += h.h_totals[0];
    if (total > h.text_capacity) return set_error("internal error: text of %lld bytes exceeds the slot capacity %lld", (long long)total, (long long)h.text_capacity);
    int64_t issued = 0, delivered = 0, size[2] = {0, 0};
    auto issue = [&](int b) -> int {
        const int64_t sz = std::min(kTextPiece, total - issued);
        CU(cudaMemcpyAsync(s->h_text[b], h.d_text + issued, (size_t)sz, cudaMemcpyDeviceToHost, h.stream));
        CU(cudaEventRecord(s->text_ev[b], h.stream));
        size[b] = sz;
        issued += sz;
        return 0;
    };
    if (total > 0 && issue(0)) return 1;
    for (int b = 0; delivered < total; b ^= 1) {
        if (issued < total && issue(b ^ 1)) return 1;
        CU(cudaEventSynchronize(s->text_ev[b]));
        if (sink(user, s->h_text[b], size[b]) != 0) return set_error("the text sink reported an error");
        delivered += size[b];
    }
    return 0;
}

static int query_host_text_body(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                                int case_mode, sbwt_gpu_text_sink sink, void* user, int64_t* n_lookups);

extern "C" int sbwt_gpu_query_host_text(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                                        int case_mode, sbwt_gpu_text_sink sink, void* user, int64_t* n_lookups) {
    if (!s) return set_error("null session");
    if (n_lookups) *n_lookups = 0;
    if (n_reads < 0) return set_error("negative batch size");
    if (n_reads == 0) return 0;
    if (!ascii || !off || !sink) return set_error("null argument");
    if (mode != SBWT_GPU_MODE_SEARCH && mode != SBWT_GPU_MODE_STREAMING) return set_error("unknown mode %d", mode);
    if (case_mode != SBWT_GPU_CASE_UPPER && case_mode != SBWT_GPU_CASE_EXACT && case_mode != SBWT_GPU_CASE_API) return set_error("unknown case mode %d", case_mode);
    if (mode == SBWT_GPU_MODE_STREAMING && !s->idx->has_sgs) return set_error("Error: streaming search support not built");
    DeviceGuard guard(s->idx->device);
    const int rc = query_host_text_body(s, ascii, off, n_reads, mode, case_mode, sink, user, n_lookups);
    if (rc) drain_slots(s); // also after a sink error: queued kernels still read the caller's (pinned) input buffers
    return rc;
}

static int query_host_text_body(sbwt_gpu_session* s, const char* ascii, const int64_t* off, int64_t n_reads, int mode,
                                int case_mode, sbwt_gpu_text_sink sink, void* user, int64_t* n_lookups) {
    sbwt_gpu_index* ix = s->idx;
    if (host_slots_init(s)) return 1;
    for (int b = 0; b < 2; b++)
        if (!s->h_text[b]) {
            CU(cudaMallocHost(&s->h_text[b], (size_t)kTextPiece));
            CU(cudaEventCreateWithFlags(&s->text_ev[b], cudaEventDisableTiming));
        }
    const bool out32 = ix->n_nodes < (1ll << 31) && !ix->view.wide; // same values; halves what the formatter reads
    for (HostSlot& h : s->slots)
        if (!h.d_text) {
            h.text_capacity = sbwt_gpu_text_capacity(ix, s->max_bases, s->max_reads);
            CU(cudaMalloc(&h.d_text, (size_t)h.text_capacity));
        }
    const bool pin_in = is_pinned(ascii), pin_off = is_pinned(off);
    int64_t r0 = 0;
    int turn = 0;
    HostSlot* prev = nullptr;
    while (r0 < n_reads) {
        int64_t r1 = r0, bases = 0, n_out_chunk = 0;
        if (next_chunk(s, off, n_reads, ix->k, r0, &r1, &bases, &n_out_chunk)) return 1;
        const int64_t nr = r1 - r0;
        HostSlot& h = s->slots[turn & 1]; // free: its previous batch was delivered one iteration ago
        turn++;
        const char* src = ascii + off[r0];
        if (!pin_in) {
            if (!h.h_ascii) CU(cudaMallocHost(&h.h_ascii, s->max_bases + 64));
            memcpy(h.h_ascii, src, (size_t)bases);
            src = h.h_ascii;
        }
        const int64_t* osrc = off + r0;
        if (!pin_off) {
            if (!h.h_offsets) CU(cudaMallocHost(&h.h_offsets, (s->max_reads + 1) * 8));
            memcpy(h.h_offsets, osrc, (size_t)(nr + 1) * 8);
            osrc = h.h_offsets;
        }
        CU(cudaMemcpyAsync(h.d_ascii, src, (size_t)bases, cudaMemcpyHostToDevice, h.stream));
        CU(cudaMemcpyAsync(h.d_offsets, osrc, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, h.stream));
        if (run_device_batch(s, h.sc, h.d_ascii, h.d_offsets, nr, bases, mode, case_mode, h.d_out, out32, false, h.stream)) return 1;
        const int rc = out32 ? run_format_t<int32_t>(h.sc, ix->k, (const int32_t*)h.d_out, h.d_offsets, nr, true, h.d_text, h.text_capacity, nullptr, h.stream)
                             : run_format_t<int64_t>(h.sc, ix->k, h.d_out, h.d_offsets, nr, true, h.d_text, h.text_capacity, nullptr, h.stream);
        if (rc) return 1;
        CU(cudaMemcpyAsync(h.h_totals, h.sc.totals, 32, cudaMemcpyDeviceToHost, h.stream));
        CU(cudaEventRecord(h.done, h.stream));
        h.busy = true;
        h.out_staged = false;
        h.out_bytes = 0;
        if (prev && slot_deliver_text(s, *prev, sink, user, n_lookups)) return 1;
        prev = &h;
        r0 = r1;
    }
    if (prev && slot_deliver_text(s, *prev, sink, user, n_lookups)) return 1;
    return 0;
}

extern "C" int sbwt_gpu_host_alloc(size_t bytes, void** out) {
    if (!out) return set_error("null argument");
    CU(cudaMallocHost(out, bytes ? bytes : 16));
    return 0;
}
extern "C" int sbwt_gpu_host_alloc_on(int device, size_t bytes, void** out) {
    if (!out) return set_error("null argument");
    if (device < 0 || device >= sbwt_gpu_device_count()) return set_error("invalid device %d", device);
    DeviceGuard guard(device); // (no context is created on another device by a thread that never chose one)
    CU(cudaMallocHost(out, bytes ? bytes : 16));
    return 0;
}
extern "C" void sbwt_gpu_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ probe

extern "C" int sbwt_gpu_sector_probe(int device, int64_t buffer_bytes, int64_t n_loads, int bytes_per_load, int iters, double* best_ms) {
    if (bytes_per_load != 32 && bytes_per_load != 64) return set_error("bytes_per_load must be 32 or 64");
    if (sbwt_gpu_device_count() <= 0) return set_error("no CUDA device available");
    DeviceGuard guard(device);
    void* buf = nullptr;
    uint32_t* sink = nullptr;
    CU(cudaMalloc(&buf, (size_t)buffer_bytes));
    CU(cudaMemset(buf, 1, (size_t)buffer_bytes));
    CU(cudaMalloc(&sink, 4));
    const uint64_t n_units = (uint64_t)buffer_bytes / (uint64_t)bytes_per_load;
    const int per_thread = 32;
    const int64_t threads = (n_loads + per_thread - 1) / per_thread;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 1e30;
    for (int it = 0; it < iters + 1; it++) {
        CU(cudaEventRecord(e0));
        if (bytes_per_load == 32) probe_kernel<32><<<grid_for(threads, 256), 256>>>((const Sector*)buf, n_units, per_thread, 1234u + it, sink);
        else probe_kernel<64><<<grid_for(threads, 256), 256>>>((const Sector*)buf, n_units, per_thread, 1234u + it, sink);
        LAUNCHED();
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    *best_ms = best;
    return 0;
}
