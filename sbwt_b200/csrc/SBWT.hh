// SBWT.hh -- host-side C++ mirror of the reference's query surface for the GPU path.
//
// Same names, argument meaning and error behaviour as the reference's
//   template <typename subset_rank_t> class SBWT          (include/sbwt/SBWT.hh:31-332)
//   class SubsetMatrixRank<bitvector_t, rank_support_t>    (include/sbwt/SubsetMatrixRank.hh:13-127)
// for the one instantiation in scope, plain_matrix_sbwt_t (include/sbwt/variants.hh:19):
//   sbwt::SBWT<sbwt::GpuSubsetMatrixRank>
// Every query method is a thin call into the C ABI of include/sbwt_b200.h; nothing is computed
// on the CPU. Construction from reads (KMC) and the select-support methods are out of
// scope.
//
// Differences a caller can observe:
//   * none in the answers. (k > 64 is answered by the literal kernels of long_kmer_kernels.cuh: correct, not tuned.)
//   * streaming_search(const char*, len) gives the reference's answer, mixed case included (SBWT_GPU_CASE_API;
//     tests/golden/*/mixed_case.*, written by the reference's own method) -- except on an index that violates the
//     edge invariant (only hand-made files do), where a lower-case base is a miss everywhere (CASE_EXACT).
//   * the batch methods (search_batch / streaming_search_batch) are additions: one call per
//     k-mer through a GPU is correct but slow, the batch forms are what `sbwt search` uses.
#pragma once

#include <cstdint>
#include <fstream>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sbwt_b200.h"
#include "sbwt_file.hpp"

namespace sbwt {

inline void gpu_check(int rc) {
    if (rc != 0) throw std::runtime_error(sbwt_gpu_last_error()); // the reference's error style: std::runtime_error
}

// The subset-rank structure of the GPU path. Satisfies the concept SBWT<> expects
// (SubsetMatrixRank.hh:31-125): rank, contains, construction from four bit vectors, serialize, load.
// The bit vectors live here on the host only for contains()/serialize(); rank() is answered by
// the device index.
class GpuSubsetMatrixRank {
public:
    std::vector<uint64_t> A_bits, C_bits, G_bits, T_bits; // LSB-first words, n_bits bits each
    int64_t n_bits = 0;
    sbwt_gpu_index* device_index = nullptr; // not owned; set by SBWT<>

    GpuSubsetMatrixRank() {}
    GpuSubsetMatrixRank(const std::vector<uint64_t>& A, const std::vector<uint64_t>& C, const std::vector<uint64_t>& G,
                        const std::vector<uint64_t>& T, int64_t n_bits)
        : A_bits(A), C_bits(C), G_bits(G), T_bits(T), n_bits(n_bits) {}

    // Count of character c in subsets up to pos, not including pos (SubsetMatrixRank.hh:31-37).
    int64_t rank(int64_t pos, char c) const {
        if (!device_index) throw std::runtime_error("GpuSubsetMatrixRank: no device index attached");
        int64_t out = 0;
        gpu_check(sbwt_gpu_rank(device_index, &pos, &c, 1, &out));
        return out;
    }

    bool contains(int64_t pos, char c) const { // SubsetMatrixRank.hh:39-48
        const std::vector<uint64_t>* v = c == 'A' ? &A_bits : c == 'C' ? &C_bits : c == 'G' ? &G_bits : c == 'T' ? &T_bits : nullptr;
        return v ? (((*v)[(size_t)pos >> 6] >> (pos & 63)) & 1) != 0 : false;
    }

    int64_t serialize(std::ostream& os) const { // SubsetMatrixRank.hh:86-100: 4 bit vectors, then 4 rank supports
        using namespace sbwt_b200::detail;
        int64_t written = 0;
        for (const auto* v : {&A_bits, &C_bits, &G_bits, &T_bits}) written += wr_words(os, n_bits, *v);
        for (const auto* v : {&A_bits, &C_bits, &G_bits, &T_bits}) {
            std::vector<uint64_t> rs = sbwt_b200::rank_support_v5_words(*v, n_bits);
            written += wr_words(os, (int64_t)rs.size() * 64, rs);
        }
        return written;
    }

    void load(std::istream& is) { // SubsetMatrixRank.hh:102-125
        using namespace sbwt_b200::detail;
        int64_t nb[4], rsb;
        A_bits = rd_words(is, &nb[0]); C_bits = rd_words(is, &nb[1]); G_bits = rd_words(is, &nb[2]); T_bits = rd_words(is, &nb[3]);
        if (nb[0] != nb[1] || nb[0] != nb[2] || nb[0] != nb[3]) throw std::runtime_error("Error: Corrupt index file (bit vector lengths differ).");
        n_bits = nb[0];
        for (int c = 0; c < 4; c++) rd_words(is, &rsb); // the file's rank_support_v5 words are not needed on the device
    }
};

template <typename subset_rank_t>
class SBWT {
private:
    subset_rank_t subset_rank;
    std::vector<uint64_t> suffix_group_starts; // empty = no streaming support
    std::vector<int64_t> C;
    std::vector<std::pair<int64_t, int64_t>> kmer_prefix_precalc;
    int64_t precalc_k = 0;
    int64_t n_nodes = 0, n_kmers = 0, k = 0;

    int device = 0;
    sbwt_gpu_index* dev = nullptr;
    mutable sbwt_gpu_session* session = nullptr;
    mutable int64_t session_bases = 0, session_reads = 0;

    void release() {
        if (session) sbwt_gpu_session_destroy(session);
        if (dev) sbwt_gpu_index_destroy(dev);
        session = nullptr;
        dev = nullptr;
        subset_rank.device_index = nullptr;
    }

    // (re)creates the device index from the host members; a missing precalc table is computed on the device
    void attach_device(bool compute_precalc) {
        release();
        const uint64_t* bits[4] = {subset_rank.A_bits.data(), subset_rank.C_bits.data(), subset_rank.G_bits.data(), subset_rank.T_bits.data()};
        const int64_t* pre = (precalc_k > 0 && !compute_precalc) ? &kmer_prefix_precalc[0].first : nullptr;
        static_assert(sizeof(std::pair<int64_t, int64_t>) == 16, "pair layout");
        gpu_check(sbwt_gpu_index_create(bits, suffix_group_starts.empty() ? nullptr : suffix_group_starts.data(), n_nodes, n_kmers, k,
                                        C.data(), pre, precalc_k, device, &dev));
        subset_rank.device_index = dev;
        if (precalc_k > 0 && compute_precalc) {
            kmer_prefix_precalc.resize((size_t)1 << (2 * precalc_k));
            gpu_check(sbwt_gpu_index_get_precalc(dev, &kmer_prefix_precalc[0].first));
        }
    }

    sbwt_gpu_session* get_session(int64_t bases, int64_t reads) const {
        if (!dev) throw std::runtime_error("SBWT: index not loaded");
        if (!session || bases > session_bases || reads > session_reads) {
            if (session) sbwt_gpu_session_destroy(session);
            session = nullptr;
            // capacities grow in powers of two: batches of slightly different sizes (variable-length reads, --devices slices)
            // must not free and reallocate several GB of device and pinned buffers inside the query path
            auto pow2 = [](int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; };
            session_bases = std::min<int64_t>(std::max<int64_t>(pow2(std::max<int64_t>(bases, session_bases)), 1 << 20), (int64_t)0xFFFF0000ll);
            session_reads = std::min<int64_t>(std::max<int64_t>(pow2(std::max<int64_t>(reads, session_reads)), 1 << 14), (int64_t)0x7FFF0000ll);
            gpu_check(sbwt_gpu_session_create(dev, session_bases, session_reads, &session));
        }
        return session;
    }

    static int64_t popcount_words(const std::vector<uint64_t>& w) {
        int64_t s = 0;
        for (uint64_t x : w) s += __builtin_popcountll(x);
        return s;
    }

public:
    SBWT() {}
    explicit SBWT(int device) : device(device) {}
    SBWT(const SBWT&) = delete;
    SBWT& operator=(const SBWT&) = delete;
    ~SBWT() { release(); }

    // SBWT(A_bits, C_bits, G_bits, T_bits, streaming_support, k, number_of_kmers, precalc_k) (SBWT.hh:336-353).
    // Bit vectors are LSB-first 64-bit words of n_nodes bits; streaming_support may be empty.
    SBWT(const std::vector<uint64_t>& A_bits, const std::vector<uint64_t>& C_bits, const std::vector<uint64_t>& G_bits,
         const std::vector<uint64_t>& T_bits, const std::vector<uint64_t>& streaming_support, int64_t n_nodes, int64_t k,
         int64_t number_of_kmers, int64_t precalc_k, int device = 0)
        : subset_rank(A_bits, C_bits, G_bits, T_bits, n_nodes), suffix_group_starts(streaming_support), precalc_k(precalc_k),
          n_nodes(n_nodes), n_kmers(number_of_kmers), k(k), device(device) {
        C = {1, 0, 0, 0}; // one ghost dollar into the root (SBWT.hh:346)
        C[1] = C[0] + popcount_words(A_bits);
        C[2] = C[1] + popcount_words(C_bits);
        C[3] = C[2] + popcount_words(G_bits);
        if (precalc_k > 20) throw std::runtime_error("Error: Can't precalc longer than 20-mers (would take over 4^20 = 2^40 bytes");
        if (precalc_k > k) throw std::runtime_error("Error: Precalc length is longer than k (" + std::to_string(precalc_k) + " > " + std::to_string(k) + ")");
        attach_device(true);
    }

    // Accessors (SBWT.hh:111-157, 253)
    const subset_rank_t& get_subset_rank_structure() const { return subset_rank; }
    const std::vector<uint64_t>& get_streaming_support() const { return suffix_group_starts; }
    const std::vector<int64_t>& get_C_array() const { return C; }
    const std::vector<std::pair<int64_t, int64_t>>& get_precalc() const { return kmer_prefix_precalc; }
    int64_t get_precalc_k() const { return precalc_k; }
    int64_t number_of_subsets() const { return n_nodes; }
    int64_t number_of_kmers() const { return n_kmers; }
    int64_t get_k() const { return k; }
    bool has_streaming_query_support() const { return !suffix_group_starts.empty(); }
    sbwt_gpu_index* get_device_index() const { return dev; }

    // Precalculate the intervals of all p-mers, on the device (SBWT.hh:617-645).
    void do_kmer_prefix_precalc(int64_t prefix_length) {
        if (prefix_length == 0) return;
        if (prefix_length > 20) throw std::runtime_error("Error: Can't precalc longer than 20-mers (would take over 4^20 = 2^40 bytes");
        if (prefix_length > k) throw std::runtime_error("Error: Precalc length is longer than k (" + std::to_string(prefix_length) + " > " + std::to_string(k) + ")");
        precalc_k = prefix_length;
        attach_device(true);
    }

    // ---- queries ---------------------------------------------------------------------------

    // SBWT.hh:384-415. Reads k bytes; only 'A','C','G','T' are valid (case-sensitive, SBWT.hh:427).
    int64_t search(const std::string& kmer) const { return search(kmer.c_str()); }
    int64_t search(const char* kmer) const {
        const int64_t off[2] = {0, k};
        int64_t out = -1;
        gpu_check(sbwt_gpu_query_host(get_session(k, 1), kmer, off, 1, SBWT_GPU_MODE_SEARCH, SBWT_GPU_CASE_EXACT, &out));
        return out;
    }

    // SBWT.hh:545-586. Throws if the index has no streaming support; empty result if len < k. Raw bytes as in the
    // reference: from-scratch searches are case-sensitive (SBWT.hh:427), a streaming step upper-cases its new
    // character (SBWT.hh:565) -- SBWT_GPU_CASE_API.
    std::vector<int64_t> streaming_search(const std::string& input) const { return streaming_search(input.c_str(), (int64_t)input.size()); }
    std::vector<int64_t> streaming_search(const char* input, int64_t len) const {
        if (suffix_group_starts.empty()) throw std::runtime_error("Error: streaming search support not built");
        const int64_t off[2] = {0, len};
        return streaming_search_batch(input, off, 1, SBWT_GPU_CASE_API);
    }

    // All k-mers of all reads; reads are ascii[offsets[i], offsets[i+1]). Results concatenated read after read.
    std::vector<int64_t> streaming_search_batch(const char* ascii, const int64_t* offsets, int64_t n_reads, int case_mode = SBWT_GPU_CASE_UPPER) const {
        if (suffix_group_starts.empty()) throw std::runtime_error("Error: streaming search support not built");
        return query_batch(ascii, offsets, n_reads, SBWT_GPU_MODE_STREAMING, case_mode);
    }
    std::vector<int64_t> search_batch(const char* ascii, const int64_t* offsets, int64_t n_reads, int case_mode = SBWT_GPU_CASE_UPPER) const {
        return query_batch(ascii, offsets, n_reads, SBWT_GPU_MODE_SEARCH, case_mode);
    }
    std::vector<int64_t> query_batch(const char* ascii, const int64_t* offsets, int64_t n_reads, int mode, int case_mode) const {
        std::vector<int64_t> out((size_t)sbwt_gpu_count_outputs(offsets, n_reads, k));
        query_batch_into(ascii, offsets, n_reads, mode, case_mode, out.data());
        return out;
    }
    void query_batch_into(const char* ascii, const int64_t* offsets, int64_t n_reads, int mode, int case_mode, int64_t* out) const {
        if (n_reads == 0) return;
        const int64_t bases = offsets[n_reads] - offsets[0];
        const int64_t cap_bases = std::min<int64_t>(std::max<int64_t>(bases, 1), (int64_t)96 << 20);
        int64_t longest = 0;
        for (int64_t i = 0; i < n_reads; i++) longest = std::max(longest, offsets[i + 1] - offsets[i]);
        gpu_check(sbwt_gpu_query_host(get_session(std::max(cap_bases, longest), std::min<int64_t>(n_reads, (int64_t)4 << 20)), ascii, offsets,
                                      n_reads, mode, case_mode, out));
    }

    // The same batch answered as the text `sbwt search` writes (print_vector, sbwt_search.cpp:21-43), formatted on the
    // device; pieces of the text reach `sink` in order. Returns the number of k-mers answered.
    int64_t query_batch_text(const char* ascii, const int64_t* offsets, int64_t n_reads, int mode, int case_mode, sbwt_gpu_text_sink sink,
                             void* user) const {
        if (n_reads == 0) return 0;
        if (mode == SBWT_GPU_MODE_STREAMING && suffix_group_starts.empty()) throw std::runtime_error("Error: streaming search support not built");
        const int64_t bases = offsets[n_reads] - offsets[0];
        const int64_t cap_bases = std::min<int64_t>(std::max<int64_t>(bases, 1), (int64_t)96 << 20);
        int64_t longest = 0, n_lookups = 0;
        for (int64_t i = 0; i < n_reads; i++) longest = std::max(longest, offsets[i + 1] - offsets[i]);
        gpu_check(sbwt_gpu_query_host_text(get_session(std::max(cap_bases, longest), std::min<int64_t>(n_reads, (int64_t)4 << 20)), ascii,
                                           offsets, n_reads, mode, case_mode, sink, user, &n_lookups));
        return n_lookups;
    }

    // SBWT.hh:418-437: extend the interval I by the characters of S -- one kernel launch for the whole string
    // (sbwt_gpu_update_interval_batch), not one device round trip per character.
    std::pair<int64_t, int64_t> update_sbwt_interval(const std::string& S, std::pair<int64_t, int64_t> I) const {
        return update_sbwt_interval(S.c_str(), (int64_t)S.size(), I);
    }
    std::pair<int64_t, int64_t> update_sbwt_interval(const char* S, int64_t S_length, std::pair<int64_t, int64_t> I) const {
        if (I.first == -1) return I;
        const int64_t off[2] = {0, S_length};
        int64_t l = I.first, r = I.second;
        gpu_check(sbwt_gpu_update_interval_batch(dev, S, off, 1, &l, &r));
        return {l, r};
    }
    // n (string, interval) pairs in one launch; strings are ascii[offsets[i], offsets[i+1])
    void update_sbwt_interval_batch(const char* ascii, const int64_t* offsets, int64_t n, int64_t* l, int64_t* r) const {
        gpu_check(sbwt_gpu_update_interval_batch(dev, ascii, offsets, n, l, r));
    }

    // SBWT.hh:369-381: follow the edge labelled c out of `node`; -1 if there is none (or c is not in ACGT).
    int64_t forward(int64_t node, char c) const {
        int64_t out = -1;
        gpu_check(sbwt_gpu_forward_batch(dev, &node, &c, 1, &out)); // (fails like the reference without streaming support)
        return out;
    }
    std::vector<int64_t> forward_batch(const std::vector<int64_t>& nodes, const std::string& chars) const {
        std::vector<int64_t> out(nodes.size());
        gpu_check(sbwt_gpu_forward_batch(dev, nodes.data(), chars.data(), (int64_t)nodes.size(), out.data()));
        return out;
    }

    // SBWT.hh:526-542: longest prefix of input that is found; returns ({l,r}, length matched).
    std::pair<std::pair<int64_t, int64_t>, int64_t> partial_search(const std::string& input) const { return partial_search(input.c_str(), (int64_t)input.size()); }
    std::pair<std::pair<int64_t, int64_t>, int64_t> partial_search(const char* input, int64_t len) const {
        const int64_t off[2] = {0, len};
        int64_t l = 0, r = 0, m = 0;
        gpu_check(sbwt_gpu_partial_search_batch(dev, input, off, 1, &l, &r, &m));
        return {{l, r}, m};
    }
    void partial_search_batch(const char* ascii, const int64_t* offsets, int64_t n, int64_t* l, int64_t* r, int64_t* matched) const {
        gpu_check(sbwt_gpu_partial_search_batch(dev, ascii, offsets, n, l, r, matched));
    }

    // SBWT.hh:701-725: the k-mer (label) of a node into buf (k bytes, '$'-padded on the left).
    void get_kmer(int64_t colex_rank, char* buf) const { gpu_check(sbwt_gpu_get_kmer_batch(dev, &colex_rank, 1, buf)); }
    void get_kmer_batch(const int64_t* colex_ranks, int64_t n, char* buf) const { gpu_check(sbwt_gpu_get_kmer_batch(dev, colex_ranks, n, buf)); }

    // SBWT.hh:750-773: the subsets as text, written by the device.
    template <typename out_stream_t>
    void ascii_export_sets(out_stream_t& out) const {
        std::vector<char> text((size_t)(4 * n_nodes + 1));
        int64_t n = 0;
        gpu_check(sbwt_gpu_ascii_export_sets(dev, text.data(), (int64_t)text.size(), &n));
        out.write(text.data(), n);
    }

    // ---- serialization (SBWT.hh:463-522); the variant string is written/read by the caller ----

    int64_t serialize(std::ostream& os) const {
        using namespace sbwt_b200::detail;
        int64_t written = wr_string(os, "v0.1");
        written += subset_rank.serialize(os);
        written += wr_words(os, suffix_group_starts.empty() ? 0 : n_nodes, suffix_group_starts);
        wr_i64(os, 32); wr(os, C.data(), 32); written += 40;
        wr_i64(os, (int64_t)kmer_prefix_precalc.size() * 16);
        wr(os, kmer_prefix_precalc.data(), kmer_prefix_precalc.size() * 16);
        written += 8 + (int64_t)kmer_prefix_precalc.size() * 16;
        wr_i64(os, precalc_k); wr_i64(os, n_nodes); wr_i64(os, n_kmers); wr_i64(os, k);
        return written + 32;
    }
    int64_t serialize(const std::string& filename) const {
        std::ofstream out(filename, std::ios::binary);
        if (!out.good()) throw std::runtime_error("Error opening file: " + filename);
        return serialize(out);
    }

    void load(std::istream& is) {
        using namespace sbwt_b200::detail;
        if (rd_string(is) != "v0.1")
            throw std::runtime_error("Error: Corrupt index file, or the index was constructed with an incompatible version of SBWT.");
        subset_rank.load(is);
        int64_t sgs_bits, nbytes;
        suffix_group_starts = rd_words(is, &sgs_bits);
        if (rd_i64(is) != 32) throw std::runtime_error("Error: Corrupt index file (C array).");
        C.assign(4, 0);
        rd(is, C.data(), 32);
        nbytes = rd_i64(is);
        if (nbytes < 0 || (nbytes & 15)) throw std::runtime_error("Error: Corrupt index file (precalc table).");
        kmer_prefix_precalc.resize((size_t)nbytes / 16);
        rd(is, kmer_prefix_precalc.data(), (size_t)nbytes);
        precalc_k = rd_i64(is);
        n_nodes = rd_i64(is);
        n_kmers = rd_i64(is);
        k = rd_i64(is);
        if (subset_rank.n_bits != n_nodes || (sgs_bits != 0 && sgs_bits != n_nodes) ||
            (int64_t)kmer_prefix_precalc.size() != (precalc_k ? (int64_t)1 << (2 * precalc_k) : 0))
            throw std::runtime_error("Error: Corrupt index file (inconsistent sizes).");
        attach_device(false);
    }
    void load(const std::string& filename) {
        std::ifstream in(filename, std::ios::binary);
        if (!in.good()) throw std::runtime_error("Error opening file: " + filename);
        load(in);
    }
};

typedef SBWT<GpuSubsetMatrixRank> plain_matrix_sbwt_t; // variants.hh:19

} // namespace sbwt
