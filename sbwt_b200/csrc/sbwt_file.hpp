// sbwt_file.hpp -- host-side reader / writer of the reference's serialized plain-matrix index.
//
// Replaces, for the plain-matrix variant only, the variant-string read of
// src/CLI/sbwt_search.cpp:194-199, SBWT::load / serialize (include/sbwt/SBWT.hh:463-516),
// SubsetMatrixRank::load / serialize (include/sbwt/SubsetMatrixRank.hh:86-125), load_string /
// serialize_string (src/globals.cpp:49-62), load_std_vector (SBWT.hh:451-459) and sdsl
// int_vector::load (sdsl-lite/include/sdsl/int_vector.hpp:1614-1628).
//
// Little-endian layout, parsed to EOF:
//   [i64 len]["plain-matrix"]                         (written by the CLI, not by SBWT::serialize)
//   [i64 len]["v0.1"]
//   4 x bit_vector { u64 nbits, ceil(nbits/64) x u64 }            A, C, G, T
//   4 x rank_support_v5 { u64 nbits = 64*W, W x u64 }             kept verbatim for serialize();
//                                                                 the device index has its own directory
//   suffix_group_starts bit_vector { u64 nbits (0 = absent), words }
//   [i64 32][4 x i64 C]  [i64 16*4^p][pairs (l,r)]  i64 p  i64 n_nodes  i64 n_kmers  i64 k
#pragma once

#include <cstdint>
#include <fstream>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sbwt_b200 {

struct PlainMatrixFile {
    int64_t n_nodes = 0, n_kmers = 0, k = 0, precalc_k = 0;
    int64_t C[4] = {0, 0, 0, 0};
    std::vector<uint64_t> bits[4];             // ceil(n_nodes/64) words each
    std::vector<uint64_t> rank_support[4];     // the file's rank_support_v5 words (not used for queries)
    std::vector<uint64_t> suffix_group_starts; // empty when the index has no streaming support
    std::vector<int64_t> precalc;              // 2 * 4^p values
};

namespace detail {
inline void rd(std::istream& in, void* dst, size_t n) {
    in.read(static_cast<char*>(dst), (std::streamsize)n);
    if ((size_t)in.gcount() != n) throw std::runtime_error("Error: Corrupt index file (truncated).");
}
inline int64_t rd_i64(std::istream& in) { int64_t x; rd(in, &x, 8); return x; }
inline std::string rd_string(std::istream& in) {
    int64_t n = rd_i64(in);
    if (n < 0 || n > 4096) throw std::runtime_error("Error: Corrupt index file (bad string length).");
    std::string s((size_t)n, '\0');
    rd(in, &s[0], (size_t)n);
    return s;
}
inline std::vector<uint64_t> rd_words(std::istream& in, int64_t* nbits_out) {
    int64_t nbits = rd_i64(in);
    if (nbits < 0) throw std::runtime_error("Error: Corrupt index file (negative vector length).");
    std::vector<uint64_t> w((size_t)((nbits + 63) / 64));
    rd(in, w.data(), w.size() * 8);
    *nbits_out = nbits;
    return w;
}
inline void wr(std::ostream& out, const void* p, size_t n) { out.write(static_cast<const char*>(p), (std::streamsize)n); }
inline void wr_i64(std::ostream& out, int64_t x) { wr(out, &x, 8); }
inline int64_t wr_string(std::ostream& out, const std::string& s) { wr_i64(out, (int64_t)s.size()); wr(out, s.data(), s.size()); return 8 + (int64_t)s.size(); }
inline int64_t wr_words(std::ostream& out, int64_t nbits, const std::vector<uint64_t>& w) { wr_i64(out, nbits); wr(out, w.data(), w.size() * 8); return 8 + (int64_t)w.size() * 8; }
} // namespace detail

// Reads the variant string that `sbwt build` puts in front of the index (sbwt_search.cpp:194-199).
inline std::string load_variant_string(std::istream& in) { return detail::rd_string(in); }

// SBWT::load(istream&): everything after the variant string. `to_eof` additionally requires the
// stream to end right after the index, as it does in a .sbwt file.
inline PlainMatrixFile load_plain_matrix(std::istream& in, bool to_eof = true) {
    using namespace detail;
    if (rd_string(in) != "v0.1") // SBWT_VERSION, SBWT.hh:27
        throw std::runtime_error("Error: Corrupt index file, or the index was constructed with an incompatible version of SBWT.");
    PlainMatrixFile F;
    int64_t nbits[4], rs_bits, sgs_bits;
    for (int c = 0; c < 4; c++) F.bits[c] = rd_words(in, &nbits[c]);
    for (int c = 0; c < 4; c++) {
        F.rank_support[c] = rd_words(in, &rs_bits);
        if (rs_bits & 63) throw std::runtime_error("Error: Corrupt index file (bad rank support).");
    }
    F.suffix_group_starts = rd_words(in, &sgs_bits);
    if (rd_i64(in) != 32) throw std::runtime_error("Error: Corrupt index file (C array).");
    rd(in, F.C, 32);
    int64_t pbytes = rd_i64(in);
    if (pbytes < 0 || (pbytes & 15)) throw std::runtime_error("Error: Corrupt index file (precalc table).");
    F.precalc.resize((size_t)(pbytes / 8));
    rd(in, F.precalc.data(), (size_t)pbytes);
    F.precalc_k = rd_i64(in);
    F.n_nodes = rd_i64(in);
    F.n_kmers = rd_i64(in);
    F.k = rd_i64(in);
    if (to_eof && in.peek() != std::istream::traits_type::eof()) throw std::runtime_error("Error: Corrupt index file (trailing bytes).");
    for (int c = 0; c < 4; c++)
        if (nbits[c] != F.n_nodes) throw std::runtime_error("Error: Corrupt index file (bit vector length != number of subsets).");
    if (sgs_bits != 0 && sgs_bits != F.n_nodes) throw std::runtime_error("Error: Corrupt index file (streaming support length).");
    if (F.precalc_k < 0 || F.precalc_k > 20 || F.precalc_k > F.k ||
        (int64_t)F.precalc.size() != (F.precalc_k ? (int64_t)2 << (2 * F.precalc_k) : 0))
        throw std::runtime_error("Error: Corrupt index file (precalc table size).");
    if (F.k < 1 || F.n_nodes < 1) throw std::runtime_error("Error: Corrupt index file (k / number of subsets).");
    return F;
}

// A whole .sbwt file: variant string + index.
inline PlainMatrixFile load_plain_matrix_file(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) throw std::runtime_error("Error opening file: " + path); // throwing_streams.hh semantics
    std::string variant = load_variant_string(in);
    if (variant != "plain-matrix")
        throw std::runtime_error("Error loading index from file: only the plain-matrix variant is supported on the GPU path (file has '" + variant + "')");
    return load_plain_matrix(in, true);
}

// SBWT::serialize(ostream&) (SBWT.hh:463-493): everything after the variant string. Returns bytes written.
inline int64_t serialize_plain_matrix(const PlainMatrixFile& F, std::ostream& out) {
    using namespace detail;
    int64_t n = wr_string(out, "v0.1");
    for (int c = 0; c < 4; c++) n += wr_words(out, F.n_nodes, F.bits[c]);
    for (int c = 0; c < 4; c++) n += wr_words(out, (int64_t)F.rank_support[c].size() * 64, F.rank_support[c]);
    n += wr_words(out, F.suffix_group_starts.empty() ? 0 : F.n_nodes, F.suffix_group_starts);
    wr_i64(out, 32); wr(out, F.C, 32); n += 40;
    wr_i64(out, (int64_t)F.precalc.size() * 8); wr(out, F.precalc.data(), F.precalc.size() * 8); n += 8 + (int64_t)F.precalc.size() * 8;
    wr_i64(out, F.precalc_k); wr_i64(out, F.n_nodes); wr_i64(out, F.n_kmers); wr_i64(out, F.k); n += 32;
    return n;
}

// The words sdsl's rank_support_v5<1,1> constructor produces (rank_support_v5.hpp:65-109): per
// 2048-bit superblock an absolute count and the counts of its first 6,12,...,30 words in 12-bit
// fields at shifts 48..0; a field exists only if the vector has at least that many words.
inline std::vector<uint64_t> rank_support_v5_words(const std::vector<uint64_t>& bits, int64_t nbits) {
    if (nbits == 0) return std::vector<uint64_t>(2, 0);
    const uint64_t W = (uint64_t)(nbits + 63) / 64, nsb = W / 32 + 1;
    std::vector<uint64_t> bb(nsb * 2, 0);
    uint64_t total = 0;
    for (uint64_t s = 0; s < nsb; s++) {
        bb[2 * s] = total;
        uint64_t second = 0, sum = 0;
        for (uint64_t j = 0; j < 32; j++) {
            const uint64_t wi = 32 * s + j;
            if (j && j % 6 == 0 && W >= wi) second |= sum << (60 - 12 * (j / 6));
            if (wi < W) sum += (uint64_t)__builtin_popcountll(bits[wi]);
        }
        bb[2 * s + 1] = second;
        total += sum;
    }
    return bb;
}

} // namespace sbwt_b200
