// sbwt_file.hpp -- host-side reader of the reference's serialized plain-matrix index.
//
// Replaces, for the plain-matrix variant only, the variant-string read of
// src/CLI/sbwt_search.cpp:194-199, SBWT::load (include/sbwt/SBWT.hh:501-516),
// SubsetMatrixRank::load (include/sbwt/SubsetMatrixRank.hh:102-125), load_string
// (src/globals.cpp:56-62), load_std_vector (SBWT.hh:451-459) and sdsl
// int_vector::load (sdsl-lite/include/sdsl/int_vector.hpp:1614-1628).
//
// Little-endian layout, parsed to EOF:
//   [i64 len]["plain-matrix"] [i64 len]["v0.1"]
//   4 x bit_vector { u64 nbits, ceil(nbits/64) x u64 }            A, C, G, T
//   4 x rank_support_v5 { u64 nbits = 64*W, W x u64 }             skipped: the device index
//                                                                 carries its own directory
//   suffix_group_starts bit_vector { u64 nbits (0 = absent), words }
//   [i64 32][4 x i64 C]  [i64 16*4^p][pairs (l,r)]  i64 p  i64 n_nodes  i64 n_kmers  i64 k
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace sbwt_b200 {

struct PlainMatrixFile {
    int64_t n_nodes = 0, n_kmers = 0, k = 0, precalc_k = 0;
    int64_t C[4] = {0, 0, 0, 0};
    std::vector<uint64_t> bits[4];            // ceil(n_nodes/64) words each
    std::vector<uint64_t> suffix_group_starts; // empty when the index has no streaming support
    std::vector<int64_t> precalc;              // 2 * 4^p values
};

namespace detail {
struct File {
    FILE* f;
    explicit File(const std::string& path) : f(std::fopen(path.c_str(), "rb")) {
        if (!f) throw std::runtime_error("Error opening file: " + path); // throwing_streams.hh semantics
    }
    ~File() { if (f) std::fclose(f); }
    void read(void* dst, size_t n) {
        if (n && std::fread(dst, 1, n, f) != n) throw std::runtime_error("Error: Corrupt index file (truncated).");
    }
    int64_t i64() { int64_t x; read(&x, 8); return x; }
    std::string str() {
        int64_t n = i64();
        if (n < 0 || n > 4096) throw std::runtime_error("Error: Corrupt index file (bad string length).");
        std::string s((size_t)n, '\0');
        read(&s[0], (size_t)n);
        return s;
    }
    std::vector<uint64_t> bitvector(int64_t* nbits_out) {
        int64_t nbits = i64();
        if (nbits < 0) throw std::runtime_error("Error: Corrupt index file (negative bit vector length).");
        std::vector<uint64_t> w((size_t)((nbits + 63) / 64));
        read(w.data(), w.size() * 8);
        *nbits_out = nbits;
        return w;
    }
    void skip_words() { // an int_vector<64>: [u64 nbits][nbits/64 words]
        int64_t nbits = i64();
        if (nbits < 0 || (nbits & 63)) throw std::runtime_error("Error: Corrupt index file (bad rank support).");
        if (std::fseek(f, (long)(nbits / 8), SEEK_CUR)) throw std::runtime_error("Error: Corrupt index file (truncated).");
    }
};
} // namespace detail

inline PlainMatrixFile load_plain_matrix_file(const std::string& path) {
    detail::File in(path);
    std::string variant = in.str();
    if (variant != "plain-matrix")
        throw std::runtime_error("Error loading index from file: only the plain-matrix variant is supported on the GPU path (file has '" + variant + "')");
    if (in.str() != "v0.1") // SBWT_VERSION, SBWT.hh:27
        throw std::runtime_error("Error: Corrupt index file, or the index was constructed with an incompatible version of SBWT.");
    PlainMatrixFile F;
    int64_t nbits[4];
    for (int c = 0; c < 4; c++) F.bits[c] = in.bitvector(&nbits[c]);
    for (int c = 0; c < 4; c++) in.skip_words();
    int64_t sgs_bits;
    F.suffix_group_starts = in.bitvector(&sgs_bits);
    if (in.i64() != 32) throw std::runtime_error("Error: Corrupt index file (C array).");
    in.read(F.C, 32);
    int64_t pbytes = in.i64();
    if (pbytes < 0 || (pbytes & 15)) throw std::runtime_error("Error: Corrupt index file (precalc table).");
    F.precalc.resize((size_t)(pbytes / 8));
    in.read(F.precalc.data(), (size_t)pbytes);
    F.precalc_k = in.i64();
    F.n_nodes = in.i64();
    F.n_kmers = in.i64();
    F.k = in.i64();
    if (std::fgetc(in.f) != EOF) throw std::runtime_error("Error: Corrupt index file (trailing bytes).");
    for (int c = 0; c < 4; c++)
        if (nbits[c] != F.n_nodes) throw std::runtime_error("Error: Corrupt index file (bit vector length != number of subsets).");
    if (sgs_bits != 0 && sgs_bits != F.n_nodes) throw std::runtime_error("Error: Corrupt index file (streaming support length).");
    if (F.precalc_k < 0 || F.precalc_k > 20 || F.precalc_k > F.k ||
        (int64_t)F.precalc.size() != (F.precalc_k ? (int64_t)2 << (2 * F.precalc_k) : 0))
        throw std::runtime_error("Error: Corrupt index file (precalc table size).");
    if (F.k < 1 || F.n_nodes < 1) throw std::runtime_error("Error: Corrupt index file (k / number of subsets).");
    return F;
}

} // namespace sbwt_b200
