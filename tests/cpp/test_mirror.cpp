// Exercises the C++ mirror of the reference's template surface (sbwt_b200/csrc/SBWT.hh) the way the
// reference's own tests and api_examples/api_example.cpp use sbwt::SBWT<>: load, search,
// streaming_search, update_sbwt_interval, partial_search, accessors, serialize, and the
// bit-vector constructor with do_kmer_prefix_precalc. Prints "key value..." lines that
// tests/test_cli.py compares with the oracle.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../sbwt_b200/csrc/SBWT.hh"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string index_path = argv[1], out_path = argv[2], read = argv[3];
    try {
        sbwt::plain_matrix_sbwt_t idx;
        {
            std::ifstream in(index_path, std::ios::binary);
            std::string variant = sbwt_b200::load_variant_string(in);
            if (variant != "plain-matrix") return 3;
            idx.load(in);
        }
        const int64_t k = idx.get_k();
        std::cout << "header " << idx.number_of_subsets() << " " << idx.number_of_kmers() << " " << k << " " << idx.get_precalc_k() << " "
                  << idx.has_streaming_query_support();
        for (int64_t c : idx.get_C_array()) std::cout << " " << c;
        std::cout << "\n";
        std::cout << "streaming";
        for (int64_t x : idx.streaming_search(read)) std::cout << " " << x;
        std::cout << "\nsearch";
        for (size_t i = 0; i + k <= read.size(); i++) std::cout << " " << idx.search(read.substr(i, k));
        std::cout << "\n";
        // update_sbwt_interval over the first k characters from the full interval == search of the first k-mer
        auto I = idx.update_sbwt_interval(read.substr(0, k), {0, idx.number_of_subsets() - 1});
        std::cout << "interval " << I.first << " " << I.second << "\n";
        auto P = idx.partial_search(read);
        std::cout << "partial " << P.first.first << " " << P.first.second << " " << P.second << "\n";
        // forward (SBWT.hh:369-381) along the read reproduces the streaming answers: ans[i+1] == forward(ans[i], read[i+k])
        {
            const std::vector<int64_t> ans = idx.streaming_search(read);
            std::cout << "forward";
            for (size_t i = 0; i + 1 < ans.size() && i < 40; i++) std::cout << " " << (ans[i] >= 0 ? idx.forward(ans[i], read[i + k]) : -2);
            std::cout << " " << idx.forward(ans[0], 'N') << "\n";
        }
        std::cout << "rank " << idx.get_subset_rank_structure().rank(idx.number_of_subsets(), 'A') << " "
                  << idx.get_subset_rank_structure().rank(idx.number_of_subsets() / 2, 'G') << " "
                  << idx.get_subset_rank_structure().rank(5, 'N') << "\n";
        // serialize: variant string + SBWT::serialize must reproduce the file
        {
            std::ofstream out(out_path, std::ios::binary);
            sbwt_b200::detail::wr_string(out, "plain-matrix");
            idx.serialize(out);
        }
        // the constructor from bit vectors computes C and the precalc table itself (on the device)
        const auto& R = idx.get_subset_rank_structure();
        sbwt::plain_matrix_sbwt_t built(R.A_bits, R.C_bits, R.G_bits, R.T_bits, idx.get_streaming_support(), idx.number_of_subsets(), k,
                                        idx.number_of_kmers(), idx.get_precalc_k());
        std::cout << "rebuilt_same_C " << (built.get_C_array() == idx.get_C_array()) << "\n";
        std::cout << "rebuilt_same_precalc " << (built.get_precalc() == idx.get_precalc()) << "\n";
        std::cout << "rebuilt_streaming";
        for (int64_t x : built.streaming_search(read)) std::cout << " " << x;
        std::cout << "\n";
        // errors keep the reference's style
        try {
            sbwt::plain_matrix_sbwt_t bad;
            std::istringstream s(std::string("\x04\0\0\0\0\0\0\0v9.9", 12));
            bad.load(s);
            std::cout << "version_error none\n";
        } catch (const std::runtime_error& e) {
            std::cout << "version_error " << e.what() << "\n";
        }
    } catch (const std::exception& e) {
        std::cout << "exception " << e.what() << "\n";
        return 1;
    }
    return 0;
}
