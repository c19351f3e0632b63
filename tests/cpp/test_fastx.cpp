// Differential test of the two batch readers of sbwt_b200/csrc/fastx.hpp: ParallelFastxReader must deliver the
// batches of FastxReader byte for byte (and fail with the same message) -- the serial one restates
// seq_io::Reader::get_next_read_to_buffer (SeqIO.hh:255-360).
// usage: test_fastx <file> <max_bases> <max_reads> <threads>   -> prints "<n_batches> <n_reads> <n_bases> <fnv hash>" or "ERROR <what>"
// for both readers, one line each.
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../sbwt_b200/csrc/fastx.hpp"

template <typename R>
static std::string run(R& reader, int64_t max_bases, int64_t max_reads) {
    std::vector<char> ascii;
    std::vector<int64_t> off;
    uint64_t h = 1469598103934665603ull;
    int64_t batches = 0, reads = 0, bases = 0;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    try {
        for (;;) {
            const int64_t n = reader.next_batch(max_bases, max_reads, ascii, off);
            if (n == 0) break;
            batches++;
            reads += n;
            bases += (int64_t)ascii.size();
            mix((uint64_t)n);
            for (int64_t x : off) mix((uint64_t)x);
            for (char c : ascii) mix((unsigned char)c);
        }
    } catch (const std::exception& e) {
        return std::string("ERROR ") + e.what() + " after " + std::to_string(reads) + " reads";
    }
    return std::to_string(batches) + " " + std::to_string(reads) + " " + std::to_string(bases) + " " + std::to_string(h);
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const std::string file = argv[1];
    const int64_t mb = atoll(argv[2]), mr = atoll(argv[3]);
    const int threads = atoi(argv[4]);
    std::string a, b;
    try {
        sbwt_b200::FastxReader r(file);
        a = run(r, mb, mr);
    } catch (const std::exception& e) { a = std::string("ERROR ") + e.what(); }
    try {
        sbwt_b200::ParallelFastxReader r(file, threads);
        b = run(r, mb, mr);
    } catch (const std::exception& e) { b = std::string("ERROR ") + e.what(); }
    std::cout << a << "\n" << b << "\n";
    return 0;
}
