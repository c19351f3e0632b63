// parse_time.cpp -- FASTA/FASTQ parse rate of ParallelFastxReader alone (no device): parse_time <file> <threads>
// g++ -O2 -std=c++17 -DFASTX_HPP=\"../../sbwt_b200/csrc/fastx.hpp\" -o parse_time parse_time.cpp -lz -lpthread
#include <chrono>
#include <cstdio>
#include <vector>
#include FASTX_HPP
int main(int argc, char** argv) {
    const int threads = atoi(argv[2]);
    auto t0 = std::chrono::steady_clock::now();
    sbwt_b200::ParallelFastxReader r(argv[1], threads);
    std::vector<char> ascii; std::vector<int64_t> off;
    int64_t reads = 0, bases = 0, batches = 0;
    for (;;) { const int64_t n = r.next_batch(64 << 20, 1 << 30, ascii, off); if (!n) break; reads += n; bases += ascii.size(); batches++; }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %lld reads %lld bases %lld batches in %.3f s = %.2f GB/s of sequence\n", threads, (long long)reads, (long long)bases, (long long)batches, s, bases / s / 1e9);
}
