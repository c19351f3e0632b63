// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin command-line driver around the UNMODIFIED reference classes, compiled
// from the sources where they lie under /root/reference (see oracle/Makefile;
// the reference's own cmake build is not used). Everything on the query path
// -- SBWT<>::load / search / streaming_search (include/sbwt/SBWT.hh),
// SubsetMatrixRank (include/sbwt/SubsetMatrixRank.hh), sdsl bit_vector and
// rank_support_v5, seq_io::Reader (SeqIO/include/SeqIO/SeqIO.hh) -- is the
// reference's code. Only the per-read loop and print_vector of
// src/CLI/sbwt_search.cpp:21-91 are restated here, because that translation
// unit also instantiates the nine out-of-scope variants, which need a
// cmake-generated header (divsufsort.h).
//
// Commands:
//   search      -i index -q reads.(fa|fq)[.gz] -o out.txt [-z]   like `sbwt search` for one file pair
//   timed       -i index -q reads -t threads                     CPU baseline; prints one JSON line
//   build-inmem -i in.fna -o out.sbwt -k K [-p P] [--no-streaming-support] [--add-reverse-complements]
//   ranks       -i index   < "pos char" lines                    prints rank_c(pos) per line
//   dump        -i index                                         prints the header numbers
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "SBWT.hh"
#include "SubsetMatrixRank.hh"
#include "SeqIO/SeqIO.hh"
#include "SeqIO/buffered_streams.hh"

using namespace sbwt;
using std::string;
using std::vector;

typedef SBWT<SubsetMatrixRank<sdsl::bit_vector, sdsl::rank_support_v5<>>> plain_matrix_t; // variants.hh:19

static string arg(int argc, char** argv, const string& name, const string& dflt = "") {
    for (int i = 0; i + 1 < argc; i++)
        if (name == argv[i]) return argv[i + 1];
    return dflt;
}
static bool flag(int argc, char** argv, const string& name) {
    for (int i = 0; i < argc; i++)
        if (name == argv[i]) return true;
    return false;
}

static void load_index(const string& path, plain_matrix_t& idx) {
    throwing_ifstream in(path, ios::binary);
    string variant = load_string(in.stream); // sbwt_search.cpp:194
    if (variant != "plain-matrix") throw std::runtime_error("not a plain-matrix index: " + variant);
    idx.load(in.stream);
}

// print_vector, sbwt_search.cpp:21-43 (restated).
template <typename writer_t>
static void print_values(const vector<int64_t>& v, writer_t& out) {
    char buffer[32];
    char newline = '\n';
    for (int64_t x : v) {
        int64_t i = 0;
        if (x == -1) {
            buffer[0] = '1';
            buffer[1] = '-';
            i = 2;
        } else {
            while (x > 0) {
                buffer[i++] = '0' + (x % 10);
                x /= 10;
            }
        }
        std::reverse(buffer, buffer + i);
        buffer[i] = ' ';
        out.write(buffer, i + 1);
    }
    out.write(&newline, 1);
}

// run_queries_streaming / run_queries_not_streaming, sbwt_search.cpp:45-91 (restated loop,
// reference classes underneath).
template <typename reader_t, typename writer_t>
static int64_t run_file(const plain_matrix_t& idx, const string& in, const string& out) {
    reader_t reader(in);
    writer_t writer(out);
    int64_t nq = 0, k = idx.get_k();
    bool streaming = idx.has_streaming_query_support();
    vector<int64_t> buf;
    while (true) {
        int64_t len = reader.get_next_read_to_buffer();
        if (len == 0) break;
        if (streaming) {
            buf = idx.streaming_search(reader.read_buf, len);
        } else {
            buf.clear();
            for (int64_t i = 0; i < len - k + 1; i++) buf.push_back(idx.search(reader.read_buf + i));
        }
        nq += buf.size();
        print_values(buf, writer);
    }
    return nq;
}

static int cmd_search(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    string q = arg(argc, argv, "-q"), o = arg(argc, argv, "-o");
    bool gz_in = seq_io::figure_out_file_format(q).gzipped, gz_out = flag(argc, argv, "-z");
    typedef seq_io::Reader<seq_io::Buffered_ifstream<seq_io::zstr::ifstream>> in_gzip;
    typedef seq_io::Reader<seq_io::Buffered_ifstream<std::ifstream>> in_plain;
    typedef seq_io::Buffered_ofstream<seq_io::zstr::ofstream> out_gzip;
    typedef seq_io::Buffered_ofstream<std::ofstream> out_plain;
    int64_t nq;
    if (gz_in && gz_out) nq = run_file<in_gzip, out_gzip>(idx, q, o);
    else if (gz_in) nq = run_file<in_gzip, out_plain>(idx, q, o);
    else if (gz_out) nq = run_file<in_plain, out_gzip>(idx, q, o);
    else nq = run_file<in_plain, out_plain>(idx, q, o);
    std::cerr << "queries: " << nq << std::endl;
    return 0;
}

// CPU baseline: reads are parsed first, then T threads each answer a contiguous
// slice with the reference's const query methods; only the query loop is timed
// (mirrors the "us/query (excluding I/O etc)" figure of sbwt_search.cpp:63,89,
// without the two chrono calls per k-mer of the non-streaming loop).
static int cmd_timed(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    int T = std::stoi(arg(argc, argv, "-t", "1"));
    int reps = std::stoi(arg(argc, argv, "-r", "1"));
    vector<string> reads;
    {
        seq_io::Reader<> reader(arg(argc, argv, "-q"));
        while (true) {
            int64_t len = reader.get_next_read_to_buffer();
            if (len == 0) break;
            reads.emplace_back(reader.read_buf, len);
        }
    }
    int64_t k = idx.get_k();
    bool streaming = idx.has_streaming_query_support();
    vector<int64_t> counts(T, 0), hits(T, 0), sums(T, 0);
    vector<uint64_t> wsums(T, 0); // sum of (1-based position inside the thread's slice) x value, modulo 2^64
    double best = 1e300;
    for (int rep = 0; rep < reps; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        vector<std::thread> th;
        for (int t = 0; t < T; t++) {
            th.emplace_back([&, t]() {
                size_t a = reads.size() * t / T, b = reads.size() * (t + 1) / T;
                int64_t n = 0, h = 0, s = 0;
                uint64_t ws = 0;
                for (size_t i = a; i < b; i++) {
                    const string& R = reads[i];
                    if (streaming) {
                        vector<int64_t> v = idx.streaming_search(R.c_str(), R.size());
                        for (int64_t x : v) { n++; h += x >= 0; s += x; ws += (uint64_t)n * (uint64_t)x; }
                    } else {
                        for (int64_t j = 0; j < (int64_t)R.size() - k + 1; j++) {
                            int64_t x = idx.search(R.c_str() + j);
                            n++; h += x >= 0; s += x; ws += (uint64_t)n * (uint64_t)x;
                        }
                    }
                }
                counts[t] = n; hits[t] = h; sums[t] = s; wsums[t] = ws;
            });
        }
        for (auto& x : th) x.join();
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::min(best, sec);
    }
    // position-weighted checksum of the whole output: sum over results of (1-based index) x value, modulo 2^64 -- a
    // permutation or a shift of the results changes it (the plain sum would not)
    int64_t n = 0, h = 0, s = 0;
    uint64_t ws = 0;
    for (int t = 0; t < T; t++) { ws += wsums[t] + (uint64_t)n * (uint64_t)sums[t]; n += counts[t]; h += hits[t]; s += sums[t]; }
    std::cout << "{\"lookups\": " << n << ", \"hits\": " << h << ", \"checksum\": " << s << ", \"weighted_checksum\": " << ws << ", \"seconds\": " << best
              << ", \"threads\": " << T << ", \"reads\": " << reads.size() << ", \"streaming\": " << (streaming ? "true" : "false")
              << "}" << std::endl;
    return 0;
}

static int cmd_build_inmem(int argc, char** argv) {
    int64_t k = std::stoll(arg(argc, argv, "-k"));
    int64_t p = std::stoll(arg(argc, argv, "-p", "8"));
    if (p > k) p = k; // sbwt_build.cpp:101-105
    bool streaming = !flag(argc, argv, "--no-streaming-support");
    bool rc = flag(argc, argv, "--add-reverse-complements");
    vector<string> seqs;
    {
        seq_io::Reader<> reader(arg(argc, argv, "-i"));
        if (rc) reader.enable_reverse_complements();
        while (true) {
            int64_t len = reader.get_next_read_to_buffer();
            if (len == 0) break;
            seqs.emplace_back(reader.read_buf, len);
        }
    }
    plain_matrix_t idx;
    build_nodeboss_in_memory(seqs, idx, k, streaming); // NodeBOSSInMemoryConstructor.hh:188-221
    idx.do_kmer_prefix_precalc(p);                      // sbwt_build.cpp:157
    throwing_ofstream out(arg(argc, argv, "-o"), ios::binary);
    serialize_string("plain-matrix", out.stream);       // sbwt_build.cpp:142
    idx.serialize(out.stream);
    std::cerr << "n_kmers " << idx.number_of_kmers() << " n_subsets " << idx.number_of_subsets() << std::endl;
    return 0;
}

static int cmd_ranks(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    int64_t pos;
    char c;
    while (std::cin >> pos >> c) std::cout << idx.get_subset_rank_structure().rank(pos, c) << "\n";
    return 0;
}

// The other read-only queries of SBWT.hh, for the golden fixtures of tests/golden/*/f4.txt:
//   partial  -q reads        one line per read: "l r matched"                     (SBWT::partial_search, SBWT.hh:526-537)
//   forward  (stdin: "node c" per line)  one result per line                        (SBWT::forward, SBWT.hh:369-381)
//   getkmer  (stdin: colex rank per line) one k-mer label per line                  (SBWT::get_kmer, SBWT.hh:701-725)
//   export   -o file         SBWT::ascii_export_sets (SBWT.hh:750-773)
//   api      [--no-streaming] (stdin: one raw sequence per line) streaming_search / search on the bytes as given
static int cmd_partial(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    seq_io::Reader<> reader(arg(argc, argv, "-q"));
    while (true) {
        int64_t len = reader.get_next_read_to_buffer();
        if (len == 0) break;
        auto res = idx.partial_search(reader.read_buf, len);
        std::cout << res.first.first << " " << res.first.second << " " << res.second << "\n";
    }
    return 0;
}

// The direct API on the caller's raw bytes (no SeqIO upper-casing in front): one sequence per line on stdin,
// SBWT::streaming_search(const char*, len) (SBWT.hh:545-581) -- or the search() loop with --no-streaming -- per line,
// the result vector printed as numbers separated by blanks. Pins the mixed-case behaviour: a from-scratch search
// takes the bytes as they are (SBWT.hh:427), a streaming step upper-cases its new character (SBWT.hh:565).
static int cmd_api(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    const bool streaming = !flag(argc, argv, "--no-streaming");
    const int64_t k = idx.get_k();
    string line;
    while (std::getline(std::cin, line)) {
        vector<int64_t> res;
        if (streaming) res = idx.streaming_search(line.c_str(), (int64_t)line.size());
        else
            for (int64_t i = 0; i + k <= (int64_t)line.size(); i++) res.push_back(idx.search(line.c_str() + i));
        for (size_t i = 0; i < res.size(); i++) std::cout << (i ? " " : "") << res[i];
        std::cout << "\n";
    }
    return 0;
}

static int cmd_forward(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    int64_t node;
    char c;
    while (std::cin >> node >> c) std::cout << idx.forward(node, c) << "\n";
    return 0;
}

static int cmd_getkmer(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    int64_t rank;
    string buf(idx.get_k(), ' ');
    while (std::cin >> rank) {
        idx.get_kmer(rank, &buf[0]);
        std::cout << buf << "\n";
    }
    return 0;
}

static int cmd_export(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    throwing_ofstream out(arg(argc, argv, "-o"), ios::binary);
    idx.ascii_export_sets(out.stream);
    return 0;
}

static int cmd_dump(int argc, char** argv) {
    plain_matrix_t idx;
    load_index(arg(argc, argv, "-i"), idx);
    std::cout << "n_nodes " << idx.number_of_subsets() << " n_kmers " << idx.number_of_kmers() << " k " << idx.get_k()
              << " precalc_k " << idx.get_precalc_k() << " streaming " << idx.has_streaming_query_support() << " C";
    for (int64_t x : idx.get_C_array()) std::cout << " " << x;
    std::cout << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    set_log_level(LogLevel::OFF);
    if (argc < 2) {
        std::cerr << "usage: sbwt_ref {search|timed|build-inmem|ranks|dump|partial|forward|getkmer|export} ..." << std::endl;
        return 1;
    }
    string cmd = argv[1];
    try {
        if (cmd == "search") return cmd_search(argc, argv);
        if (cmd == "timed") return cmd_timed(argc, argv);
        if (cmd == "build-inmem") return cmd_build_inmem(argc, argv);
        if (cmd == "ranks") return cmd_ranks(argc, argv);
        if (cmd == "dump") return cmd_dump(argc, argv);
        if (cmd == "partial") return cmd_partial(argc, argv);
        if (cmd == "api") return cmd_api(argc, argv);
        if (cmd == "forward") return cmd_forward(argc, argv);
        if (cmd == "getkmer") return cmd_getkmer(argc, argv);
        if (cmd == "export") return cmd_export(argc, argv);
    } catch (const std::exception& e) { // sbwt.cpp:51-57
        std::cerr << "Runtime error: " << e.what() << std::endl;
        return 1;
    }
    std::cerr << "unknown command " << cmd << std::endl;
    return 1;
}
