// cli_search.cpp -- `sbwt search` on the GPU path.
//
// Mirrors search_main / run_queries / run_file / run_queries_streaming /
// run_queries_not_streaming / print_vector of the reference (src/CLI/sbwt_search.cpp:21-260) and
// the dispatcher's error handling (src/CLI/sbwt.cpp:42-57): same options, same list-file mode,
// same text output (one "<value> " per k-mer, '\n' per read, -1 for a miss), same exit codes,
// and the same two timing log lines. Only the plain-matrix variant is served; the others are
// refused (they are out of scope).
//
// Differences: reads are answered in batches and the output text is formatted on the device
// (sbwt_gpu_query_host_text through sbwt::SBWT<>), so "us/query (excluding I/O etc)" is the time spent inside
// the batch calls, which here includes handing the text to the output file; extra options --device,
// --batch-bases and --threads (gzip blocks are deflated in parallel).
//
//   sbwt_search [search] -o <out | list.txt> -i <index.sbwt> -q <reads.(fa|fq)[.gz] | list.txt> [-z]
#include <algorithm>
#include <atomic>
#include <mutex>
#include <deque>
#include <condition_variable>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <exception>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <zlib.h>

#include "SBWT.hh"

// The readers write a batch's bases straight into page-locked memory (the device then takes them by DMA, without a
// staging copy): the allocator fastx.hpp gives its batch vector. g_pinned_device is the device whose context allocates.
static int g_pinned_device = 0;
template <typename T>
struct PinnedAllocator {
    typedef T value_type;
    PinnedAllocator() = default;
    template <typename U>
    PinnedAllocator(const PinnedAllocator<U>&) {}
    T* allocate(size_t n) {
        void* p = nullptr;
        if (sbwt_gpu_host_alloc_on(g_pinned_device, n * sizeof(T), &p)) throw std::bad_alloc();
        return (T*)p;
    }
    void deallocate(T* p, size_t) { sbwt_gpu_host_free(p); }
    template <typename U>
    bool operator==(const PinnedAllocator<U>&) const { return true; }
    template <typename U>
    bool operator!=(const PinnedAllocator<U>&) const { return false; }
};
#define SBWT_B200_ASCII_ALLOCATOR PinnedAllocator
#include "fastx.hpp"

using std::string;
using std::vector;

static const auto program_start = std::chrono::steady_clock::now();

static long long cur_time_micros() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - program_start).count();
}

// write_log, src/globals.cpp:93-105 (timestamp in the reference's "milliseconds since start" quirk)
static void write_log(const string& message) {
    std::time_t t = std::time(nullptr);
    string ts = std::asctime(std::localtime(&t));
    ts.pop_back();
    std::cerr << std::setprecision(4) << std::fixed << cur_time_micros() / 1000.0 << " " << ts << " " << message << std::endl;
}

static vector<string> readlines(const string& filename) { // src/globals.cpp:27-34
    vector<string> lines;
    std::ifstream in(filename);
    if (!in.good()) throw std::runtime_error("Error opening file: " + filename);
    string line;
    while (std::getline(in, line)) lines.push_back(line);
    return lines;
}

static void check_readable(const string& filename) { // src/globals.cpp:39-41
    std::ifstream f(filename);
    if (!f.good()) throw std::runtime_error("Error opening file: " + filename);
}

static void check_writable(const string& filename) { // src/globals.cpp:44-46 (opens with app: creates the file)
    std::ofstream f(filename, std::ofstream::out | std::ofstream::app);
    if (!f.good()) throw std::runtime_error("Error opening file: " + filename);
}

// Output file. The text itself (print_vector, sbwt_search.cpp:21-43) is produced on the device
// (sbwt_gpu_query_host_text) and arrives here in pieces. With -z every piece is cut into blocks that are
// deflated in parallel, each as its own gzip member; the concatenation is a valid gzip file that
// decompresses to the same bytes as the reference's zstr output (tests/test_CLI.hh:12-18 compares
// decompressed content).
struct Writer {
    FILE* fp = nullptr;
    bool gzip = false;
    int threads = 8;
    static constexpr size_t kGzBlock = (size_t)1 << 20;
    Writer(const string& name, bool gzip, int threads) : gzip(gzip), threads(std::max(1, threads)) {
        fp = fopen(name.c_str(), "wb");
        if (!fp) throw std::runtime_error("Error opening file: " + name);
    }
    ~Writer() {
        if (fp) fclose(fp); // (error paths; finish() closes and checks on the normal one)
    }
    void put(const char* p, size_t n) {
        if (n && fwrite(p, 1, n, fp) != n) throw std::runtime_error("Error writing output");
    }
    static void deflate_member(const char* src, size_t n, string& out) {
        z_stream z;
        memset(&z, 0, sizeof z);
        if (deflateInit2(&z, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
            throw std::runtime_error("Error writing gzip output");
        out.resize(deflateBound(&z, (uLong)n) + 32);
        z.next_in = (Bytef*)src;
        z.avail_in = (uInt)n;
        z.next_out = (Bytef*)&out[0];
        z.avail_out = (uInt)out.size();
        const int rc = deflate(&z, Z_FINISH);
        const size_t produced = out.size() - z.avail_out;
        deflateEnd(&z);
        if (rc != Z_STREAM_END) throw std::runtime_error("Error writing gzip output");
        out.resize(produced);
    }
    void write(const char* text, size_t n) {
        if (n == 0) return;
        if (!gzip) { put(text, n); return; }
        const size_t nb = (n + kGzBlock - 1) / kGzBlock;
        vector<string> parts(nb);
        std::atomic<size_t> next{0};
        std::atomic<bool> failed{false};
        auto work = [&]() {
            for (size_t b; (b = next.fetch_add(1)) < nb;) {
                try { deflate_member(text + b * kGzBlock, std::min(kGzBlock, n - b * kGzBlock), parts[b]); }
                catch (...) { failed = true; }
            }
        };
        vector<std::thread> th;
        for (int t = 1; t < (int)std::min<size_t>((size_t)threads, nb); t++) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
        if (failed) throw std::runtime_error("Error writing gzip output");
        for (const string& part : parts) put(part.data(), part.size());
    }
    void finish() { // an empty output is still a valid (empty) gzip stream
        if (gzip && ftell(fp) == 0) {
            string part;
            deflate_member("", 0, part);
            put(part.data(), part.size());
        }
        // a full disk shows up here at the latest: do not report success over a truncated file
        const bool bad = fflush(fp) != 0 || ferror(fp);
        const int rc = fclose(fp);
        fp = nullptr;
        if (bad || rc != 0) throw std::runtime_error("Error writing output");
    }
};

struct SinkState {
    Writer* writer;
    string error;
};

static int text_sink(void* user, const char* text, int64_t n_bytes) {
    SinkState* st = (SinkState*)user;
    try {
        st->writer->write(text, (size_t)n_bytes);
        return 0;
    } catch (const std::exception& e) {
        st->error = e.what();
        return 1;
    }
}

struct Options {
    int64_t batch_bases = (int64_t)64 << 20;
    int64_t batch_reads = (int64_t)1 << 20;
    int threads = 8;
};

// run_file (sbwt_search.cpp:93-105): streaming search when the index supports it, else search() per k-mer.
// One batch over several devices (--devices): the reads are cut into contiguous ranges of (almost) equal total bases,
// one host thread per replica produces the text of its range, and the ranges are written in order. Nothing is exchanged
// between the devices. The text is STREAMED: every replica hands its pieces to a bounded
// queue, the calling thread writes the queue of range 0 while it fills, then range 1, ...; a replica whose queue is full
// waits (its pipeline stalls), so the memory held is bounded whatever the batch size.
struct PieceQueue {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<string> pieces;
    size_t bytes = 0;
    bool done = false, abandoned = false;
    static constexpr size_t kCap = (size_t)256 << 20;
};

static int queue_sink(void* user, const char* text, int64_t n_bytes) {
    PieceQueue* q = static_cast<PieceQueue*>(user);
    std::unique_lock<std::mutex> g(q->mu);
    q->cv.wait(g, [&] { return q->bytes < PieceQueue::kCap || q->abandoned; });
    if (q->abandoned) return 1;
    q->pieces.emplace_back(text, (size_t)n_bytes);
    q->bytes += (size_t)n_bytes;
    q->cv.notify_all();
    return 0;
}

static int64_t query_batch_multi(const vector<const sbwt::plain_matrix_sbwt_t*>& replicas, const char* ascii,
                                 const int64_t* offsets, int64_t n, int mode, Writer& writer) {
    const size_t D = replicas.size();
    vector<int64_t> cuts(D + 1, n);
    cuts[0] = 0;
    const int64_t total = offsets[(size_t)n] - offsets[0];
    for (size_t d = 1; d < D; d++) {
        const int64_t target = offsets[0] + total * (int64_t)d / (int64_t)D;
        const int64_t c = std::lower_bound(offsets, offsets + n + 1, target) - offsets;
        cuts[d] = std::min<int64_t>(n, std::max<int64_t>(cuts[d - 1], c));
    }
    vector<PieceQueue> queues(D);
    vector<int64_t> lookups(D, 0);
    vector<std::exception_ptr> errors(D);
    auto work = [&](size_t d) {
        try {
            const int64_t r0 = cuts[d], nr = cuts[d + 1] - r0;
            if (nr > 0)
                lookups[d] = replicas[d]->query_batch_text(ascii, offsets + r0, nr, mode, SBWT_GPU_CASE_UPPER, queue_sink, &queues[d]);
        } catch (...) { errors[d] = std::current_exception(); }
        std::lock_guard<std::mutex> g(queues[d].mu);
        queues[d].done = true;
        queues[d].cv.notify_all();
    };
    vector<std::thread> th;
    for (size_t d = 0; d < D; d++) th.emplace_back(work, d);
    std::exception_ptr write_error;
    for (size_t d = 0; d < D; d++) { // in order: everything of range d before anything of range d + 1
        PieceQueue& q = queues[d];
        for (;;) {
            string piece;
            {
                std::unique_lock<std::mutex> g(q.mu);
                q.cv.wait(g, [&] { return !q.pieces.empty() || q.done; });
                if (q.pieces.empty()) break;
                piece.swap(q.pieces.front());
                q.pieces.pop_front();
                q.bytes -= piece.size();
                q.cv.notify_all();
            }
            if (!write_error && !errors[d]) {
                try { writer.write(piece.data(), piece.size()); }
                catch (...) { write_error = std::current_exception(); }
            }
            if (write_error)
                for (PieceQueue& x : queues) { std::lock_guard<std::mutex> g(x.mu); x.abandoned = true; x.cv.notify_all(); }
        }
        if (errors[d] && !write_error) { // a failed range: nothing after it is written (what was before it already is)
            write_error = errors[d];
            for (PieceQueue& x : queues) { std::lock_guard<std::mutex> g(x.mu); x.abandoned = true; x.cv.notify_all(); }
        }
    }
    for (auto& t : th) t.join();
    if (write_error) std::rethrow_exception(write_error);
    int64_t n_lookups = 0;
    for (size_t d = 0; d < D; d++) n_lookups += lookups[d];
    return n_lookups;
}

static int64_t run_file(const string& infile, const string& outfile, const vector<const sbwt::plain_matrix_sbwt_t*>& replicas, bool gzip_output,
                        const Options& opt, long long& query_micros) {
    const sbwt::plain_matrix_sbwt_t& index = *replicas[0];
    const long long file_t0 = cur_time_micros();
    sbwt_b200::ParallelFastxReader reader(infile, opt.threads); // same batches and errors as the serial FastxReader
    Writer writer(outfile, gzip_output, opt.threads);
    SinkState sink{&writer, ""};
    const bool streaming = index.has_streaming_query_support();
    write_log(string(streaming ? "Running streaming queries from input file " : "Running non-streaming queries from input file ") + infile +
              " to output file " + outfile);
    // two batch buffers: batch i + 1 is parsed by a helper thread while the device answers batch i. A parse error
    // surfaces when its batch is due, after the output of everything before it has been written (as in the reference,
    // which parses and queries read by read)
    // (the bases land in page-locked memory as they are parsed -- PinnedAllocator above --, the offsets are copied there)
    struct Batch {
        sbwt_b200::AsciiVec ascii;
        vector<int64_t> offsets;
        int64_t n = 0;
        std::exception_ptr error;
        int64_t* p_off = nullptr;
        size_t cap_off = 0;
        ~Batch() { sbwt_gpu_host_free(p_off); }
    } batches[2];
    auto stage = [&](Batch& b) {
        if (b.n <= 0) return;
        const size_t no = b.offsets.size();
        if (no > b.cap_off) {
            sbwt_gpu_host_free(b.p_off); b.p_off = nullptr; b.cap_off = 0;
            void* p = nullptr;
            if (sbwt_gpu_host_alloc_on(g_pinned_device, (no + no / 8 + 512) * sizeof(int64_t), &p)) throw std::runtime_error(sbwt_gpu_last_error());
            b.p_off = (int64_t*)p; b.cap_off = no + no / 8 + 512;
        }
        memcpy(b.p_off, b.offsets.data(), no * sizeof(int64_t));
    };
    auto parse = [&](Batch& b) {
        try {
            if (b.ascii.capacity() == 0) b.ascii.reserve((size_t)opt.batch_bases + ((size_t)1 << 20)); // (one page-locked allocation instead of a doubling sequence)
            b.n = reader.next_batch(opt.batch_bases, opt.batch_reads, b.ascii, b.offsets);
            stage(b);
        }
        catch (...) { b.error = std::current_exception(); b.n = 0; }
    };
    int64_t n_queries = 0;
    long long first_parse_micros = cur_time_micros(), join_micros = 0;
    parse(batches[0]);
    first_parse_micros = cur_time_micros() - first_parse_micros;
    for (int turn = 0;; turn ^= 1) {
        Batch& b = batches[turn];
        if (b.error) std::rethrow_exception(b.error);
        if (b.n == 0) break;
        std::thread ahead(parse, std::ref(batches[turn ^ 1]));
        const long long t0 = cur_time_micros();
        try {
            const int mode = streaming ? SBWT_GPU_MODE_STREAMING : SBWT_GPU_MODE_SEARCH;
            if (replicas.size() > 1) n_queries += query_batch_multi(replicas, b.ascii.data(), b.p_off, b.n, mode, writer);
            else n_queries += index.query_batch_text(b.ascii.data(), b.p_off, b.n, mode, SBWT_GPU_CASE_UPPER, text_sink, &sink);
        } catch (const std::runtime_error&) {
            ahead.join();
            if (!sink.error.empty()) throw std::runtime_error(sink.error);
            throw;
        } catch (...) {
            ahead.join();
            throw;
        }
        query_micros += cur_time_micros() - t0;
        const long long tj = cur_time_micros();
        ahead.join();
        join_micros += cur_time_micros() - tj;
    }
    const long long tf = cur_time_micros();
    writer.finish();
    if (getenv("SBWT_B200_CLI_TIMING"))
        write_log("timing: first batch parsed in " + std::to_string(first_parse_micros / 1e6) + " s, device calls " + std::to_string(query_micros / 1e6) +
                  " s, waiting for the parser " + std::to_string(join_micros / 1e6) + " s, closing the output " + std::to_string((cur_time_micros() - tf) / 1e6) + " s");
    {
        const double file_s = (double)(cur_time_micros() - file_t0) / 1e6;
        write_log("queries: " + std::to_string(n_queries) + " in " + std::to_string(file_s) + " s of parsing + querying + writing (index load excluded): " +
                  std::to_string((double)n_queries / std::max(file_s, 1e-9)) + " lookups/s");
    }
    write_log("us/query: " + std::to_string((double)query_micros / std::max<int64_t>(n_queries, 1)) + " (excluding I/O etc)");
    return n_queries;
}

static void print_help(const char* prog) {
    std::cerr << "Query all k-mers of all input reads.\nUsage:\n  " << prog << " [OPTION...]\n\n"
              << "  -o, --out-file arg     Output filename.\n"
              << "  -i, --index-file arg   Index input file.\n"
              << "  -q, --query-file arg   The query in FASTA or FASTQ format, possibly gzipped. Multi-line FASTQ is not\n"
              << "                         supported. If the file extension is .txt, this is interpreted as a list of query\n"
              << "                         files, one per line. In this case, --out-file is also interpreted as a list of\n"
              << "                         output files in the same manner, one line for each input file.\n"
              << "  -z, --gzip-output      Writes output in gzipped form. This can shrink the output files by an order of\n"
              << "                         magnitude.\n"
              << "      --device arg       CUDA device (default 0)\n"
              << "      --devices arg      Comma-separated CUDA devices: the index is replicated on each of them and every\n"
              << "                         batch of reads is split over them (no data is exchanged between devices)\n"
              << "      --batch-bases arg  Read bases per GPU batch (default 67108864)\n"
              << "      --threads arg      Host threads for parsing the query file and for gzip output (default 8)\n"
              << "  -h, --help             Print usage\n" << std::endl;
}

static int search_main(int argc, char** argv) {
    const long long micros_start = cur_time_micros();
    string out_file, index_file, query_file;
    bool gzip_output = false, have_o = false, have_i = false, have_q = false;
    int device = 0;
    vector<int> devices; // --devices a,b,...: one replica of the index per listed device, every batch split over them
    Options opt;
    if (argc == 1) { print_help(argv[0]); return 1; }
    // cxxopts' forms (sbwt_search.cpp:149-165): --name value, --name=value, -n value, -nvalue, grouped short flags (-zo out)
    struct Opt { char short_name; const char* long_name; bool takes_value; };
    static const Opt table[] = {{'o', "out-file", true}, {'i', "index-file", true}, {'q', "query-file", true}, {'z', "gzip-output", false},
                                {'h', "help", false}, {0, "device", true}, {0, "devices", true}, {0, "batch-bases", true}, {0, "threads", true}};
    bool want_help = false;
    auto apply = [&](const Opt& o, const string& v) {
        const string name = o.long_name;
        if (name == "out-file") { out_file = v; have_o = true; }
        else if (name == "index-file") { index_file = v; have_i = true; }
        else if (name == "query-file") { query_file = v; have_q = true; }
        else if (name == "gzip-output") gzip_output = true;
        else if (name == "help") want_help = true;
        else if (name == "device") device = std::stoi(v);
        else if (name == "devices") {
            std::stringstream ss(v);
            for (string tok; std::getline(ss, tok, ',');) devices.push_back(std::stoi(tok));
        }
        else if (name == "batch-bases") opt.batch_bases = std::stoll(v);
        else if (name == "threads") opt.threads = std::stoi(v);
    };
    for (int i = 1; i < argc; i++) {
        const string a = argv[i];
        auto next_value = [&](const string& shown) -> string {
            if (i + 1 >= argc) throw std::runtime_error("Option '" + shown + "' is missing an argument");
            return argv[++i];
        };
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            const string name = a.substr(2, eq == string::npos ? string::npos : eq - 2);
            const Opt* o = nullptr;
            for (const Opt& t : table) if (name == t.long_name) o = &t;
            if (!o) throw std::runtime_error("Option '" + name + "' does not exist");
            if (!o->takes_value) apply(*o, "");
            else apply(*o, eq != string::npos ? a.substr(eq + 1) : next_value(name));
        } else if (a.size() > 1 && a[0] == '-' && a != "--") {
            for (size_t j = 1; j < a.size(); j++) {
                const Opt* o = nullptr;
                for (const Opt& t : table) if (t.short_name && a[j] == t.short_name) o = &t;
                if (!o) throw std::runtime_error(string("Option '") + a[j] + "' does not exist");
                if (!o->takes_value) { apply(*o, ""); continue; }
                apply(*o, j + 1 < a.size() ? a.substr(j + 1) : next_value(string(1, a[j]))); // the rest of the token, or the next one
                break;
            }
        } // (anything else is a positional argument: cxxopts keeps those aside without complaint, and so does this)
    }
    if (want_help) { print_help(argv[0]); return 1; }
    if (opt.batch_bases <= 0) throw std::runtime_error("Option 'batch-bases' must be positive");
    if (opt.threads <= 0) throw std::runtime_error("Option 'threads' must be positive");
    if (!have_i) throw std::runtime_error("Option 'index-file' has no value");
    check_readable(index_file);
    if (!have_q) throw std::runtime_error("Option 'query-file' has no value");
    vector<string> input_files, output_files;
    const bool multi_file = query_file.size() >= 4 && query_file.substr(query_file.size() - 4) == ".txt";
    if (multi_file) input_files = readlines(query_file);
    else input_files = {query_file};
    for (const string& f : input_files) check_readable(f);
    if (!have_o) throw std::runtime_error("Option 'out-file' has no value");
    if (multi_file) output_files = readlines(out_file);
    else output_files = {out_file};
    for (const string& f : output_files) check_writable(f);

    std::ifstream in(index_file, std::ios::binary);
    if (!in.good()) throw std::runtime_error("Error opening file: " + index_file);
    const string variant = sbwt_b200::load_variant_string(in); // sbwt_search.cpp:194
    static const char* known[] = {"plain-matrix", "rrr-matrix", "mef-matrix", "plain-split", "rrr-split", "mef-split",
                                  "plain-concat", "mef-concat", "plain-subsetwt", "rrr-subsetwt"};
    if (std::find(std::begin(known), std::end(known), variant) == std::end(known)) {
        std::cerr << "Error loading index from file: unrecognized variant specified in the file" << std::endl;
        return 1;
    }
    write_log("Loading the index variant " + variant);
    if (variant != "plain-matrix") {
        std::cerr << "Error: the GPU query path serves the plain-matrix variant only (index is " << variant << ")" << std::endl;
        return 1;
    }
    if (devices.empty()) devices.push_back(device);
    g_pinned_device = devices[0];
    vector<std::unique_ptr<sbwt::plain_matrix_sbwt_t>> owned;
    vector<const sbwt::plain_matrix_sbwt_t*> replicas;
    for (size_t d = 0; d < devices.size(); d++) {
        owned.emplace_back(new sbwt::plain_matrix_sbwt_t(devices[d]));
        if (d == 0) owned.back()->load(in);
        else { // every replica reads the file itself (the variant string first, sbwt_search.cpp:194)
            std::ifstream again(index_file, std::ios::binary);
            if (!again.good()) throw std::runtime_error("Error opening file: " + index_file);
            sbwt_b200::load_variant_string(again);
            owned.back()->load(again);
        }
        replicas.push_back(owned.back().get());
    }

    if (input_files.size() != output_files.size())
        throw std::runtime_error("Number of input and output files does not match (" + std::to_string(input_files.size()) + " vs " +
                                 std::to_string(output_files.size()) + ")");
    int64_t number_of_queries = 0;
    for (size_t i = 0; i < input_files.size(); i++) {
        long long micros = 0;
        number_of_queries += run_file(input_files[i], output_files[i], replicas, gzip_output, opt, micros);
    }
    const long long total_micros = cur_time_micros() - micros_start;
    write_log("us/query end-to-end: " + std::to_string((double)total_micros / std::max<int64_t>(number_of_queries, 1)));
    return 0;
}

int main(int argc, char** argv) { // sbwt.cpp:19-59: strips the sub-command, reports exceptions, returns 1
    try {
        if (argc >= 2 && string(argv[1]) == "search") return search_main(argc - 1, argv + 1);
        if (argc >= 2 && (string(argv[1]) == "build" || string(argv[1]) == "build-variant" || string(argv[1]) == "ascii-export")) {
            std::cerr << "Error: only `search` is available in the GPU query build" << std::endl;
            return 1;
        }
        return search_main(argc, argv);
    } catch (const std::runtime_error& e) {
        std::cerr << "Runtime error: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
