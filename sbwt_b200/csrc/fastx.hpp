// fastx.hpp -- FASTA / FASTQ (.gz) batch reader with the observable behaviour of the reference's
// seq_io::Reader::get_next_read_to_buffer (SeqIO/include/SeqIO/SeqIO.hh:255-360) over
// Buffered_ifstream (SeqIO/include/SeqIO/buffered_streams.hh:67-111):
//   * file type from the extension only (SeqIO.hh:19-20, 62-93), ".gz" through zlib;
//   * FASTA sequences may span lines and are concatenated; FASTQ is strictly 4 lines per record;
//   * the same failures: file not starting with '>' / '@', empty line, empty sequence, a last
//     line without '\n' ("ended unexpectedly") -- thrown as std::runtime_error;
//   * bytes are handed on untouched ('\r' stays a base); upper-casing is NOT done here -- the
//     device packer folds case (SBWT_GPU_CASE_UPPER), which is what SeqIO.hh:294-297 does per byte.
// Instead of one read per call it fills a batch: concatenated bases + offsets, the layout the
// C ABI takes.
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <zlib.h>

// The bases of a batch are written into a std::vector<char, SBWT_B200_ASCII_ALLOCATOR<char>>: std::allocator unless the
// including file defines another one first (the command line uses page-locked memory, which the device reads by DMA).
#ifndef SBWT_B200_ASCII_ALLOCATOR
#define SBWT_B200_ASCII_ALLOCATOR std::allocator
#endif

namespace sbwt_b200 {

typedef std::vector<char, SBWT_B200_ASCII_ALLOCATOR<char>> AsciiVec;

enum class SeqFormat { FASTA, FASTQ };

struct FileFormat {
    SeqFormat format;
    bool gzipped;
};

// figure_out_file_format, SeqIO.hh:62-93
inline FileFormat figure_out_file_format(std::string filename) {
    const std::string original = filename;
    bool gz = false;
    if (filename.size() >= 3 && filename.compare(filename.size() - 3, 3, ".gz") == 0) {
        filename.resize(filename.size() - 3);
        gz = true;
    }
    const size_t dot = filename.rfind('.');
    if (dot != std::string::npos) {
        const std::string ext = filename.substr(dot);
        for (const char* s : {".fasta", ".fna", ".ffn", ".faa", ".frn", ".fa"})
            if (ext == s) return {SeqFormat::FASTA, gz};
        for (const char* s : {".fastq", ".fq"})
            if (ext == s) return {SeqFormat::FASTQ, gz};
    }
    throw std::runtime_error("Unknown file format: " + original);
}

class FastxReader {
    std::string filename;
    SeqFormat format;
    bool gz;
    FILE* fp = nullptr;
    gzFile gzf = nullptr;
    std::vector<char> buf;
    const char* mem = nullptr; // memory source (ParallelFastxReader's fallback): [mem, mem + mem_len) ...
    size_t mem_len = 0, mem_pos = 0;
    gzFile cont = nullptr;     // ... continued, when it is used up, by the rest of this gzip stream (not owned)
    size_t pos = 0, size = 0;
    size_t taken = 0;    // bytes handed out by get() before the current buffer
    bool is_eof = false; // true once a get() ran past the end (Buffered_ifstream::eof)

    bool refill() {
        taken += size;
        if (mem && mem_pos < mem_len) {
            size = std::min(buf.size(), mem_len - mem_pos);
            memcpy(buf.data(), mem + mem_pos, size);
            mem_pos += size;
        } else if (mem) {
            size = 0;
            if (cont) {
                int n = gzread(cont, buf.data(), (unsigned)buf.size());
                if (n < 0) throw std::runtime_error("Error reading gzip file " + filename);
                size = (size_t)n;
            }
        } else if (gz) {
            int n = gzread(gzf, buf.data(), (unsigned)buf.size());
            if (n < 0) throw std::runtime_error("Error reading gzip file " + filename);
            size = (size_t)n;
        } else {
            size = fread(buf.data(), 1, buf.size(), fp);
        }
        pos = 0;
        return size > 0;
    }
    inline bool get(char& c) {
        if (is_eof) return false;
        if (pos == size && !refill()) { is_eof = true; return false; }
        c = buf[pos++];
        return true;
    }
    // getline into out (appending); returns nothing -- callers test is_eof like the reference does
    inline void getline_append(AsciiVec* out) {
        for (;;) {
            char c;
            if (!get(c)) return;
            if (c == '\n') return;
            if (out) out->push_back(c);
        }
    }

public:
    explicit FastxReader(const std::string& filename) : filename(filename), buf(4 << 20) {
        FileFormat ff = figure_out_file_format(filename);
        format = ff.format;
        gz = ff.gzipped;
        if (gz) {
            gzf = gzopen(filename.c_str(), "rb");
            if (!gzf) throw std::runtime_error("Error opening file " + filename);
            gzbuffer(gzf, 1 << 20);
        } else {
            fp = fopen(filename.c_str(), "rb");
            if (!fp) throw std::runtime_error("Error opening file " + filename);
        }
        char c = 0;
        get(c); // read_first_char_and_sanity_check, SeqIO.hh:178-189
        if (format == SeqFormat::FASTA && c != '>') throw std::runtime_error("ERROR: FASTA file " + filename + " does not start with '>'");
        if (format == SeqFormat::FASTQ && c != '@') throw std::runtime_error("ERROR: FASTQ file " + filename + " does not start with '@'");
    }
    // Serial parser over a memory range that starts at a record (whose first byte is consumed here, like the
    // look-ahead get() at the end of next_read does); `filename` only feeds the error messages.
    FastxReader(const std::string& filename, SeqFormat format, const char* mem, size_t len, gzFile cont = nullptr)
        : filename(filename), format(format), gz(false), buf(1 << 20), mem(mem ? mem : ""), mem_len(len), cont(cont) {
        char c = 0;
        get(c);
    }
    // bytes consumed so far, not counting the look-ahead byte of the next record
    size_t consumed() const { return is_eof ? taken + pos : taken + pos - 1; }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;
    ~FastxReader() {
        if (fp) fclose(fp);
        if (gzf) gzclose(gzf);
    }

    // Appends the next read's bases to `ascii`; returns its length, 0 at end of file.
    int64_t next_read(AsciiVec& ascii) {
        if (is_eof) return 0;
        const size_t start = ascii.size();
        if (format == SeqFormat::FASTA) {
            char c = 0;
            getline_append(nullptr); // header
            if (is_eof) throw std::runtime_error("FASTA file " + filename + " ended unexpectedly.");
            get(c);
            if (c == '\n') throw std::runtime_error("Empty line in FASTA file " + filename + ".");
            if (c == '>') throw std::runtime_error("Empty sequence in FASTA file " + filename + ".");
            while (c != '>') {
                ascii.push_back(c);
                getline_append(&ascii);
                if (is_eof) throw std::runtime_error("FASTA file " + filename + " ended unexpectedly.");
                get(c); // first byte of the next line
                if (c == '\n') throw std::runtime_error("Empty line inside sequence in file " + filename + ".");
                if (is_eof) break;
            }
        } else {
            getline_append(nullptr);
            if (is_eof) throw std::runtime_error("FASTQ file " + filename + " ended unexpectedly.");
            getline_append(&ascii);
            if (is_eof) throw std::runtime_error("FASTQ file " + filename + " ended unexpectedly.");
            getline_append(nullptr);
            if (is_eof) throw std::runtime_error("FASTQ file " + filename + " ended unexpectedly.");
            getline_append(nullptr);
            if (is_eof) throw std::runtime_error("FASTQ file " + filename + " ended unexpectedly.");
            char c;
            get(c); // the '@' of the next record, or end of file
            if (ascii.size() == start) throw std::runtime_error("Error: empty sequence in FASTQ file.");
        }
        return (int64_t)(ascii.size() - start);
    }

    // Fills a batch of whole reads: stops before exceeding max_bases (unless the batch is empty) or
    // max_reads. offsets gets n+1 entries starting at 0. Returns the number of reads (0 = end of file).
    // A malformed record does not swallow the reads in front of it: they are returned as a (short) batch and the error is
    // raised by the next call -- the reference answers and prints read by read, so everything before the bad record has
    // been written when it throws (sbwt_search.cpp:45-65 over SeqIO.hh:255-360).
    int64_t next_batch(int64_t max_bases, int64_t max_reads, AsciiVec& ascii, std::vector<int64_t>& offsets) {
        ascii.clear();
        offsets.clear();
        offsets.push_back(0);
        if (pending_error) {
            std::exception_ptr e = pending_error;
            pending_error = nullptr;
            std::rethrow_exception(e);
        }
        while ((int64_t)offsets.size() - 1 < max_reads && (int64_t)ascii.size() < max_bases) {
            try {
                if (next_read(ascii) == 0) break;
            } catch (...) {
                ascii.resize((size_t)offsets.back()); // (whatever the failed record had appended)
                if (offsets.size() == 1) throw;
                pending_error = std::current_exception();
                break;
            }
            offsets.push_back((int64_t)ascii.size());
        }
        return (int64_t)offsets.size() - 1;
    }

private:
    std::exception_ptr pending_error;
};

// Parallel batch reader for uncompressed FASTA / FASTQ files: the same batches, byte for byte, as
// FastxReader::next_batch (and the same exceptions), produced by `threads` host threads.
//
// The file is mapped; for every batch a window of it is cut into lines by a parallel newline scan.
// Line structure alone fixes the records -- FASTQ is strictly four lines per record (SeqIO.hh:316-343: no
// '@' test after the first record), a FASTA header is any line that starts with '>' after a sequence line
// (SeqIO.hh:262-312) -- so record boundaries need no guessing, and the sequence lines are copied into the
// batch in parallel at prefix-summed offsets. Anything the reference treats specially (empty lines, empty
// sequences, a last line without '\n', a FASTQ record whose header line is empty) is not reasoned about
// here: the first batch that meets such a line is handed, from its first record on, to the serial parser
// above, which raises the reference's error or carries on exactly as the reference would.
class ParallelFastxReader {
    std::string filename;
    SeqFormat format;
    int threads;
    int fd = -1;
    const char* data = nullptr;
    size_t size = 0, cur = 0; // cur: first byte of the next record
    std::unique_ptr<FastxReader> serial; // the fallback after an anomaly (or for inputs that cannot be mapped)
    // gzip input: the stream is inflated (sequentially: a deflate stream cannot be entered in the middle) into a sliding
    // buffer; data / size / cur then refer to that buffer and the parsing is the same parallel code
    gzFile gzf = nullptr;
    std::vector<char> zbuf;
    bool z_eof = false;
    bool all_here = true; // everything up to the end of the input is in [data, data + size)
    // BGZF input (bgzip, htslib: gzip members of at most 64 KiB whose header carries the member's size in a 'BC' extra
    // field): member boundaries are known without inflating, so the members are inflated in parallel. A plain gzip
    // stream has no such index and is inflated by one thread ahead of the parallel parser.
    const unsigned char* zmap = nullptr;
    size_t zmap_size = 0, zmap_pos = 0;
    bool bgzf = false;

    // size of the BGZF member at `at` (0 = not a BGZF member header)
    size_t bgzf_member_size(size_t at) const {
        if (at + 18 > zmap_size) return 0;
        const unsigned char* h = zmap + at;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
        const size_t xlen = h[10] | (h[11] << 8);
        if (at + 12 + xlen > zmap_size) return 0;
        for (size_t x = 12; x + 4 <= 12 + xlen;) {
            const size_t slen = h[x + 2] | (h[x + 3] << 8);
            if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2 && x + 6 <= 12 + xlen) {
                const size_t bsize = (size_t)(h[x + 4] | (h[x + 5] << 8)) + 1;
                return (bsize >= 12 + xlen + 8 && at + bsize <= zmap_size) ? bsize : 0;
            }
            x += 4 + slen;
        }
        return 0;
    }

    void bgzf_fill(size_t want) {
        if (cur > 0 && cur == size) { size = 0; cur = 0; }
        if (cur > ((size_t)64 << 20)) {
            memmove(zbuf.data(), zbuf.data() + cur, size - cur);
            size -= cur;
            cur = 0;
        }
        while (!z_eof && size - cur < want) {
            // the next members, up to ~64 MB of text: headers are hopped over, the sizes come from each member's ISIZE
            struct Member { size_t at, len, out, out_len, hdr; };
            std::vector<Member> ms;
            size_t out_total = 0;
            while (zmap_pos < zmap_size && out_total < ((size_t)64 << 20)) {
                const size_t len = bgzf_member_size(zmap_pos);
                if (!len) throw std::runtime_error("Error reading gzip file " + filename);
                const unsigned char* h = zmap + zmap_pos;
                const size_t isize = (size_t)h[len - 4] | ((size_t)h[len - 3] << 8) | ((size_t)h[len - 2] << 16) | ((size_t)h[len - 1] << 24);
                ms.push_back(Member{zmap_pos, len, out_total, isize, 12 + (size_t)(h[10] | (h[11] << 8))});
                out_total += isize;
                zmap_pos += len;
            }
            if (zmap_pos >= zmap_size) z_eof = true;
            if (zbuf.size() < size + out_total) zbuf.resize(std::max(zbuf.size() * 2, size + out_total));
            char* base = zbuf.data() + size;
            const unsigned char* zm = zmap;
            std::atomic<bool> bad{false};
            const Member* mp = ms.data();
            parallel_for(ms.size(), [=, &bad](size_t a, size_t b, size_t) {
                for (size_t i = a; i < b; i++) {
                    if (mp[i].out_len == 0) continue; // (bgzip's end-of-file marker is an empty member)
                    z_stream z;
                    memset(&z, 0, sizeof z);
                    if (inflateInit2(&z, -15) != Z_OK) { bad = true; return; } // raw deflate: the member's header was parsed above
                    z.next_in = (Bytef*)(zm + mp[i].at + mp[i].hdr);
                    z.avail_in = (uInt)(mp[i].len - mp[i].hdr - 8);
                    z.next_out = (Bytef*)(base + mp[i].out);
                    z.avail_out = (uInt)mp[i].out_len;
                    const int rc = inflate(&z, Z_FINISH);
                    if (rc != Z_STREAM_END || z.avail_out != 0) bad = true;
                    inflateEnd(&z);
                }
            });
            if (bad) throw std::runtime_error("Error reading gzip file " + filename);
            size += out_total;
        }
        data = zbuf.data();
        all_here = z_eof;
    }

    // make at least `want` bytes after `cur` available (or reach the end of the stream)
    void z_fill(size_t want) {
        if (bgzf) { bgzf_fill(want); return; }
        if (!gzf) return;
        if (cur > 0 && cur == size) { size = 0; cur = 0; }
        if (cur > ((size_t)64 << 20)) { // drop what has been consumed
            memmove(zbuf.data(), zbuf.data() + cur, size - cur);
            size -= cur;
            cur = 0;
        }
        while (!z_eof && size - cur < want) {
            const size_t step = (size_t)16 << 20;
            if (zbuf.size() < size + step) zbuf.resize(std::max(zbuf.size() * 2, size + step));
            const int n = gzread(gzf, zbuf.data() + size, (unsigned)step);
            if (n < 0) throw std::runtime_error("Error reading gzip file " + filename);
            if (n == 0) z_eof = true;
            size += (size_t)n;
        }
        data = zbuf.data();
        all_here = z_eof;
    }

    struct Line {
        size_t start, end; // [start, end): without the '\n'
        bool header() const { return (end >> 63) != 0; } // the line begins with '>' (noted by the scan, while the byte is in cache)
        size_t stop() const { return end & ~((size_t)1 << 63); }
    };
    // scratch kept across batches (a 64 MB batch of short reads has millions of lines: allocating, zero-filling and
    // page-faulting these arrays anew for every batch cost a third of the parser's time)
    template <typename T>
    struct RawBuf {
        std::unique_ptr<T[]> p;
        size_t cap = 0, n = 0;
        void resize(size_t m) { // contents are not kept and not initialised
            if (m > cap) { cap = m + m / 4 + 1024; p.reset(new T[cap]); }
            n = m;
        }
        size_t size() const { return n; }
        T* data() { return p.get(); }
        const T* data() const { return p.get(); }
        const T& operator[](size_t i) const { return p[i]; }
    };
    typedef RawBuf<Line> LineBuf;
    struct Rec {
        size_t first_line, n_lines; // sequence lines
        int64_t len;
    };
    mutable LineBuf lines_buf;
    RawBuf<Rec> recs_buf;
    mutable std::vector<std::vector<size_t>> nl_buf;
    double file_bytes_per_base = 1.5; // how much of the file a batch of max_bases covers: learnt from the previous batch

    template <typename F>
    void parallel_for(size_t n, F f) const {
        const size_t T = std::max<size_t>(1, std::min<size_t>((size_t)threads, n));
        if (T == 1) { f(0, n, 0); return; }
        std::vector<std::thread> th;
        for (size_t t = 1; t < T; t++) th.emplace_back([=] { f(n * t / T, n * (t + 1) / T, t); });
        f(0, n / T, 0);
        for (auto& x : th) x.join();
    }

    // complete lines of [from, to): every '\n' found ends one. Two parallel passes: collect the newline positions of every
    // slice, then write the lines of every slice at its prefix-summed place.
    void scan_lines(size_t from, size_t to, LineBuf& lines) const {
        const size_t T = (size_t)std::max(1, threads);
        std::vector<std::vector<size_t>>& nl = nl_buf;
        nl.resize(T);
        for (auto& v : nl) v.clear();
        const char* dend = data + to;
        parallel_for(to - from, [&](size_t a, size_t b, size_t t) {
            const char* p = data + from + a;
            const char* e = data + from + b;
            nl[t].reserve((size_t)(e - p) / 64 + 16);
            while (p < e) {
                const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
                if (!q) break;
                // bit 63: the line after this newline begins with '>' (its first byte is in the cache line just scanned)
                nl[t].push_back((size_t)(q - data) | ((q + 1 < dend && q[1] == '>') ? (size_t)1 << 63 : 0));
                p = q + 1;
            }
        });
        constexpr size_t HB = (size_t)1 << 63;
        std::vector<size_t> first(T + 1, 0), start(T, from); // first line index of slice t; start of its first line (+ its header bit)
        size_t prev_start = from | ((from < to && data[from] == '>') ? HB : 0);
        for (size_t t = 0; t < T; t++) {
            first[t + 1] = first[t] + nl[t].size();
            start[t] = prev_start;
            if (!nl[t].empty()) prev_start = ((nl[t].back() & ~HB) + 1) | (nl[t].back() & HB);
        }
        lines.resize(first[T]);
        Line* out = lines.data();
        const std::vector<size_t>* nlp = nl.data();
        const size_t* firstp = first.data();
        const size_t* startp = start.data();
        parallel_for(T, [=](size_t a, size_t b, size_t) {
            for (size_t t = a; t < b; t++) {
                size_t st = startp[t]; // (bit 63: this line is a header)
                Line* o = out + firstp[t];
                for (size_t x : nlp[t]) {
                    *o++ = Line{st & ~HB, (x & ~HB) | (st & HB)};
                    st = ((x & ~HB) + 1) | (x & HB);
                }
            }
        });
    }

    // Every record the window's lines hold, in parallel over slices of the lines. A FASTA record is a header line and
    // the lines up to the next header; the last record of the window runs to its end (whether it is whole is the
    // caller's business). Returns false when the window holds a line the serial parser has to judge -- an empty line,
    // a header directly behind a header, an empty FASTQ header or sequence line -- wherever it is: the caller then
    // hands everything from `cur` on to the serial parser, which is always right and only slower.
    bool form_records(const LineBuf& lines, RawBuf<Rec>& recs) const {
        const size_t nl = lines.size();
        const Line* L = lines.data();
        std::atomic<bool> bad{false};
        if (format == SeqFormat::FASTQ) {
            const size_t n = nl / 4;
            recs.resize(n);
            Rec* R = recs.data();
            parallel_for(n, [=, &bad](size_t a, size_t b, size_t) {
                for (size_t r = a; r < b; r++) {
                    const Line& h = L[4 * r];
                    const Line& q = L[4 * r + 1];
                    if (h.stop() == h.start || q.stop() == q.start) bad = true;
                    R[r] = Rec{4 * r + 1, 1, (int64_t)(q.stop() - q.start)};
                }
            });
            return !bad;
        }
        const size_t T = (size_t)std::max(1, threads);
        std::vector<size_t> cnt(T + 1, 0);
        size_t* cntp = cnt.data();
        if (nl == 0 || !L[0].header()) { recs.resize(0); return nl == 0; }
        parallel_for(nl, [=, &bad](size_t a, size_t b, size_t t) {
            size_t c = 0;
            for (size_t i = a; i < b; i++) {
                if (L[i].stop() == L[i].start) bad = true; // an empty line
                if (L[i].header()) {
                    c++;
                    if (i + 1 < nl && L[i + 1].header()) bad = true; // an empty sequence
                }
            }
            cntp[t + 1] = c;
        });
        if (bad) return false;
        for (size_t t = 0; t < T; t++) cnt[t + 1] += cnt[t];
        recs.resize(cnt[T]);
        Rec* R = recs.data();
        parallel_for(nl, [=](size_t a, size_t b, size_t t) {
            size_t r = cntp[t], i = a;
            while (i < b && !L[i].header()) i++; // (the lines before belong to a record of the slice in front)
            while (i < b) {
                size_t j = i + 1;
                int64_t len = 0;
                while (j < nl && !L[j].header()) { len += (int64_t)(L[j].stop() - L[j].start); j++; }
                R[r++] = Rec{i + 1, j - (i + 1), len};
                i = j;
            }
        });
        return true;
    }

    int64_t from_serial(int64_t max_bases, int64_t max_reads, AsciiVec& ascii, std::vector<int64_t>& offsets) {
        return serial->next_batch(max_bases, max_reads, ascii, offsets);
    }

public:
    ParallelFastxReader(const std::string& filename, int threads) : filename(filename), threads(std::max(1, threads)) {
        const FileFormat ff = figure_out_file_format(filename);
        format = ff.format;
        if (ff.gzipped) {
            fd = open(filename.c_str(), O_RDONLY);
            struct stat zst;
            if (fd >= 0 && fstat(fd, &zst) == 0 && zst.st_size >= 28) {
                void* p = mmap(nullptr, (size_t)zst.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (p != MAP_FAILED) {
                    zmap = (const unsigned char*)p;
                    zmap_size = (size_t)zst.st_size;
                    bgzf = bgzf_member_size(0) != 0;
                    if (!bgzf) { munmap(p, zmap_size); zmap = nullptr; }
                }
            }
            if (!bgzf) {
                if (fd >= 0) { close(fd); fd = -1; }
                gzf = gzopen(filename.c_str(), "rb");
                if (!gzf) throw std::runtime_error("Error opening file " + filename);
                gzbuffer(gzf, 1 << 20);
            }
            z_fill(1);
            const char c = size ? data[0] : 0; // read_first_char_and_sanity_check, SeqIO.hh:178-189
            if (format == SeqFormat::FASTA && c != '>') throw std::runtime_error("ERROR: FASTA file " + filename + " does not start with '>'");
            if (format == SeqFormat::FASTQ && c != '@') throw std::runtime_error("ERROR: FASTQ file " + filename + " does not start with '@'");
            return;
        }
        fd = open(filename.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("Error opening file " + filename);
        struct stat st;
        if (fstat(fd, &st) != 0) throw std::runtime_error("Error opening file " + filename);
        size = (size_t)st.st_size;
        if (size > 0) {
            void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { // e.g. a pipe: read it serially
                close(fd);
                fd = -1;
                serial.reset(new FastxReader(filename));
                return;
            }
            data = (const char*)p;
            madvise((void*)data, size, MADV_SEQUENTIAL);
        }
        // read_first_char_and_sanity_check, SeqIO.hh:178-189 (an empty file reads a 0 byte there)
        const char c = size ? data[0] : 0;
        if (format == SeqFormat::FASTA && c != '>') throw std::runtime_error("ERROR: FASTA file " + filename + " does not start with '>'");
        if (format == SeqFormat::FASTQ && c != '@') throw std::runtime_error("ERROR: FASTQ file " + filename + " does not start with '@'");
    }
    ParallelFastxReader(const ParallelFastxReader&) = delete;
    ParallelFastxReader& operator=(const ParallelFastxReader&) = delete;
    ~ParallelFastxReader() {
        serial.reset();
        if (gzf) gzclose(gzf);
        else if (zmap) munmap((void*)zmap, zmap_size);
        else if (data) munmap((void*)data, size);
        if (fd >= 0) close(fd);
    }

    // Same contract as FastxReader::next_batch.
    int64_t next_batch(int64_t max_bases, int64_t max_reads, AsciiVec& ascii, std::vector<int64_t>& offsets) {
        if (serial) return from_serial(max_bases, max_reads, ascii, offsets);
        offsets.clear();
        offsets.push_back(0);
        if (max_reads <= 0 || max_bases <= 0) { ascii.clear(); return 0; }
        z_fill(1);
        if (cur >= size) { ascii.clear(); return 0; }
        LineBuf& lines = lines_buf;
        RawBuf<Rec>& recs = recs_buf;
        size_t n_take = 0; // records of the window that make the batch
        size_t window = (size_t)((double)max_bases * file_bytes_per_base) + ((size_t)1 << 20);
        bool anomaly = false;
        size_t next_cur = cur;
        for (;;) {
            z_fill(window);
            const size_t wend = std::min(size, cur + window);
            const bool at_eof = wend == size && all_here;
            scan_lines(cur, wend, lines);
            anomaly = !form_records(lines, recs);
            bool full = false; // the batch reached max_reads / max_bases
            next_cur = cur;
            n_take = 0;
            if (!anomaly) {
                const size_t nl = lines.size(), n_all = recs.size();
                const Line* L = lines.data();
                const Rec* R = recs.data();
                const size_t after_all = nl ? L[nl - 1].stop() + 1 : cur; // first byte behind the window's last complete line
                // FASTA: the window's last record ends where the lines end -- whole only at the end of the file, when its
                // last line is terminated and it has a sequence; anything else at the end of the file is the serial parser's
                size_t n_whole = n_all;
                bool tail_anomaly = false;
                if (format == SeqFormat::FASTA && n_all > 0) {
                    const bool last_whole = at_eof && after_all == size && R[n_all - 1].n_lines > 0;
                    if (!last_whole) { n_whole = n_all - 1; tail_anomaly = at_eof; }
                }
                int64_t bases = 0;
                for (size_t r = 0; r < n_whole; r++) {
                    bases += R[r].len;
                    if ((int64_t)(r + 1) >= max_reads || bases >= max_bases) { full = true; n_take = r + 1; break; }
                }
                if (!full) { n_take = n_whole; anomaly = tail_anomaly; }
                if (format == SeqFormat::FASTQ) next_cur = n_take == 0 ? cur : (4 * n_take < nl ? L[4 * n_take].start : L[4 * n_take - 1].stop() + 1);
                else next_cur = n_take == 0 ? cur : (n_take < n_all ? L[R[n_take].first_line - 1].start : after_all);
                // a tail that is not a whole record (or lacks its last '\n') is the serial parser's business
                if (!full && !anomaly && at_eof && next_cur < size) anomaly = true;
            }
            if (anomaly || full || at_eof) break;
            window *= 2; // the window ended before the batch was full
        }
        if (anomaly) {
            if (bgzf) z_fill((size_t)1 << 62); // the serial parser continues in memory: inflate the rest of the members first
            serial.reset(new FastxReader(filename, format, data + cur, size - cur, gzf)); // (gzip: continues with the rest of the stream)
            return from_serial(max_bases, max_reads, ascii, offsets);
        }
        const size_t n = n_take;
        offsets.resize(n + 1);
        for (size_t r = 0; r < n; r++) offsets[r + 1] = offsets[r] + recs[r].len;
        ascii.resize((size_t)offsets[n]); // (not cleared first: only growth beyond the previous batch is zero-filled)
        char* out = ascii.data();
        const int64_t* off = offsets.data();
        const Line* L = lines.data();
        const Rec* R = recs.data();
        const char* d = data;
        parallel_for(n, [=](size_t a, size_t b, size_t) {
            for (size_t r = a; r < b; r++) {
                char* o = out + off[r];
                for (size_t l = R[r].first_line; l < R[r].first_line + R[r].n_lines; l++) {
                    memcpy(o, d + L[l].start, L[l].stop() - L[l].start);
                    o += L[l].stop() - L[l].start;
                }
            }
        });
        if (offsets[n] > 0) file_bytes_per_base = std::min(8.0, std::max(1.0, 1.03 * (double)(next_cur - cur) / (double)offsets[n]));
        cur = next_cur;
        return (int64_t)n;
    }
};

} // namespace sbwt_b200
