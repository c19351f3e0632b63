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

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

namespace sbwt_b200 {

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
    size_t pos = 0, size = 0;
    bool is_eof = false; // true once a get() ran past the end (Buffered_ifstream::eof)

    bool refill() {
        if (gz) {
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
    inline void getline_append(std::vector<char>* out) {
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
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;
    ~FastxReader() {
        if (fp) fclose(fp);
        if (gzf) gzclose(gzf);
    }

    // Appends the next read's bases to `ascii`; returns its length, 0 at end of file.
    int64_t next_read(std::vector<char>& ascii) {
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
    int64_t next_batch(int64_t max_bases, int64_t max_reads, std::vector<char>& ascii, std::vector<int64_t>& offsets) {
        ascii.clear();
        offsets.clear();
        offsets.push_back(0);
        while ((int64_t)offsets.size() - 1 < max_reads && (int64_t)ascii.size() < max_bases) {
            if (next_read(ascii) == 0) break;
            offsets.push_back((int64_t)ascii.size());
        }
        return (int64_t)offsets.size() - 1;
    }
};

} // namespace sbwt_b200
