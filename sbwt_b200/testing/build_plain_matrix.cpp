// build_plain_matrix.cpp -- TEST / BENCH INFRASTRUCTURE, not part of the product path.
//
// Index construction is out of scope for this repository (SURVEY.md section 2:
// KMC, EM-sort, NodeBOSS constructors), but the GPU box has no copy of the
// reference, so the synthetic indexes named by BASELINE.json have to be
// produced from inside the repo. This is an independent, in-memory, sort-based
// constructor of the *canonical* plain-matrix SBWT:
//
//   nodes   = distinct k-mers  +  every proper prefix ($-padded) of every k-mer
//             that has no predecessor k-mer, in colex order, root first;
//   column  = out-edge labels of the node, recorded only at the first node of
//             its suffix group (nodes sharing the last k-1 characters);
//   file    = the reference's serialized layout,
//             including sdsl's rank_support_v5 directories and the p-mer
//             interval table, byte-for-byte what `sbwt build` writes.
//
// tests/test_builder.py byte-compares its output with files written by the
// reference's own constructors (KMC-based `sbwt build` and the in-memory
// constructor compiled from the reference tree).
//
// usage: build_plain_matrix -i in.fna[,in2.fna...] -o out.sbwt -k K [-p P]
//            [--no-streaming-support] [--add-reverse-complements] [-t threads]
//   input: FASTA (multi-line allowed), or with --raw a file of bare sequence
//   bytes with '\n' between sequences. Case-insensitive; any byte outside ACGT
//   breaks k-mers.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

typedef unsigned __int128 u128;

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static double t_start;
static bool verbose = true;
static void logmsg(const char* msg) {
    if (verbose) fprintf(stderr, "[build %.2fs] %s\n", now() - t_start, msg);
}

static std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(1); }
    std::vector<uint8_t> buf;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)sz);
    if (sz && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "short read %s\n", path.c_str()); exit(1); }
    fclose(f);
    return buf;
}

static inline int code_of(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// A key holds a k-mer with character j at digit j (2 bits each), so the LAST
// character is the most significant digit and integer order == colex order.
template <class Key>
struct Builder {
    int k;
    bool rc;
    std::vector<Key> kmers;

    void add_stream(const uint8_t* s, size_t n, bool fasta) {
        Key fwd = 0, rev = 0;
        const Key mask = (k * 2 == (int)sizeof(Key) * 8) ? ~(Key)0 : (((Key)1 << (2 * k)) - 1);
        int run = 0;
        bool in_header = false;
        for (size_t i = 0; i < n; i++) {
            uint8_t ch = s[i];
            if (fasta) {
                if (in_header) { if (ch == '\n') in_header = false; continue; }
                if (ch == '>') { in_header = true; run = 0; continue; }
                if (ch == '\n' || ch == '\r') continue; // multi-line sequence
            } else if (ch == '\n') { run = 0; continue; }
            int c = code_of(ch);
            if (c > 3) { run = 0; continue; }
            fwd = (fwd >> 2) | ((Key)c << (2 * (k - 1)));
            rev = ((rev << 2) | (Key)(3 - c)) & mask;
            if (++run >= k) {
                kmers.push_back(fwd);
                if (rc) kmers.push_back(rev);
            }
        }
    }
};

template <class Key>
static void sort_unique(std::vector<Key>& v) {
#ifdef _OPENMP
    __gnu_parallel::sort(v.begin(), v.end());
#else
    std::sort(v.begin(), v.end());
#endif
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

struct BitVec {
    uint64_t n = 0;
    std::vector<uint64_t> w;
    void resize(uint64_t bits) { n = bits; w.assign((bits + 63) / 64 + 1, 0); } // +1 padding word, not serialized
    inline void set(uint64_t i) { w[i >> 6] |= 1ULL << (i & 63); }
    uint64_t nwords() const { return (n + 63) / 64; }
};

// Cumulative popcount per word, for the builder's own rank (precalc table).
struct RankHelper {
    const BitVec* bv;
    std::vector<uint64_t> cum; // cum[i] = ones in words [0,i)
    void init(const BitVec* b) {
        bv = b;
        cum.resize(b->w.size() + 1);
        uint64_t s = 0;
        for (size_t i = 0; i < b->w.size(); i++) { cum[i] = s; s += (uint64_t)__builtin_popcountll(b->w[i]); }
        cum[b->w.size()] = s;
    }
    inline int64_t rank(int64_t pos) const {
        uint64_t wi = (uint64_t)pos >> 6, off = (uint64_t)pos & 63;
        return (int64_t)(cum[wi] + (off ? (uint64_t)__builtin_popcountll(bv->w[wi] & ((1ULL << off) - 1)) : 0));
    }
};

static void put(FILE* f, const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) { fprintf(stderr, "write failed\n"); exit(1); }
}
static void put_i64(FILE* f, int64_t x) { put(f, &x, 8); }
static void put_string(FILE* f, const char* s) { put_i64(f, (int64_t)strlen(s)); put(f, s, strlen(s)); }
static void put_bitvec(FILE* f, const BitVec& b) {
    uint64_t n = b.n;
    put(f, &n, 8);
    put(f, b.w.data(), b.nwords() * 8);
}

// The directory sdsl's rank_support_v5<1,1> constructor produces for this bit
// vector: per 2048-bit superblock one absolute count and five 12-bit fields
// (11 used) holding the counts of the first 6,12,..,30 words, at shifts 48..0.
// A field exists only if the vector has at least that many words.
static std::vector<uint64_t> v5_directory(const BitVec& b) {
    uint64_t W = b.nwords();
    if (b.n == 0) return std::vector<uint64_t>(2, 0);
    uint64_t nsb = W / 32 + 1;
    std::vector<uint64_t> bb(nsb * 2, 0);
    uint64_t total = 0;
    for (uint64_t s = 0; s < nsb; s++) {
        bb[2 * s] = total;
        uint64_t second = 0, sum = 0;
        for (uint64_t j = 0; j < 32; j++) {
            uint64_t wi = 32 * s + j;
            if (j && j % 6 == 0 && W >= wi) second |= sum << (60 - 12 * (j / 6));
            if (wi < W) sum += (uint64_t)__builtin_popcountll(b.w[wi]);
        }
        bb[2 * s + 1] = second;
        total += sum;
    }
    return bb;
}

template <class Key>
static int run(int k, int p, bool streaming, bool rc, bool raw, const std::vector<std::string>& inputs,
               const std::string& out_path) {
    Builder<Key> B;
    B.k = k;
    B.rc = rc;
    for (const std::string& path : inputs) {
        std::vector<uint8_t> buf = read_file(path);
        B.add_stream(buf.data(), buf.size(), !raw);
    }
    logmsg("k-mers extracted");
    std::vector<Key>& K = B.kmers;
    sort_unique(K);
    const int64_t nk = (int64_t)K.size();
    logmsg("k-mers sorted");

    const int top = 2 * (k - 1);
    const Key pmask = ((Key)1 << top) - 1; // digits 0..k-2: the (k-1)-prefix, aligned like the suffix key (K >> 2)

    // Block boundaries by last character.
    int64_t blk[5];
    blk[0] = 0;
    for (int c = 0; c < 4; c++) {
        Key lim = (c == 3) ? ~(Key)0 : (((Key)(c + 1)) << top);
        blk[c + 1] = (c == 3) ? nk : (int64_t)(std::lower_bound(K.begin(), K.end(), lim) - K.begin());
    }

    // edges[i] (4 bits) is meaningful at suffix-group starts; has_in marks k-mers with a predecessor.
    std::vector<uint8_t> edges((size_t)nk, 0), has_in((size_t)nk, 0);
    auto scan_char = [&](int c) {
        int64_t i = blk[c], g = 0;
        while (i < blk[c + 1] && g < nk) {
            // advance g to a group start
            Key s = K[g] >> 2;
            Key pfx = K[i] & pmask;
            if (pfx == s) {
                has_in[i] = 1;
                edges[g] |= (uint8_t)(1 << c); // only this thread writes bit c... but bytes are shared:
                i++;
                // move to next group
                int64_t g2 = g + 1;
                while (g2 < nk && (K[g2] >> 2) == s) g2++;
                g = g2;
            } else if (pfx < s) {
                i++; // no predecessor: a source k-mer
            } else {
                int64_t g2 = g + 1;
                while (g2 < nk && (K[g2] >> 2) == s) g2++;
                g = g2;
            }
        }
    };
    // (edges[] bytes are shared between characters, so the four scans run one after another.)
    for (int c = 0; c < 4; c++) scan_char(c);
    logmsg("edges computed");

    // Dummy nodes: (padded key, length, edge flags).
    struct Dummy { Key pk; int len; uint8_t e; };
    std::vector<Dummy> D;
    for (int64_t i = 0; i < nk; i++) {
        if (has_in[i]) continue;
        Key z = K[i];
        for (int j = 0; j < k; j++) {
            Key d = (j == 0) ? (Key)0 : (z & (((Key)1 << (2 * j)) - 1));
            Key pk = (j == 0) ? (Key)0 : (d << (2 * (k - j)));
            uint8_t e = (uint8_t)(1 << (int)((z >> (2 * j)) & 3));
            D.push_back({pk, j, e});
        }
    }
    D.push_back({(Key)0, 0, 0}); // the root always exists, even when no k-mer is a source
    std::sort(D.begin(), D.end(), [](const Dummy& a, const Dummy& b) { return a.pk != b.pk ? a.pk < b.pk : a.len < b.len; });
    {
        size_t o = 0;
        for (size_t i = 0; i < D.size(); i++) {
            if (o && D[o - 1].pk == D[i].pk && D[o - 1].len == D[i].len) D[o - 1].e |= D[i].e;
            else D[o++] = D[i];
        }
        D.resize(o);
    }
    const int64_t nd = (int64_t)D.size();
    const int64_t n = nk + nd;
    logmsg("dummies sorted");

    BitVec bits[4], sgs;
    for (int c = 0; c < 4; c++) bits[c].resize((uint64_t)n);
    sgs.resize((uint64_t)n);
    {
        int64_t i = 0, d = 0, col = 0;
        while (i < nk || d < nd) {
            bool take_dummy = d < nd && (i >= nk || D[d].pk <= K[i]);
            if (take_dummy) {
                for (int c = 0; c < 4; c++) if (D[d].e >> c & 1) bits[c].set((uint64_t)col);
                sgs.set((uint64_t)col);
                d++;
            } else {
                bool start = (i == 0) || ((K[i] >> 2) != (K[i - 1] >> 2));
                if (start) {
                    sgs.set((uint64_t)col);
                    for (int c = 0; c < 4; c++) if (edges[i] >> c & 1) bits[c].set((uint64_t)col);
                }
                i++;
            }
            col++;
        }
    }
    logmsg("columns emitted");

    RankHelper R[4];
    for (int c = 0; c < 4; c++) R[c].init(&bits[c]);
    int64_t C[4];
    C[0] = 1;
    for (int c = 0; c < 3; c++) C[c + 1] = C[c] + R[c].rank(n);

    if (p > k) p = k;
    std::vector<int64_t> precalc;
    if (p > 0) {
        uint64_t np = 1ULL << (2 * p);
        precalc.resize(2 * np);
        for (uint64_t idx = 0; idx < np; idx++) {
            int64_t l = 0, r = n - 1;
            for (int j = 0; j < p && l != -1; j++) {
                int c = (int)((idx >> (2 * j)) & 3);
                l = C[c] + R[c].rank(l);
                r = C[c] + R[c].rank(r + 1) - 1;
                if (l > r) l = r = -1;
            }
            precalc[2 * idx] = l;
            precalc[2 * idx + 1] = r;
        }
    }
    logmsg("precalc done");

    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot open %s for writing\n", out_path.c_str()); return 1; }
    put_string(f, "plain-matrix");
    put_string(f, "v0.1");
    for (int c = 0; c < 4; c++) put_bitvec(f, bits[c]);
    for (int c = 0; c < 4; c++) {
        std::vector<uint64_t> bb = v5_directory(bits[c]);
        uint64_t nbits = bb.size() * 64;
        put(f, &nbits, 8);
        put(f, bb.data(), bb.size() * 8);
    }
    if (streaming) put_bitvec(f, sgs);
    else { uint64_t z = 0; put(f, &z, 8); }
    put_i64(f, 32);
    put(f, C, 32);
    put_i64(f, (int64_t)precalc.size() * 8);
    put(f, precalc.data(), precalc.size() * 8);
    put_i64(f, p);
    put_i64(f, n);
    put_i64(f, nk);
    put_i64(f, k);
    fclose(f);
    logmsg("written");
    printf("{\"n_kmers\": %lld, \"n_nodes\": %lld, \"k\": %d, \"precalc_k\": %d, \"streaming\": %s}\n", (long long)nk,
           (long long)n, k, p, streaming ? "true" : "false");
    return 0;
}

int main(int argc, char** argv) {
    t_start = now();
    std::vector<std::string> inputs;
    std::string out;
    int k = 0, p = 8, threads = 0;
    bool streaming = true, rc = false, raw = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "-i") {
            std::string v = next();
            size_t pos = 0;
            while (true) {
                size_t q = v.find(',', pos);
                inputs.push_back(v.substr(pos, q == std::string::npos ? q : q - pos));
                if (q == std::string::npos) break;
                pos = q + 1;
            }
        } else if (a == "-o") out = next();
        else if (a == "-k") k = atoi(next().c_str());
        else if (a == "-p") p = atoi(next().c_str());
        else if (a == "-t") threads = atoi(next().c_str());
        else if (a == "--no-streaming-support") streaming = false;
        else if (a == "--add-reverse-complements") rc = true;
        else if (a == "--raw") raw = true;
        else if (a == "-q") verbose = false;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    if (inputs.empty() || out.empty() || k < 1 || k > 64 || p < 0 || p > 14) {
        fprintf(stderr, "usage: build_plain_matrix -i in.fna[,..] -o out.sbwt -k K(1..64) [-p P(0..14)] [--no-streaming-support] [--add-reverse-complements] [--raw] [-t threads] [-q]\n");
        return 1;
    }
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    if (k <= 32) return run<uint64_t>(k, p, streaming, rc, raw, inputs, out);
    return run<u128>(k, p, streaming, rc, raw, inputs, out);
}
