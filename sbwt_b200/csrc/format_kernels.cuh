// format_kernels.cuh -- print_vector (src/CLI/sbwt_search.cpp:21-43) on the device.
//
// The reference writes, per read, "<value> " for every k-mer and then '\n'; -1 prints as "-1", a
// value that is neither -1 nor positive prints as an empty field (its digit loop runs while x > 0).
// Here the text of a whole batch is produced in HBM and leaves the device as bytes:
//   fmt_len_kernel    bytes of each read's line                     (8 B read per value)
//   exclusive scan    line start offsets (the scans of aux_kernels.cuh)
//   fmt_emit_kernel   digits -> per-warp shared-memory staging -> 16-byte coalesced stores
// A warp owns kFmtReadsPerWarp consecutive reads, hence one contiguous span of the text; only the
// first and last partial 16-byte blocks of that span are written bytewise.
#pragma once

#include <cstdint>

namespace sbwt_b200 {

constexpr int kFmtThreads = 128;
constexpr int kFmtWarps = kFmtThreads / 32;
constexpr int kFmtReadsPerWarp = 8;
constexpr int kFmtBufBytes = 1536; // per warp: < kFmtFlushAt pending + one chunk of 32 values (<= 32 * 21 + 1 bytes)
constexpr int kFmtFlushAt = 512;

__constant__ uint64_t c_pow10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                     1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                     100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                                     1000000000000000000ull, 10000000000000000000ull};

// bytes print_vector writes for one value, the trailing ' ' included
__device__ __forceinline__ int fmt_value_len(int64_t x) {
    if (x == -1) return 3;
    if (x <= 0) return 1;
    const uint64_t v = (uint64_t)x;
    const int t = ((64 - __clzll((long long)v)) * 1233) >> 12; // floor(log10 v) or one less
    return t + (v >= c_pow10[t] ? 1 : 0) + 1;
}

template <typename T>
__global__ void __launch_bounds__(kFmtThreads) fmt_len_kernel(const T* __restrict__ vals, const int64_t* __restrict__ voff,
                                                              const int64_t* __restrict__ vtotal, int64_t n_reads,
                                                              int64_t* __restrict__ tlen) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kFmtThreads + threadIdx.x) >> 5;
    const int64_t r0 = warp * kFmtReadsPerWarp;
    for (int i = 0; i < kFmtReadsPerWarp; i++) {
        const int64_t r = r0 + i;
        if (r >= n_reads) break;
        const int64_t v0 = voff[r], v1 = r + 1 < n_reads ? voff[r + 1] : *vtotal;
        int64_t s = 0;
        for (int64_t q = v0 + lane; q < v1; q += 32) s += fmt_value_len((int64_t)__ldcs(vals + q));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if (lane == 0) tlen[r] = s + 1; // + '\n'
    }
}

// toff = exclusive scan of tlen, *ttotal its total. Writes nothing when the text would not fit.
template <typename T>
__global__ void __launch_bounds__(kFmtThreads) fmt_emit_kernel(const T* __restrict__ vals, const int64_t* __restrict__ voff,
                                                               const int64_t* __restrict__ vtotal,
                                                               const int64_t* __restrict__ toff,
                                                               const int64_t* __restrict__ ttotal, int64_t n_reads,
                                                               char* __restrict__ text, int64_t capacity) {
    __shared__ __align__(16) unsigned char sbuf[kFmtWarps][kFmtBufBytes];
    if (*ttotal > capacity) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kFmtThreads + threadIdx.x) >> 5;
    const int64_t r0 = warp * kFmtReadsPerWarp;
    if (r0 >= n_reads) return;
    const int64_t r1 = r0 + kFmtReadsPerWarp < n_reads ? r0 + kFmtReadsPerWarp : n_reads;
    unsigned char* buf = sbuf[threadIdx.x >> 5];

    // buf[0] stands for the 16-byte aligned address gbase; the first `hole` bytes of the first block are not ours
    char* const gp = text + toff[r0];
    char* gbase = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(gp) & ~(uintptr_t)15);
    int fill = (int)(gp - gbase);
    int hole = fill;

    auto flush_blocks = [&]() { // stores the whole 16-byte blocks, keeps the remainder at the start of buf
        const int nblk = fill >> 4;
        for (int j = lane; j < nblk; j += 32) {
            if (j == 0 && hole) {
                for (int b = hole; b < 16; b++) gbase[b] = (char)buf[b];
            } else {
                __stcs(reinterpret_cast<uint4*>(gbase) + j, reinterpret_cast<const uint4*>(buf)[j]);
            }
        }
        __syncwarp();
        const int rem = fill & 15;
        unsigned char t = 0;
        if (lane < rem) t = buf[16 * nblk + lane];
        __syncwarp();
        if (nblk) {
            if (lane < rem) buf[lane] = t;
            gbase += 16 * nblk;
            fill = rem;
            hole = 0;
        }
        __syncwarp();
    };

    for (int64_t r = r0; r < r1; r++) {
        const int64_t v0 = voff[r], v1 = r + 1 < n_reads ? voff[r + 1] : *vtotal;
        if (v0 >= v1) { // a read shorter than k: an empty line
            if (lane == 0) buf[fill] = '\n';
            fill += 1;
            __syncwarp();
            if (fill >= kFmtFlushAt) flush_blocks();
            continue;
        }
        for (int64_t qb = v0; qb < v1; qb += 32) {
            const int64_t q = qb + lane;
            const bool active = q < v1;
            const int64_t x = active ? (int64_t)__ldcs(vals + q) : 0;
            const bool last = active && q == v1 - 1;
            const int vlen = active ? fmt_value_len(x) : 0;
            const int len = vlen + (last ? 1 : 0);
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (active) {
                unsigned char* p = buf + fill + incl - len;
                p[vlen - 1] = ' ';
                if (x == -1) {
                    p[0] = '-';
                    p[1] = '1';
                } else if (x > 0) {
                    unsigned char* e = p + vlen - 1;
                    if ((uint64_t)x >> 32) {
                        uint64_t v = (uint64_t)x;
                        do { const uint64_t d = v / 10; *--e = (unsigned char)('0' + (int)(v - d * 10)); v = d; } while (v);
                    } else {
                        uint32_t v = (uint32_t)x;
                        do { const uint32_t d = v / 10; *--e = (unsigned char)('0' + (v - d * 10)); v = d; } while (v);
                    }
                }
                if (last) p[vlen] = '\n';
            }
            fill += total;
            __syncwarp();
            if (fill >= kFmtFlushAt) flush_blocks();
        }
    }
    flush_blocks();
    // the tail (and, for a span inside one block, everything): bytes [hole, fill)
    for (int b = hole + lane; b < fill; b += 32) gbase[b] = (char)buf[b];
}

} // namespace sbwt_b200
