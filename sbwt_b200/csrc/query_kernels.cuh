// query_kernels.cuh -- the other read-only queries of the index, batched:
//   SBWT::update_sbwt_interval  include/sbwt/SBWT.hh:423-437
//   SBWT::partial_search        SBWT.hh:526-537
//   SBWT::forward               SBWT.hh:369-381
//   SubsetMatrixRank::contains  include/sbwt/SubsetMatrixRank.hh:39-48
//   SBWT::get_kmer              SBWT.hh:701-725
//   SBWT::ascii_export_sets     SBWT.hh:750-773
// One thread per query on the classic sectors (every index has them); each answers with exactly the arithmetic of the
// reference function it replaces. These are not the hot path: one launch per batch instead of one device round trip
// per character is what they are for.
#pragma once

#include "device_index.cuh"

namespace sbwt_b200 {

// get_char_idx (SBWT.hh:49-57): case-sensitive
__device__ __forceinline__ int code_exact(uint8_t ch) { return ch == 'A' ? 0 : (ch == 'C' ? 1 : (ch == 'G' ? 2 : (ch == 'T' ? 3 : -1))); }

// C[c] + rank_c(pos), 0 <= pos <= n_nodes
template <bool WIDE>
__device__ __forceinline__ int64_t classic_lf(const DeviceIndexView& ix, int64_t pos, int c) {
    const BlockPos bp = split_pos<WIDE>(pos);
    const Sector s = ld_sector(sector_addr<WIDE>(ix, bp.blk, c));
    return lf_value<WIDE>(ix, s, bp.blk, bp.off, c);
}

template <bool WIDE>
__device__ __forceinline__ uint32_t classic_bit(const DeviceIndexView& ix, int64_t pos, int c) {
    const BlockPos bp = split_pos<WIDE>(pos);
    const Sector s = ld_sector(sector_addr<WIDE>(ix, bp.blk, c));
    return sector_bit(s, bp.off);
}

constexpr int kExtendUpdate = 0;  // update_sbwt_interval: raw bytes, {-1,-1} passes through, failure gives {-1,-1}
constexpr int kExtendPartial = 1; // partial_search: toupper, starts from {0, n-1}, failure keeps the last interval

// string t = ascii[offsets[t] - offsets[0] .. offsets[t+1] - offsets[0]); l/r in and out; matched[t] = characters consumed
template <bool WIDE>
__global__ void __launch_bounds__(256) extend_kernel(const DeviceIndexView ix, const uint8_t* __restrict__ ascii,
                                                     const int64_t* __restrict__ offsets, int64_t n, int mode,
                                                     int64_t* __restrict__ l, int64_t* __restrict__ r, int64_t* __restrict__ matched) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint8_t* s = ascii + (offsets[t] - offsets[0]);
    const int64_t len = offsets[t + 1] - offsets[t];
    int64_t lo = mode == kExtendPartial ? 0 : l[t], hi = mode == kExtendPartial ? ix.n_nodes - 1 : r[t];
    int64_t i = 0;
    if (!(mode == kExtendUpdate && lo == -1)) { // SBWT.hh:424
        for (; i < len; i++) {
            uint8_t ch = s[i];
            if (mode == kExtendPartial && ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 32); // SBWT.hh:530
            const int c = code_exact(ch);
            int64_t nl = -1, nr = -1;
            if (c >= 0) {
                nl = classic_lf<WIDE>(ix, lo, c);
                nr = classic_lf<WIDE>(ix, hi + 1, c) - 1;
            }
            if (c < 0 || nl > nr) { // invalid character / not found (SBWT.hh:428,433)
                if (mode == kExtendUpdate) lo = hi = -1;
                break;
            }
            lo = nl;
            hi = nr;
        }
    }
    l[t] = lo;
    r[t] = hi;
    if (matched) matched[t] = i;
}

// SBWT.hh:369-381: to the start of node's suffix group, then the edge labelled c; -1 if there is none
template <bool WIDE>
__global__ void __launch_bounds__(256) forward_kernel(const DeviceIndexView ix, const int64_t* __restrict__ nodes,
                                                      const char* __restrict__ chars, int64_t n, int64_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t g = nodes[t];
    while (true) { // the first node is always marked
        const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
        if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
        g = (g & ~31ll) - 1;
    }
    const int c = code_exact((uint8_t)chars[t]);
    if (c < 0) { out[t] = -1; return; } // rank of any other byte is 0 at both positions (SubsetMatrixRank.hh:36)
    const int64_t v = classic_lf<WIDE>(ix, g, c);
    out[t] = classic_bit<WIDE>(ix, g, c) ? v : -1; // r1 == r2 <=> the bit at g is clear
}

template <bool WIDE>
__global__ void __launch_bounds__(256) contains_kernel(const DeviceIndexView ix, const int64_t* __restrict__ pos,
                                                       const char* __restrict__ chars, int64_t n, uint8_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = code_exact((uint8_t)chars[t]);
    out[t] = c < 0 ? 0 : (uint8_t)classic_bit<WIDE>(ix, pos[t], c);
}

// SBWT.hh:701-725: the label of node colex_rank, '$'-padded; the backward step is the reference's search for the largest p
// with rank_c(p) <= rank (here a bisection over [0, n_nodes] on the same monotone predicate)
template <bool WIDE>
__global__ void __launch_bounds__(256) get_kmer_kernel(const DeviceIndexView ix, const int64_t* __restrict__ ranks, int64_t n,
                                                       int64_t C0, int64_t C1, int64_t C2, int64_t C3, char* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t node = ranks[t];
    const int64_t Cc[4] = {C0, C1, C2, C3};
    char* buf = out + t * ix.k;
    for (int i = 0; i < ix.k; i++) {
        if (node == 0) {
            buf[ix.k - 1 - i] = '$';
        } else {
            int c = 0;
            while (c + 1 < 4 && node >= Cc[c + 1]) c++;
            buf[ix.k - 1 - i] = "ACGT"[c];
            int64_t lo = 0, hi = ix.n_nodes; // invariant: C[c] + rank_c(lo) <= node
            while (lo < hi) {
                const int64_t mid = lo + (hi - lo + 1) / 2;
                if (classic_lf<WIDE>(ix, mid, c) <= node) lo = mid;
                else hi = mid - 1;
            }
            node = lo;
        }
    }
}

// SBWT.hh:750-773, two passes: text bytes per 32 columns, then (after an exclusive scan) the text itself
__device__ __forceinline__ void export_planes(const DeviceIndexView& ix, int64_t w, uint32_t (&pl)[4]) {
    const int64_t b = w / kPayloadWords;
    const int wi = (int)(w - b * kPayloadWords);
#pragma unroll
    for (int c = 0; c < 4; c++) pl[c] = b < ix.n_blocks ? ix.sectors[4 * b + c].w[1 + wi] : 0u;
}

__global__ void __launch_bounds__(256) export_len_kernel(const DeviceIndexView ix, int64_t n_words, int64_t* __restrict__ len) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t pl[4];
    export_planes(ix, w, pl);
    const int64_t first = w * 32;
    const uint32_t valid = ix.n_nodes - first >= 32 ? 0xFFFFFFFFu : ((1u << (uint32_t)(ix.n_nodes - first)) - 1u);
    const uint32_t any = pl[0] | pl[1] | pl[2] | pl[3];
    // every column writes its characters, an empty one a single '$'
    len[w] = __popc(pl[0] & valid) + __popc(pl[1] & valid) + __popc(pl[2] & valid) + __popc(pl[3] & valid) + __popc(~any & valid);
}

__global__ void __launch_bounds__(256) export_emit_kernel(const DeviceIndexView ix, int64_t n_words, const int64_t* __restrict__ off,
                                                          char* __restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t pl[4];
    export_planes(ix, w, pl);
    const int64_t first = w * 32;
    const int cols = ix.n_nodes - first >= 32 ? 32 : (int)(ix.n_nodes - first);
    char* p = out + off[w];
    for (int j = 0; j < cols; j++) {
        char* start = p;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if ((pl[c] >> j) & 1u) *p++ = "ACGT"[c];
        if (p == start) *p++ = '$';
        else p[-1] = (char)(p[-1] + 32); // the last character of a set is lower-cased
    }
}

} // namespace sbwt_b200
