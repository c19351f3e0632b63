// long_kmer_kernels.cuh -- k > 64: the queries of walk_kernel.cuh without its register-held k-mer windows.
//
// The reference's search has no limit on k (SBWT.hh:390-437 loops over the k characters; only the construction side is
// sized by MAX_KMER_LENGTH, at most 255). The phase-sorted walk kernel keeps a k-mer in 2 or 4 registers, which is
// what bounds it to k <= 64; these two kernels take the characters from the packed reads as they go and follow the
// reference literally on the classic sectors:
//   long_search_kernel     one lane per k-mer: the k interval steps of SBWT::search from [0, n-1] (SBWT.hh:390-415, 423-437;
//                          the reference's table jump over the first p characters gives the same interval)
//   long_streaming_kernel  one thread per work item (a window of a read's k-mers): SBWT::streaming_search's loop
//                          (SBWT.hh:545-581): from scratch after a miss, else walk back to the suffix-group start and
//                          take one step; a window's first k-mer is searched from scratch
// Correct for every k >= 1; meant for the long k-mers the hot kernel does not take, not tuned.
#pragma once

#include "aux_kernels.cuh"
#include "query_kernels.cuh"
#include "walk_kernel.cuh"

namespace sbwt_b200 {

__device__ __forceinline__ int long_code(const uint32_t* __restrict__ codes, uint32_t pos) {
    return (int)((__ldg(codes + (pos >> 4)) >> ((pos & 15u) * 2u)) & 3u);
}

// SBWT::search of the k-mer at base b: the colex rank, or -1 (the caller has checked that no base of it is invalid)
template <bool WIDE>
__device__ __forceinline__ int64_t long_search(const DeviceIndexView& ix, const uint32_t* __restrict__ codes, uint32_t b, uint32_t k) {
    int64_t l = 0, r = ix.n_nodes - 1;
    for (uint32_t j = 0; j < k; j++) {
        const int c = long_code(codes, b + j);
        const int64_t nl = classic_lf<WIDE>(ix, l, c), nr = classic_lf<WIDE>(ix, r + 1, c) - 1;
        if (nl > nr) return -1; // SBWT.hh:433
        l = nl;
        r = nr;
    }
    return l; // a k-mer's interval is a singleton (SBWT.hh:410-413)
}

template <bool OUT32>
__device__ __forceinline__ void long_store(const WalkParams& P, uint32_t o, int64_t v) {
    if (OUT32) P.out32[o] = (int32_t)v;
    else P.out[o] = v;
}

template <bool WIDE, bool OUT32>
__global__ void __launch_bounds__(256) long_search_kernel(const WalkParams P, uint32_t k) {
    const int64_t n_items = *P.n_items;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = warp0; it < n_items; it += n_warps) {
        const WalkItem w = P.items[it];
        if ((uint32_t)lane >= w.cnt) continue;
        const uint32_t b = w.base + lane;
        const int64_t bad = next_set_bit(P.invalid, b, (int64_t)b + k);
        long_store<OUT32>(P, w.out + lane, bad < (int64_t)b + k ? -1 : long_search<WIDE>(P.ix, P.codes, b, k));
    }
}

template <bool WIDE, bool OUT32>
__global__ void __launch_bounds__(128) long_streaming_kernel(const WalkParams P, uint32_t k) {
    const int64_t n_items = *P.n_items;
    const DeviceIndexView& ix = P.ix;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (int64_t)gridDim.x * blockDim.x) {
        const WalkItem w = P.items[it];
        const int64_t end = (int64_t)w.base + w.cnt + k - 1; // bases covered by the item's k-mers: [base, end)
        int64_t bad = next_set_bit(P.invalid, w.base, end);
        int64_t prev = -1;
        for (uint32_t i = 0; i < w.cnt; i++) {
            const int64_t b = (int64_t)w.base + i;
            if (bad < b) bad = next_set_bit(P.invalid, b, end);
            int64_t ans;
            if (i == 0 || prev < 0) { // from scratch (SBWT.hh:553, 557-559)
                ans = bad < b + k ? -1 : long_search<WIDE>(ix, P.codes, (uint32_t)b, k);
            } else {                  // to the start of the suffix group, then one step (SBWT.hh:561-575)
                int64_t g = prev;
                while (true) { // the first column is always marked
                    const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
                    if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
                    g = (g & ~31ll) - 1;
                }
                if (bad == b + k - 1) ans = -1; // the new character is not one of ACGT (SBWT.hh:568)
                else {
                    const int c = long_code(P.codes, (uint32_t)(b + k - 1));
                    const int64_t nl = classic_lf<WIDE>(ix, g, c), nr = classic_lf<WIDE>(ix, g + 1, c) - 1;
                    ans = nl == nr ? nl : -1; // SBWT.hh:573-574
                }
            }
            long_store<OUT32>(P, w.out + i, ans);
            prev = ans;
        }
    }
}

} // namespace sbwt_b200
