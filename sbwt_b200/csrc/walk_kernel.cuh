// walk_kernel.cuh -- the colex-interval walk on the device (sm_100a).
//
// Replaces SBWT::search (include/sbwt/SBWT.hh:390-415), SBWT::update_sbwt_interval
// (SBWT.hh:423-437) and SBWT::streaming_search (SBWT.hh:545-581) together with the per-read
// loops of src/CLI/sbwt_search.cpp:45-91.
//
// Work item = up to `window` consecutive k-mers of one read (a whole read when it is short).
// One LANE owns one item at a time and is a small state machine; every trip of the warp loop
// each lane performs at most one memory round trip of whatever kind it needs next:
//
//   START   first p characters -> row of the precalc table (kmer_prefix_precalc, SBWT.hh:404)
//   WALK    one interval step  [l,r] -> [C[c]+rank_c(l), C[c]+rank_c(r+1)-1]  (two sectors,
//           one when l and r+1 fall into the same 224-column block)
//   STREAM  previous k-mer was found at column `col`: one sector holding bit_c(col) and
//           C[c]+rank_c(col). In a reference-built index only suffix-group starts carry edges
//           (SURVEY.md section 8(a) note 7), so bit_c(col)==1 already proves col is the group
//           start and the answer is C[c]+rank_c(col); otherwise the literal walk-back over
//           suffix_group_starts (SBWT.hh:562-563) runs on a slow path.
//
// so lanes that sit in a long from-scratch walk, lanes that stream along a matching read and
// lanes that just fetched a new read all keep one or two independent sector loads in flight --
// the pointer chase is hidden by the ~2000 resident lanes per SM, not by ILP inside a lane.
// Finished lanes refill from the warp's own contiguous item range (ballot/popc, no atomics).
#pragma once

#include "device_index.cuh"

namespace sbwt_b200 {

struct WalkParams {
    DeviceIndexView ix;
    const uint64_t* codes;    // 2-bit bases, 32 per word
    const uint32_t* invalid;  // 1 bit per base, 32 per word
    const int64_t* item_base; // global base index of the item's first k-mer
    const int64_t* item_out;  // index of its first result in `out`
    const int32_t* item_cnt;  // number of k-mers in the item
    const int64_t* n_items;   // device scalar
    int64_t* out;
    unsigned long long* stats; // [lookups, hits, rank_ops, sectors] (COUNT only)
};

enum : int { M_NEED = 0, M_START = 1, M_WALK = 2, M_STREAM = 3, M_DONE = 4 };
enum : int { K_NONE = 0, K_PRE = 1, K_WALK = 2, K_STREAM = 3 };

// Sliding window over the packed read: base t of the window (t = 0 is the first character of
// the current k-mer) sits at bits [2(t%32), 2(t%32)+2) of b[t/32]; v holds the invalid flags.
template <int NW>
struct Window {
    uint64_t b[NW];
    uint32_t v[NW];
    uint64_t nb; // packed word holding the next base to shift in
    uint32_t nv;
    int64_t np;  // global index of that base

    __device__ __forceinline__ void init(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ invalid, int64_t g) {
        const int64_t wi = g >> 5;
        const int sh = (int)(g & 31);
        uint64_t lo = codes[wi];
        uint32_t vlo = invalid[wi];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint64_t hi = codes[wi + w + 1];
            const uint32_t vhi = invalid[wi + w + 1];
            b[w] = sh ? ((lo >> (2 * sh)) | (hi << (64 - 2 * sh))) : lo;
            v[w] = __funnelshift_r(vlo, vhi, sh);
            lo = hi;
            vlo = vhi;
        }
        nb = lo;
        nv = vlo;
        np = g + 32 * NW;
    }

    __device__ __forceinline__ void shift(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ invalid) {
        const int s = (int)(np & 31);
        const uint64_t code = (nb >> (2 * s)) & 3ull;
        const uint32_t inv = (nv >> s) & 1u;
#pragma unroll
        for (int w = 0; w < NW - 1; w++) {
            b[w] = (b[w] >> 2) | (b[w + 1] << 62);
            v[w] = (v[w] >> 1) | (v[w + 1] << 31);
        }
        b[NW - 1] = (b[NW - 1] >> 2) | (code << 62);
        v[NW - 1] = (v[NW - 1] >> 1) | (inv << 31);
        np++;
        if ((np & 31) == 0) {
            nb = codes[np >> 5];
            nv = invalid[np >> 5];
        }
    }

    __device__ __forceinline__ int code_at(int j) const {
        uint64_t w = b[0];
#pragma unroll
        for (int i = 1; i < NW; i++) w = ((j >> 5) == i) ? b[i] : w;
        return (int)((w >> (2 * (j & 31))) & 3ull);
    }
    __device__ __forceinline__ uint32_t invalid_at(int j) const {
        uint32_t w = v[0];
#pragma unroll
        for (int i = 1; i < NW; i++) w = ((j >> 5) == i) ? v[i] : w;
        return (w >> (j & 31)) & 1u;
    }
    // any invalid base among the first k positions
    __device__ __forceinline__ bool any_invalid(int k) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const int rem = k - 32 * i; // bases of the k-mer that live in word i
            const uint32_t m = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            acc |= v[i] & m;
        }
        return acc != 0;
    }
};

template <int NW, bool STREAMING, bool WIDE, bool COUNT>
__global__ void __launch_bounds__(256) walk_kernel(const WalkParams P) {
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_items = *P.n_items;
    // contiguous item range of this warp
    int64_t next = (int64_t)(((__int128)n_items * gw) / nw);
    const int64_t end = (int64_t)(((__int128)n_items * (gw + 1)) / nw);

    const int k = ix.k, p = ix.p;
    const uint64_t pmask = p ? ((1ull << (2 * p)) - 1ull) : 0ull;
    const Sector* const pre_base = reinterpret_cast<const Sector*>(ix.precalc);

    Window<NW> win;
    int64_t l = 0, r = 0;     // current interval; in STREAM mode l is the previous column
    int64_t outp = 0;         // next result slot
    int remaining = 0;        // k-mers left in the item, current one included
    int j = 0;                // characters of the current k-mer already consumed
    int mode = M_NEED;
    unsigned long long st_lookups = 0, st_hits = 0, st_ranks = 0, st_sectors = 0;

    while (true) {
        // ---- refill finished lanes from the warp's range
        const unsigned need = __ballot_sync(FULL, mode == M_NEED);
        if (need) {
            const int64_t mine = next + __popc(need & lt_mask);
            next += __popc(need);
            if (mode == M_NEED) {
                if (mine < end) {
                    const int64_t g = P.item_base[mine];
                    outp = P.item_out[mine];
                    remaining = P.item_cnt[mine];
                    win.init(P.codes, P.invalid, g);
                    mode = M_START;
                } else {
                    mode = M_DONE;
                }
            }
        }
        if (__all_sync(FULL, mode == M_DONE)) break;

        // ---- classify: every live lane becomes one of  PRE (table row), STEP (one interval step)
        //      or INVALID (a k-mer covering a non-ACGT byte: answer -1 without touching memory)
        const bool from_stream = STREAMING && mode == M_STREAM;
        bool invalid = false, pre = false, step = false;
        if (mode == M_START) {
            invalid = win.any_invalid(k);             // SBWT.hh:399,428
            if (!invalid) {
                if (p > 0) pre = true;
                else { l = 0; r = ix.n_nodes - 1; j = 0; step = true; }
            }
        } else if (mode == M_WALK) {
            step = true;
        } else if (from_stream) {
            invalid = win.invalid_at(k - 1) != 0;     // SBWT.hh:568
            if (!invalid) { r = l; j = k - 1; step = true; } // one step on [col, col] with the new character
        }

        // ---- addresses + loads: the same instructions for every kind of lane
        const int c = win.code_at(j);
        const BlockPos b0 = split_pos<WIDE>(l), b1 = split_pos<WIDE>(r + 1);
        const uint64_t pidx = win.b[0] & pmask; // first character = least significant digit (SBWT.hh:396-401)
        const Sector* a0 = pre ? pre_base + (pidx >> 1) : sector_addr<WIDE>(ix, b0.blk, c);
        const bool two = step && (b1.blk != b0.blk);
        Sector s0, s1;
        if (pre || step) s0 = ld_sector(a0);
        if (two) s1 = ld_sector(sector_addr<WIDE>(ix, b1.blk, c));

        // ---- consume
        int64_t nl, nr;
        {
            uint32_t vl = sector_rank(s0, b0.off);
            uint32_t vr = sector_rank(two ? s1 : s0, b1.off);
            nl = (int64_t)vl;
            nr = (int64_t)vr;
            if (WIDE && step) {
                nl += __ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b0.blk >> ix.sb_shift));
                nr += __ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b1.blk >> ix.sb_shift));
            }
            nr -= 1;
        }
        if (pre) {
            const bool hi = (pidx & 1) != 0;
            const uint32_t e0 = hi ? s0.w[4] : s0.w[0], e1 = hi ? s0.w[5] : s0.w[1];
            const uint32_t e2 = hi ? s0.w[6] : s0.w[2], e3 = hi ? s0.w[7] : s0.w[3];
            nl = (int64_t)(((uint64_t)e1 << 32) | e0);
            nr = (int64_t)(((uint64_t)e3 << 32) | e2);
        }
        bool miss = pre ? (nl < 0) : (nl > nr); // absent p-mer (SBWT.hh:424) / empty interval (SBWT.hh:433)
        if (COUNT) {
            if (step) { st_ranks += 2; st_sectors += two ? 2 : 1; }
            if (pre) st_sectors++;
        }
        if (STREAMING && from_stream && step && (miss || !ix.edges_at_starts)) {
            // literal form (SBWT.hh:562-563): the step must start from the suffix-group start of col.
            // With edges only at group starts a set bit proves col is the start, so only a clear bit
            // (or an index that violates the invariant) gets here.
            int64_t s = l;
            while (true) {
                const uint32_t w = __ldg(ix.sgs + (s >> 5)) & (0xFFFFFFFFu >> (31 - (int)(s & 31)));
                if (w) { s = (s & ~31ll) + (31 - __clz(w)); break; }
                s = (s & ~31ll) - 1;
            }
            if (COUNT) st_sectors++;
            if (s != l) {
                const BlockPos bs = split_pos<WIDE>(s);
                const Sector ss = ld_sector(sector_addr<WIDE>(ix, bs.blk, c));
                if (COUNT) st_sectors += bs.blk != b0.blk;
                miss = sector_bit(ss, bs.off) == 0;
                nl = lf_value<WIDE>(ix, ss, bs.blk, bs.off, c);
                nr = nl;
            }
        }
        const int nj = pre ? p : j + 1;
        const bool done = (pre || step) && !miss && nj == k; // a k-mer interval is a singleton (SBWT.hh:410-413)
        if (pre || step) {
            l = nl;
            r = nr;
            j = nj;
            mode = M_WALK;
        }

        // ---- emit: one result, advance to the next k-mer of the item
        if (invalid || ((pre || step) && (miss || done))) {
            const int64_t ans = done ? nl : -1;
            __stcs(P.out + outp, ans);
            outp++;
            if (COUNT) { st_lookups++; st_hits += ans >= 0; }
            mode = (STREAMING && done) ? M_STREAM : M_START;
            if (--remaining == 0) mode = M_NEED;
            else win.shift(P.codes, P.invalid);
        }
    }

    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_lookups += __shfl_xor_sync(FULL, st_lookups, o);
            st_hits += __shfl_xor_sync(FULL, st_hits, o);
            st_ranks += __shfl_xor_sync(FULL, st_ranks, o);
            st_sectors += __shfl_xor_sync(FULL, st_sectors, o);
        }
        if (lane == 0) {
            atomicAdd(P.stats + 0, st_lookups);
            atomicAdd(P.stats + 1, st_hits);
            atomicAdd(P.stats + 2, st_ranks);
            atomicAdd(P.stats + 3, st_sectors);
        }
    }
}

} // namespace sbwt_b200
