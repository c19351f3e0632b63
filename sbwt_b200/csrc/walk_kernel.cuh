// walk_kernel.cuh -- the colex-interval walk on the device (sm_100a).
//
// Replaces SBWT::search (include/sbwt/SBWT.hh:390-415), SBWT::update_sbwt_interval
// (SBWT.hh:423-437) and SBWT::streaming_search (SBWT.hh:545-581) together with the per-read
// loops of src/CLI/sbwt_search.cpp:45-91.
//
// Work item = up to `window` consecutive k-mers of one read (a whole read when it is short).
// One LANE owns one item at a time and is a small state machine; every trip of the warp loop
// each lane performs at most one memory round trip, and every kind of round trip runs through
// the SAME instructions (the first profile of this kernel was integer-ALU bound with 14 of 32
// lanes active per instruction, profiles/r01_walk_v1_summary.txt):
//
//   STEP   one interval step  [l,r] -> [C[c]+rank_c(l), C[c]+rank_c(r+1)-1]  (two sectors, one
//          when l and r+1 fall into the same 224-column block). A streaming step (previous k-mer
//          found at column col, SBWT.hh:561-575) is the same step on [col, col] with the new
//          character: in a reference-built index only suffix-group starts carry edges
//          (SURVEY.md section 8(a) note 7), so a set bit_c(col) already proves col is the group
//          start; a clear bit (or an index violating the invariant) takes the literal
//          walk-back over suffix_group_starts on a slow path.
//
// Everything that happens once per k-mer rather than once per step -- storing the result,
// sliding the read window, validity checks, and the lookup of the first p characters in the
// search table (kmer_prefix_precalc, SBWT.hh:404; a second, dependent load in the same trip,
// hidden by the other resident warps because the kernel is issue-bound, not latency-bound) --
// lives in one divergent ADVANCE block, so lanes in long from-scratch walks, lanes streaming along a matching
// read and lanes that just fetched a read share the step code at full width. The pointer chase
// is hidden by the ~1000-2000 resident lanes per SM, each with one or two sector loads in
// flight. Finished lanes refill from the warp's own contiguous item range (ballot/popc).
#pragma once

#include <type_traits>

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
    int index_evict_last;      // L2 policy of the index / table loads
};

enum : int { M_NEED = 0, M_STEP = 1, M_DONE = 2 };

// Sliding window over the packed read: base t of the window (t = 0 is the first character of
// the current k-mer) sits at bits [2(t%32), 2(t%32)+2) of b[t/32]; v holds the invalid flags.
// nb / nv hold the not yet consumed bases of the packed word the window will slide into.
template <int NW>
struct Window {
    uint64_t b[NW];
    uint32_t v[NW];
    uint64_t nb; // next bases, already shifted so that the next base sits at bits [0,2)
    uint32_t nv; // likewise for the invalid flags
    uint32_t left; // bases left in nb / nv
    uint32_t wi;   // index of the word nb came from

    __device__ __forceinline__ void init(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ invalid, int64_t g) {
        const uint32_t w0 = (uint32_t)(g >> 5);
        const int sh = (int)(g & 31);
        uint64_t lo = codes[w0];
        uint32_t vlo = invalid[w0];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint64_t hi = codes[w0 + w + 1];
            const uint32_t vhi = invalid[w0 + w + 1];
            b[w] = sh ? ((lo >> (2 * sh)) | (hi << (64 - 2 * sh))) : lo;
            v[w] = __funnelshift_r(vlo, vhi, sh);
            lo = hi;
            vlo = vhi;
        }
        wi = w0 + NW;
        nb = lo >> (2 * sh);
        nv = vlo >> sh;
        left = 32 - sh;
    }

    __device__ __forceinline__ void shift(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ invalid) {
        const uint64_t code = nb & 3ull;
        const uint32_t inv = nv & 1u;
#pragma unroll
        for (int w = 0; w < NW - 1; w++) {
            b[w] = (b[w] >> 2) | (b[w + 1] << 62);
            v[w] = (v[w] >> 1) | (v[w + 1] << 31);
        }
        b[NW - 1] = (b[NW - 1] >> 2) | (code << 62);
        v[NW - 1] = (v[NW - 1] >> 1) | (inv << 31);
        nb >>= 2;
        nv >>= 1;
        if (--left == 0) {
            wi++;
            nb = codes[wi];
            nv = invalid[wi];
            left = 32;
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
    // any invalid base among the first k positions; kmask[i] = mask of the k-mer's bases in word i
    __device__ __forceinline__ bool any_invalid(const uint32_t* kmask) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) acc |= v[i] & kmask[i];
        return acc != 0;
    }
};

template <int NW, bool STREAMING, bool WIDE, bool COUNT>
__global__ void __launch_bounds__(256, (NW == 1 && !WIDE) ? 5 : 4) walk_kernel(const WalkParams P) {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t; // columns fit 32 bits in narrow mode
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

    const int k = ix.k, p = ix.tp; // p: characters answered by the search table
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const Sector* const pre_base = reinterpret_cast<const Sector*>(ix.table);
    const Sector* const sec_base = ix.sectors;
    const uint64_t pol = make_l2_policy(P.index_evict_last != 0);
    const pos_t last_col = (pos_t)(ix.n_nodes - 1);
    uint32_t kmask[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const int rem = k - 32 * i;
        kmask[i] = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    }

    Window<NW> win;
    pos_t l = 0, r = 0;       // current interval
    int64_t* outq = P.out;    // next result slot
    int remaining = 0;        // k-mers left in the item, current one included
    int j = 0;                // characters of the current k-mer already consumed
    int mode = M_NEED;
    bool fs = false;          // this STEP is a streaming step (previous k-mer was found at column l)
    unsigned long long st_lookups = 0, st_hits = 0, st_ranks = 0, st_sectors = 0;

    // Runs when a lane moves on to a new k-mer (window already positioned on it). Answers on the
    // spot the k-mers that need no walk -- those covering a non-ACGT byte (SBWT.hh:399,428,568),
    // those whose first p characters are absent from the table (SBWT.hh:424) and, when p == k,
    // those the table answers completely -- and leaves the lane in STEP mode for the first k-mer
    // that needs the index, or in NEED mode when the item is exhausted.
    auto setup = [&](bool stream) {
        while (true) {
            if (remaining == 0) { mode = M_NEED; return; }
            int64_t ans = -1;
            if (STREAMING && stream) {
                if (!win.invalid_at(k - 1)) { r = l; j = k - 1; fs = true; mode = M_STEP; return; }
            } else if (!win.any_invalid(kmask)) {
                fs = false;
                if (p == 0) { l = 0; r = last_col; j = 0; mode = M_STEP; return; }
                const uint32_t pidx = (uint32_t)win.b[0] & pmask; // first character = least significant digit (SBWT.hh:396-401)
                const Sector t = ld_sector(pre_base + (pidx >> (WIDE ? 1 : 2)), pol);
                if (COUNT) st_sectors++;
                bool absent;
                if (WIDE) {
                    const bool hi = pidx & 1u;
                    const uint32_t e0 = hi ? t.w[4] : t.w[0], e1 = hi ? t.w[5] : t.w[1];
                    const uint32_t e2 = hi ? t.w[6] : t.w[2], e3 = hi ? t.w[7] : t.w[3];
                    l = (pos_t)(((uint64_t)e1 << 32) | e0);
                    r = (pos_t)(((uint64_t)e3 << 32) | e2);
                    absent = (int32_t)e1 < 0;
                } else {
                    const bool q1 = pidx & 1u, q2 = pidx & 2u;
                    const uint32_t a_l = q1 ? t.w[2] : t.w[0], a_r = q1 ? t.w[3] : t.w[1];
                    const uint32_t b_l = q1 ? t.w[6] : t.w[4], b_r = q1 ? t.w[7] : t.w[5];
                    l = (pos_t)(q2 ? b_l : a_l);
                    r = (pos_t)(q2 ? b_r : a_r);
                    absent = l == (pos_t)0xFFFFFFFFu;
                }
                if (!absent) {
                    if (p < k) { j = p; mode = M_STEP; return; }
                    ans = (int64_t)l; // p == k: the row is the answer (a singleton, SBWT.hh:410-413)
                }
            }
            __stcs(outq, ans);
            outq++;
            if (COUNT) { st_lookups++; st_hits += ans >= 0; }
            stream = STREAMING && ans >= 0;
            if (--remaining) win.shift(P.codes, P.invalid);
        }
    };

    while (true) {
        // ---- refill finished lanes from the warp's range
        const unsigned need = __ballot_sync(FULL, mode == M_NEED);
        if (need) {
            const int64_t mine = next + __popc(need & lt_mask);
            next += __popc(need);
            if (mode == M_NEED) {
                if (mine < end) {
                    const int64_t g = P.item_base[mine];
                    outq = P.out + P.item_out[mine];
                    remaining = P.item_cnt[mine];
                    win.init(P.codes, P.invalid, g);
                    setup(false);
                } else {
                    mode = M_DONE;
                }
            }
            if (__all_sync(FULL, mode == M_DONE)) break;
        }

        // ---- STEP: the same instructions for every live lane
        const bool step = mode == M_STEP;
        const int c = win.code_at(j);
        const BlockPos b0 = split_pos<WIDE>((int64_t)l), b1 = split_pos<WIDE>((int64_t)r + 1);
        const bool two = step && (b1.blk != b0.blk);
        Sector s0, s1;
        if (step) s0 = ld_sector(sec_base + ((b0.blk << 2) + c), pol);
        if (two) s1 = ld_sector(sec_base + ((b1.blk << 2) + c), pol);
        else s1 = s0;

        const SectorPrefix pf0 = sector_prefix(s0);
        const SectorPrefix pf1 = two ? sector_prefix(s1) : pf0;
        pos_t nl = (pos_t)sector_rank_fast(s0, pf0, b0.off);
        pos_t nr = (pos_t)sector_rank_fast(s1, pf1, b1.off);
        if (WIDE && step) {
            nl += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b0.blk >> ix.sb_shift));
            nr += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b1.blk >> ix.sb_shift));
        }
        nr -= 1;
        bool miss = nl > nr; // empty interval (SBWT.hh:433)
        if (COUNT && step) { st_ranks += 2; st_sectors += two ? 2 : 1; }
        if (STREAMING && fs && step && (miss || !ix.edges_at_starts)) {
            // literal form (SBWT.hh:562-563): the step has to start from the suffix-group start of
            // column l. With edges only at group starts a set bit proves l is the start, so only a
            // clear bit (or an index that violates the invariant) gets here.
            int64_t s = (int64_t)l;
            while (true) {
                const uint32_t w = __ldg(ix.sgs + (s >> 5)) & (0xFFFFFFFFu >> (31 - (int)(s & 31)));
                if (w) { s = (s & ~31ll) + (31 - __clz(w)); break; }
                s = (s & ~31ll) - 1;
            }
            if (COUNT) st_sectors++;
            if (s != (int64_t)l) {
                const BlockPos bs = split_pos<WIDE>(s);
                const Sector ss = ld_sector(sec_base + ((bs.blk << 2) + c), pol);
                if (COUNT) st_sectors += bs.blk != b0.blk;
                miss = sector_bit(ss, bs.off) == 0;
                nl = (pos_t)lf_value<WIDE>(ix, ss, bs.blk, bs.off, c);
                nr = nl;
            }
        }
        const bool done = step && !miss && (j + 1 == k); // a k-mer interval is a singleton (SBWT.hh:410-413)
        if (step) {
            l = nl;
            r = nr;
            j++;
            fs = false;
        }

        // ---- ADVANCE: one result, slide to the next k-mer of the item
        if (step && (miss || done)) {
            __stcs(outq, done ? (int64_t)nl : (int64_t)-1);
            outq++;
            if (COUNT) { st_lookups++; st_hits += done; }
            if (--remaining) win.shift(P.codes, P.invalid);
            setup(done);
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
