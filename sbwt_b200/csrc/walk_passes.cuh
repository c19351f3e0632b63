// walk_passes.cuh -- the streaming walk as a few specialised passes over global work lists (sm_100a).
//
// walk_kernel (walk_kernel.cuh) runs NARROW, CHAIN and PROBE inside one persistent kernel: every warp sorts its own work
// through shared-memory queues. That kernel is compiled for the phase that needs most registers and shared memory, so all
// phases run at its occupancy (24 warps per SM), and each phase is a chain of dependent memory latencies a warp sits
// through alone: the round-2 captures show a third of the c2 walk in NARROW / PROBE visits of a few microseconds each, and
// the CHAIN loop waiting with six warps per scheduler. Here the same three phases are three small
// kernels, each at the occupancy its own state allows, handing work over through global lists:
//
//   first_pass_kernel   one lane per work item (or fresh item): table jump + interval steps on the item's first k-mer
//                  (SBWT::search, SBWT.hh:390-415). Survivors -> S list; a dead first k-mer writes -1 and its item's
//                  valid remainder -> P list; whatever covers an invalid base -> L list (answered per k-mer later).
//   chain_pass_kernel   one lane per survivor: the rest of its own k-mer, then the following k-mers of its item, one sector
//                  and one result per step (SBWT::streaming_search, SBWT.hh:561-575), results written by rounds through
//                  a shared-memory stage. A miss ends the chain; the rest of the item -> P list.
//   probe_pass_kernel   ranges of presumed misses are probed at every D-th k-mer (see walk_kernel.cuh: a walk that dies after
//                  j characters proves every k-mer covering those characters absent); proven -1s are written with
//                  coalesced stores, the first segment that is not proven absent restarts the read there -> F list.
//
// The host runs   first(items) -> chain -> probe -> first(F) -> chain -> probe -> ...   for a fixed number of rounds
// (a read needs one round per stretch of found k-mers; counts never leave the device, so an empty round is three empty
// launches), and whatever is left -- the L list, and the F list of the last round -- goes through walk_kernel, which
// answers anything. Results are the reference's, as in walk_kernel: the same Stepper arithmetic, the same control flow.
#pragma once

#include "walk_kernel.cuh"

namespace sbwt_b200 {

constexpr int kPassThreads = 256;

// a survivor of first_pass_kernel: its first k-mer has consumed `meta & 0xFF` characters and stands on column col
template <bool WIDE>
struct SurvRec;
template <>
struct __align__(16) SurvRec<false> {
    uint32_t base, out, meta, col32; // meta = jl | (k-mers of the item that may be streamed << 8)
    __device__ __forceinline__ uint32_t col() const { return col32; }
    __device__ __forceinline__ void set_col(uint32_t c) { col32 = c; }
};
template <>
struct __align__(16) SurvRec<true> {
    uint32_t base, out, meta, pad;
    int64_t col64, pad2;
    __device__ __forceinline__ int64_t col() const { return col64; }
    __device__ __forceinline__ void set_col(int64_t c) { col64 = c; pad = 0; pad2 = 0; }
};

struct PassLists {
    void* surv;                 // S: SurvRec<WIDE>[cap]
    uint4* ranges;              // P: {base, out, cnt, -}
    WalkItem* fresh;            // F: items created by probe_kernel (restarts inside a read); two buffers take turns
    WalkItem* left;             // L: what walk_kernel answers at the end
    unsigned long long* n;      // [0] = |S|, [1] = |P|, [3] = |L|, [4] = cursor of the running kernel, [2] / [5] = |F| of the two buffers
    unsigned long long* n_fresh; // the counter of `fresh`
    uint32_t cap, cap_left;
};

// one atomicAdd per warp: lane `want`s a slot; returns its index (or ~0u)
__device__ __forceinline__ uint32_t warp_append(unsigned long long* counter, bool want) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return want ? (uint32_t)base + __popc(m & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}

struct PassStats {
    unsigned long long lookups = 0, hits = 0, ranks = 0, sectors = 0;
    __device__ __forceinline__ void flush(unsigned long long* stats) {
        const unsigned FULL = 0xFFFFFFFFu;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lookups += __shfl_xor_sync(FULL, lookups, s);
            hits += __shfl_xor_sync(FULL, hits, s);
            ranks += __shfl_xor_sync(FULL, ranks, s);
            sectors += __shfl_xor_sync(FULL, sectors, s);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(stats + 0, lookups);
            atomicAdd(stats + 1, hits);
            atomicAdd(stats + 2, ranks);
            atomicAdd(stats + 3, sectors);
        }
    }
};

template <bool WIDE, int LAY>
__device__ __forceinline__ Stepper<WIDE, LAY> make_stepper(const DeviceIndexView& ix) {
    Stepper<WIDE, LAY> ST;
    ST.sec = ix.sectors; ST.cmp = ix.compact; ST.cbase = ix.cbase; ST.sbbase = ix.sbbase; ST.n_sb = ix.n_sb; ST.sb_shift = ix.sb_shift;
    ST.pol = make_l2_policy(true);
    ST.pol_cold = make_l2_policy(false);
    return ST;
}

// ------------------------------------------------------------------------------------------------ first k-mers
// items: n_items_ptr[0] work items. use_nvalid = false for fresh items (all of their k-mers are free of invalid bases).
template <bool WIDE, bool COUNT, bool OUT32, int KW, int LAY>
__global__ void __launch_bounds__(kPassThreads) first_pass_kernel(const WalkParams P, const WalkItem* __restrict__ items,
                                                              const unsigned long long* __restrict__ n_items_ptr, PassLists LS, int last_round) {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    const DeviceIndexView& ix = P.ix;
    const uint32_t n_items = (uint32_t)*n_items_ptr;
    const uint32_t k = (uint32_t)ix.k, p = (uint32_t)ix.tp;
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const Stepper<WIDE, LAY> ST = make_stepper<WIDE, LAY>(ix);
    uint64_t pol_t = ST.pol;
    if (P.table_streams) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_t));
    PassStats st;
    SurvRec<WIDE>* const surv = reinterpret_cast<SurvRec<WIDE>*>(LS.surv);
    const uint32_t n_round = (n_items + 31u) & ~31u; // whole warps take part in the appends
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool have = i < n_items;
        uint32_t b = 0, o = 0, cnt = 0, nvalid = 0;
        if (have) {
            const uint4 it = __ldg(reinterpret_cast<const uint4*>(items) + i);
            b = it.x; o = it.y; cnt = it.z; nvalid = it.w;
        }
        // what covers an invalid base is answered one k-mer at a time by walk_kernel (as its TODO ranges are)
        const bool rest = have && nvalid < cnt;
        const uint32_t at_l = warp_append(LS.n + 3, rest);
        if (rest && at_l < LS.cap_left) LS.left[at_l] = WalkItem{b + nvalid, o + nvalid, cnt - nvalid, 0u};
        bool alive = have && nvalid > 0;
        const bool act = alive;
        pos_t l = 0, r = (pos_t)(ix.n_nodes - 1);
        uint32_t jl = p;
        if (act) {
            const KmerWin<KW> win = load_win<KW>(P.codes, b);
            if (p != 0) { // first character = least significant digit of the table index (SBWT.hh:396-401)
                const TableRow<WIDE> row = TableRow<WIDE>::load(ix.table, win.w[0] & pmask, pol_t);
                l = (pos_t)row.l;
                r = (pos_t)row.r;
                if (row.absent()) alive = false;
                if (COUNT) st.sectors++;
            }
            uint32_t single = (alive && l == r) ? 1u : 0u;
            while (alive && jl < k && single <= kSingleHold) {
                const int c = (int)win_char<KW>(win, jl);
                pos_t nl = 0, nr = 0;
                const uint32_t ns = ST.narrow(l, r, c, nl, nr);
                if (COUNT) { st.ranks += 2; st.sectors += ns; }
                if (nl > nr) alive = false; // empty interval (SBWT.hh:433)
                else {
                    l = nl;
                    r = nr;
                    jl++;
                    single = (nl == nr) ? single + 1 : 0u;
                }
            }
        }
        const bool dead = act && !alive;
        if (dead) {
            store_result<OUT32>(P, o, -1);
            if (COUNT) st.lookups++;
        }
        // the valid k-mers after an absent first one are probed
        const bool pp = dead && nvalid > 1;
        const uint32_t at_p = warp_append(LS.n + 1, pp);
        if (pp && at_p < LS.cap) LS.ranges[at_p] = make_uint4(b + 1u, o + 1u, nvalid - 1u, 0u);
        const uint32_t at_s = warp_append(LS.n + 0, alive);
        if (alive && at_s < LS.cap) {
            SurvRec<WIDE> rec;
            rec.base = b; rec.out = o; rec.meta = jl | (nvalid << 8);
            rec.set_col(l);
            surv[at_s] = rec;
        }
        (void)last_round;
    }
    if (COUNT) st.flush(P.stats);
}

// ------------------------------------------------------------------------------------------------ chains
template <bool WIDE>
struct ChainShared {
    ChainStage<WIDE, 1> stage[kPassThreads / 32];
};

template <bool WIDE, bool COUNT, bool OUT32, int LAY>
__global__ void __launch_bounds__(kPassThreads, WIDE ? 3 : 5) chain_pass_kernel(const WalkParams P, PassLists LS) {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    typedef ChainStage<WIDE, 1> CS;
    constexpr uint32_t R = CS::kRound;
    extern __shared__ __align__(16) unsigned char chain_smem[];
    CS& STG = reinterpret_cast<ChainShared<WIDE>*>(chain_smem)->stage[threadIdx.x >> 5];
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const uint32_t k = (uint32_t)ix.k;
    const uint32_t n_surv = (uint32_t)min(LS.n[0], (unsigned long long)LS.cap);
    const uint32_t ophR = OUT32 ? (uint32_t)(((uintptr_t)P.out32 >> 2) & (R - 1u)) : (uint32_t)(((uintptr_t)P.out >> 3) & (R - 1u));
    const Stepper<WIDE, LAY> ST = make_stepper<WIDE, LAY>(ix);
    const SurvRec<WIDE>* const surv = reinterpret_cast<const SurvRec<WIDE>*>(LS.surv);
    PassStats st;

    auto resolve = [&](pos_t colv, int c, bool streaming_step, pos_t& ncol, bool& hit) {
        int64_t cblk;
        ST.classic_step(colv, c, ncol, hit, cblk);
        if (!hit && streaming_step) { // literal walk-back (SBWT.hh:562-563): the step starts from the suffix-group start of col
            int64_t g = (int64_t)colv;
            while (true) {
                const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
                if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
                g = (g & ~31ll) - 1;
            }
            if (COUNT) st.sectors++;
            if (g != (int64_t)colv) {
                int64_t gblk;
                ST.classic_step((pos_t)g, c, ncol, hit, gblk);
                if (COUNT) st.sectors += gblk != cblk;
            }
        }
    };

    while (true) {
        uint32_t g0 = 0;
        if (lane == 0) g0 = (uint32_t)atomicAdd(LS.n + 4, 32ull);
        g0 = __shfl_sync(FULL, g0, 0);
        if (g0 >= n_surv) break;
        bool act = g0 + (uint32_t)lane < n_surv;
        pos_t col = 0;
        uint32_t pos = 0, cw = 0, nx = 0, o = 0, oend = 0, quiet = 0;
        bool fs = false;
        if (act) {
            const SurvRec<WIDE> rec = surv[g0 + lane];
            o = rec.out;
            col = rec.col();
            const uint32_t j = rec.meta & 0xFFu;
            oend = o + (rec.meta >> 8);
            pos = rec.base + j;
            cw = __ldg(P.codes + (pos >> 4));
            nx = __ldg(P.codes + (pos >> 4) + 1);
            if (j == k) { // first_pass_kernel already completed the first k-mer (SBWT.hh:410-413)
                store_result<OUT32>(P, o, (int64_t)col);
                if (COUNT) { st.lookups++; st.hits++; }
                o++;
                fs = true;
                if (o == oend) act = false;
            } else {
                quiet = k - j - 1u;
            }
        }
        auto advance = [&]() {
            pos++;
            const uint32_t ph = pos & 15u;
            if (ph == 0) cw = nx;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
                         : "+r"(nx) : "l"(P.codes + (pos >> 4) + 1), "r"(ph));
        };
        auto push_rest = [&](bool want, uint32_t kstart, uint32_t ov) { // the rest of the item after a miss (SBWT.hh:557-559)
            const uint32_t at = warp_append(LS.n + 1, want);
            if (want && at < LS.cap) LS.ranges[at] = make_uint4(kstart + 1u, ov, oend - ov, 0u);
        };

        while (true) {
            // ---- own k-mers: steps without a result
            while (true) {
                const bool run = act && quiet > 0;
                if (!__any_sync(FULL, run)) break;
                bool want = false;
                uint32_t kstart = 0;
                if (run) {
                    const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                    typename Stepper<WIDE, LAY>::Load L;
                    ST.issue(L, col, c);
                    pos_t ncol = 0;
                    bool hit = false;
                    if (!ST.eval(L, col, c, ncol, hit)) resolve(col, c, false, ncol, hit);
                    if (COUNT) { st.ranks += 2; st.sectors += 1; }
                    if (hit) {
                        col = ncol;
                        advance();
                        quiet--;
                    } else { // the item's first k-mer is absent
                        kstart = pos - (k - 1u - quiet);
                        store_result<OUT32>(P, o, -1);
                        if (COUNT) st.lookups++;
                        o++;
                        act = false;
                        want = o < oend;
                    }
                }
                if (__any_sync(FULL, want)) push_rest(want, kstart, o);
            }
            // ---- one round
            uint32_t lo = 0, width = 0;
            if (act) {
                lo = (o + ophR) & (R - 1u);
                width = min(R - lo, oend - o);
            }
            const uint32_t s_first = __reduce_min_sync(FULL, act ? lo : R);
            const uint32_t s_last = __reduce_max_sync(FULL, lo + width);
            if (s_last == 0) break;
            bool ended = false;
            for (uint32_t sl = s_first; sl < s_last; sl++) {
                const bool run = (sl - lo) < width;
                const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                typename Stepper<WIDE, LAY>::Load L;
                if (run) ST.issue(L, col, c);
                pos_t ncol = 0;
                bool ok = true;
                if (run) {
                    bool hit = false;
                    const bool fast = ST.eval(L, col, c, ncol, hit);
                    ok = fast && hit;
                }
                if (__all_sync(FULL, ok)) {
                    if (run) {
                        STG.v[0][lane][sl] = ncol;
                        col = ncol;
                        advance();
                    }
                } else {
                    bool want = false;
                    uint32_t kstart = 0;
                    if (run) {
                        bool hit = ok;
                        if (!hit) resolve(col, c, fs || sl > lo, ncol, hit);
                        if (hit) {
                            STG.v[0][lane][sl] = ncol;
                            col = ncol;
                            advance();
                        } else { // absent: the chain ends with a -1
                            STG.v[0][lane][sl] = (pos_t)-1;
                            kstart = pos - (k - 1u);
                            width = sl + 1u - lo;
                            ended = true;
                            want = o + width < oend;
                            if (COUNT) st.hits--;
                        }
                    }
                    if (__any_sync(FULL, want)) push_rest(want, kstart, o + width);
                }
            }
            STG.obase[0][lane] = o - lo;
            STG.mask[0][lane] = width ? (((1u << width) - 1u) << lo) : 0u;
            __syncwarp();
            CS::template flush<OUT32>(STG, 0, lane, P.out, P.out32);
            __syncwarp();
            if (act) {
                if (COUNT) { st.ranks += 2ull * width; st.sectors += width; st.lookups += width; st.hits += width; }
                if (width) fs = true;
                o += width;
                if (ended || o == oend) act = false;
            }
        }
    }
    if (COUNT) st.flush(P.stats);
}

// ------------------------------------------------------------------------------------------------ probes
// A warp takes four ranges at a time, eight lanes (= eight probes per pass) each, and walks each range to its end or to
// the first segment that is not proven absent.
template <bool WIDE, bool COUNT, bool OUT32, int KW, int LAY>
__global__ void __launch_bounds__(kPassThreads) probe_pass_kernel(const WalkParams P, PassLists LS, int last_round) {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const uint32_t k = (uint32_t)ix.k, p = (uint32_t)ix.tp, D = P.probe_stride;
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const uint32_t n_ranges = (uint32_t)min(LS.n[1], (unsigned long long)LS.cap);
    const Stepper<WIDE, LAY> ST = make_stepper<WIDE, LAY>(ix);
    uint64_t pol_t = ST.pol;
    if (P.table_streams) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_t));
    PassStats st;
    while (true) {
        uint32_t g0 = 0;
        if (lane == 0) g0 = (uint32_t)atomicAdd(LS.n + 4, 4ull);
        g0 = __shfl_sync(FULL, g0, 0);
        if (g0 >= n_ranges) break;
        // the group's range: [tb, tb + tc) k-mers, results from `to`
        uint32_t tb = 0, to = 0, tc = 0;
        if (g0 + (uint32_t)grp < n_ranges) {
            const uint4 rg = __ldg(LS.ranges + g0 + grp);
            tb = rg.x; to = rg.y; tc = rg.z;
        }
        while (__any_sync(FULL, tc > 0)) {
            // probe `sub` of this pass: the last k-mer of segment [sub D, sub D + D) of the range
            const uint32_t seg_lo = (uint32_t)sub * D;
            const bool act = seg_lo < tc;
            const uint32_t xm = min(seg_lo + D, tc) - 1u;
            bool alive = act;
            uint32_t jl = p;
            if (act) {
                const KmerWin<KW> win = load_win<KW>(P.codes, tb + xm);
                pos_t l = 0, r = (pos_t)(ix.n_nodes - 1);
                if (p != 0) {
                    const TableRow<WIDE> row = TableRow<WIDE>::load(ix.table, win.w[0] & pmask, pol_t);
                    l = (pos_t)row.l;
                    r = (pos_t)row.r;
                    if (row.absent()) { alive = false; jl = p - 1u; }
                    if (COUNT) st.sectors++;
                }
                // (as in walk_kernel's NARROW: until the interval is empty, the k-mer complete, or a singleton that keeps living
                // -- the last two are "not proven absent": the segment restarts the read)
                uint32_t single = (alive && l == r) ? 1u : 0u;
                while (alive && jl < k && single <= kSingleHold) {
                    const int c = (int)win_char<KW>(win, jl);
                    pos_t nl = 0, nr = 0;
                    const uint32_t ns = ST.narrow(l, r, c, nl, nr);
                    if (COUNT) { st.ranks += 2; st.sectors += ns; }
                    if (nl > nr) alive = false;
                    else {
                        l = nl;
                        r = nr;
                        jl++;
                        single = (nl == nr) ? single + 1 : 0u;
                    }
                }
            }
            // a dead probe proves the k-mers [xm + jl + 1 - k, xm] of its range absent; "covered": its whole segment
            const bool covered = act && !alive && (int)(xm + jl + 1u) - (int)k <= (int)seg_lo;
            const unsigned badmask = __ballot_sync(FULL, act && !covered);
            const unsigned gb = (badmask >> (grp * 8)) & 0xFFu;
            uint32_t nfill;
            bool fresh = false;
            if (gb) { // restart the read at the first segment that is not proven absent
                nfill = (uint32_t)(__ffs(gb) - 1) * D;
                fresh = true;
            } else {
                nfill = min(tc, 8u * D);
            }
            // the proven misses, coalesced per group
            for (uint32_t x = (uint32_t)sub; x < nfill; x += 8u) {
                store_result<OUT32>(P, to + x, -1);
                if (COUNT) st.lookups++;
            }
            const uint32_t nb = tb + nfill, no = to + nfill, nc = tc - min(nfill, tc);
            const bool want = fresh && sub == 0 && nc > 0; // one lane of the group appends
            if (last_round) { // no further round: walk_kernel takes it from here
                const uint32_t at = warp_append(LS.n + 3, want);
                if (want && at < LS.cap_left) LS.left[at] = WalkItem{nb, no, nc, nc};
            } else {
                const uint32_t at = warp_append(LS.n_fresh, want);
                if (want && at < LS.cap) LS.fresh[at] = WalkItem{nb, no, nc, nc};
            }
            if (fresh) tc = 0;
            else { tb = nb; to = no; tc = nc; }
        }
    }
    if (COUNT) st.flush(P.stats);
}

} // namespace sbwt_b200
