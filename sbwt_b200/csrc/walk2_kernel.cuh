// walk2_kernel.cuh -- the colex-interval walk on the device (sm_100a), phase-sorted.
//
// Replaces SBWT::search (include/sbwt/SBWT.hh:390-415), SBWT::update_sbwt_interval
// (SBWT.hh:423-437) and SBWT::streaming_search (SBWT.hh:545-581) together with the per-read
// loops of src/CLI/sbwt_search.cpp:45-91.
//
// A k-mer's walk has two very different phases. Right after the table jump (kmer_prefix_precalc,
// SBWT.hh:404) the interval [l, r] is WIDE: a step needs rank_c(l) and rank_c(r+1), often from two
// sectors, and an absent k-mer dies here within a few steps. Once the interval is a SINGLETON
// (l == r) a step is one sector, one rank and one bit test -- and this is also exactly the
// streaming step of SBWT.hh:561-575, because in a reference-built index only suffix-group starts
// carry edges (SURVEY.md section 8(a) note 7): a set bit_c(col) proves col is its group's start.
// A warp whose lanes are in different phases pays for both on every trip, so the two phases are
// run as two lock-step loops, and every warp sorts its own work between them through small
// queues in shared memory (no global traffic, no second launch):
//
//   NARROW  32 lanes = 32 k-mers (consecutive k-mers of one read; in streaming mode the first
//           k-mers of 32 work items). Each lane loads its k-mer into registers, takes its table
//           row and runs general interval steps until its interval is empty (result -1), the
//           k-mer is complete, or the interval has been a singleton for kSingleHold + 1 steps.
//           Survivors go to a queue.
//   CHAIN   32 lanes = 32 queued survivors. Singleton steps only: first the remaining characters
//           of the survivor's own k-mer (SBWT.hh:425-436 on a singleton interval), then -- in
//           streaming mode -- the following k-mers of its work item, one step and one result
//           each (SBWT.hh:561-575). A clear bit there takes the literal walk-back over
//           suffix_group_starts (SBWT.hh:562-563) on a slow path. When a chain ends in a miss the
//           rest of the item goes to the warp's TODO queue and is answered by NARROW + CHAIN with
//           one lane per k-mer (streaming_search's answers equal search()'s, SURVEY.md 8(a) note 3).
//
//   PROBE   (streaming mode) runs of absent k-mers are not searched one by one. If the walk over
//           S = read[m .. m+j] dies at character j, no node's label ends with S, and because the
//           SBWT holds every prefix of every k-mer as (a suffix of) some node, no indexed k-mer
//           contains S anywhere: every k-mer of the read that covers [m, m+j] is absent, i.e. the
//           k-mers starting in [m+j-k+1, m]. So a range of k-mers left behind by a miss is probed at
//           every D-th k-mer only (D = P.probe_stride, about k - log4(n) - 2); probes that die early
//           enough prove their whole segment absent and the -1s are written with coalesced stores.
//           The first segment that is not proven absent restarts the read there as a fresh item
//           (FIRST pass: one full search, then the streaming chain again), exactly the control flow
//           of SBWT.hh:556-576 -- only the proofs of absence are cheaper. Results are identical.
//
// Work (32-k-mer chunks in search mode, work items in streaming mode) is handed out through one
// global cursor, so the grid is persistent and self-balancing.
#pragma once

#include <type_traits>

#include "device_index.cuh"
#include "walk_kernel.cuh"

namespace sbwt_b200 {

constexpr int kW2Threads = 256;
constexpr int kW2Warps = kW2Threads / 32;
constexpr int kQCap = 64; // entries per queue: at most 32 are waiting when up to 32 more are pushed
#ifndef SBWT_B200_W2_MINBLOCKS
#define SBWT_B200_W2_MINBLOCKS 4
#endif
#ifndef SBWT_B200_SINGLE_HOLD
#define SBWT_B200_SINGLE_HOLD 2
#endif
// SBWT_B200_DEBUG_NOSTORE (measurement knob) costs a uniform load and a branch at every store of the hot loop, so it
// only exists in builds made with -DSBWT_B200_DEBUG_KNOBS (tools/gpu_variants.sh)
#ifdef SBWT_B200_DEBUG_KNOBS
constexpr bool kDebugKnobs = true;
#else
constexpr bool kDebugKnobs = false;
#endif
#ifndef SBWT_B200_FAST_CHAIN
#define SBWT_B200_FAST_CHAIN 1
#endif
constexpr bool kFastChain = SBWT_B200_FAST_CHAIN != 0; // the steady-state loop of CHAIN (A/B: -DSBWT_B200_FAST_CHAIN=0)
constexpr uint32_t kSingleHold = SBWT_B200_SINGLE_HOLD; // extra NARROW steps on a singleton interval before it is queued (drops most chance survivors)

template <bool WIDE>
struct W2Queues {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    // survivors that only have their own k-mer left (search mode, and the TODO ranges of streaming mode)
    uint32_t v_base[kQCap], v_out[kQCap], v_meta[kQCap];
    pos_t v_col[kQCap];
    // survivors that go on streaming through their work item
    uint32_t s_base[kQCap], s_out[kQCap], s_meta[kQCap];
    pos_t s_col[kQCap];
    // ranges of k-mers to be answered one lane per k-mer
    uint32_t t_base[kQCap], t_out[kQCap], t_cnt[kQCap];
    // streaming mode: ranges (free of invalid bases) to be probed, and fresh items (restarts inside a read)
    uint32_t p_base[kQCap], p_out[kQCap], p_cnt[kQCap];
    uint32_t f_base[kQCap], f_out[kQCap], f_cnt[kQCap];
    uint32_t own[32]; // PROBE: lane -> (range, probe number)
};

// the k-mer starting at base b as 2-bit codes, 16 per word, character j at bits [2j, 2j+2) of the window
template <int KW>
struct KmerWin {
    uint32_t w[2 * KW];
};

template <int KW>
__device__ __forceinline__ KmerWin<KW> load_win(const uint32_t* __restrict__ codes, uint32_t b) {
    const uint32_t* p = codes + (b >> 4);
    const uint32_t sh = (b & 15u) * 2u;
    uint32_t x[2 * KW + 1];
#pragma unroll
    for (int i = 0; i < 2 * KW + 1; i++) x[i] = __ldg(p + i);
    KmerWin<KW> win;
#pragma unroll
    for (int i = 0; i < 2 * KW; i++) win.w[i] = __funnelshift_r(x[i], x[i + 1], sh);
    return win;
}

template <int KW>
__device__ __forceinline__ uint32_t win_char(const KmerWin<KW>& win, uint32_t j) {
    uint32_t w = win.w[0];
#pragma unroll
    for (int i = 1; i < 2 * KW; i++) w = ((j >> 4) == (uint32_t)i) ? win.w[i] : w;
    return (w >> ((j & 15u) * 2u)) & 3u;
}

// does the k-mer starting at base b cover a base outside ACGT (SBWT.hh:399,428)?
template <int KW>
__device__ __forceinline__ bool kmer_invalid(const uint32_t* __restrict__ inv, uint32_t b, int k) {
    const uint32_t* p = inv + (b >> 5);
    const uint32_t sh = b & 31u;
    const uint32_t f0 = __ldg(p), f1 = __ldg(p + 1);
    const uint32_t m0 = __funnelshift_r(f0, f1, sh);
    if (KW == 1) return (m0 & (k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u))) != 0;
    const uint32_t f2 = __ldg(p + 2);
    const uint32_t m1 = __funnelshift_r(f1, f2, sh);
    const int k1 = k - 32; // KW == 2 is used for 32 < k <= 64
    return (m0 | (m1 & (k1 >= 32 ? 0xFFFFFFFFu : ((1u << k1) - 1u)))) != 0;
}

template <bool OUT32>
__device__ __forceinline__ void store_result(const WalkParams& P, uint32_t o, int64_t v) {
    if (kDebugKnobs && P.debug_no_store) { // measurement only: 1 = no store, 2 = plain write-back store instead of the streaming one
        if (P.debug_no_store == 2) {
            if (OUT32) P.out32[o] = (int32_t)v;
            else P.out[o] = v;
        }
        return;
    }
    if (OUT32) asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(P.out32 + o), "r"((int32_t)v) : "memory");
    else asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(P.out + o), "l"(v) : "memory");
}

// One whole 32-byte sector of results (4 x int64 or 8 x int32) in one store (STG.E.256).
__device__ __forceinline__ void st_sector_cs(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// CHAIN writes one result per lane per step, each lane into its own read's slice of the output: as
// single stores those are 32 partial-sector writes per warp and step, which L2 has to merge (and
// fill from DRAM when the sector is gone again before its last part arrives: ncu r01d, 463 M
// write-lookup misses and 17 GB of extra DRAM reads per 9.6 GB of results). So every lane collects
// the results of one output sector in shared memory ([slot][lane]: conflict-free) and writes the
// sector with one 256-bit store; only the ragged ends of a lane's run are written one by one.
template <bool OUT32>
struct OutStage {
    typedef typename std::conditional<OUT32, int32_t, int64_t>::type val_t;
    static constexpr uint32_t R = OUT32 ? 8u : 4u; // results per sector
    val_t v[R][32];
};

// LITERAL (streaming mode on an index that violates "only suffix-group starts carry edges", i.e. a
// hand-made file): streaming answers may then differ from search() answers, so the reference's
// control flow is followed to the letter -- after a miss the k-mers are searched one at a time and
// streaming resumes from the first one found (SBWT.hh:556-576); invalid bases are met by the chain.
// COMPACT: rank steps are answered from the one-hot layout (device_index.cuh: one 32-byte csector holds all four
// characters of 96 columns); a block flagged there (some column with no edge or several) is answered from the classic
// sectors on a rare divergent path. Narrow, non-LITERAL kernels only; the host picks it when the index is eligible.
template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32, int KW, bool LITERAL, bool COMPACT>
__global__ void __launch_bounds__(kW2Threads, WIDE ? 3 : SBWT_B200_W2_MINBLOCKS) walk2_kernel(const WalkParams P) {
    static_assert(STREAMING || !LITERAL, "LITERAL is a streaming-mode variant");
    static_assert(!COMPACT || (!WIDE && !LITERAL), "the compact layout serves narrow indexes that keep the edge invariant");
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    __shared__ W2Queues<WIDE> queues[kW2Warps];
    W2Queues<WIDE>& Q = queues[threadIdx.x >> 5];
    __shared__ OutStage<OUT32> stages[kW2Warps];
    OutStage<OUT32>& OS = stages[threadIdx.x >> 5];
    constexpr uint32_t OR = OutStage<OUT32>::R;
    // phase of result 0 inside its sector
    const uint32_t oph = OUT32 ? (uint32_t)(((uintptr_t)P.out32 >> 2) & 7u) : (uint32_t)(((uintptr_t)P.out >> 3) & 3u);
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t n_items = (uint32_t)*P.n_items;
    const uint32_t k = (uint32_t)ix.k, p = (uint32_t)ix.tp; // p: characters answered by the search table
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const Sector* const sec_base = ix.sectors;
    uint64_t pol = make_l2_policy(P.index_evict_last != 0);
    const uint64_t pol_t = pol; // table rows
    if (P.index_evict_last == 2) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(P.l2_frac));
    if (P.index_evict_last == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(P.l2_frac));
    const Sector* const cmp_base = ix.compact;
    const uint32_t* const cbase = ix.cbase;
    const uint64_t pol_cl = COMPACT ? make_l2_policy(false) : pol; // classic sectors are the cold structure of a compact index
    const pos_t last_col = (pos_t)(ix.n_nodes - 1);
    constexpr uint32_t kGrab = STREAMING ? 32u : 8u; // items (streaming) or chunks (search) per cursor bump

    uint32_t nV = 0, nS = 0, nT = 0, nP = 0, nF = 0; // queue fill, warp-uniform
    uint32_t in_next = 0, in_end = 0;    // grabbed input range
    bool input_done = false;
    unsigned long long st_lookups = 0, st_hits = 0, st_ranks = 0, st_sectors = 0;
    // probe stride in k-mers (0 = no probing: every k-mer after a miss is searched on its own)
    const uint32_t D = (STREAMING && !LITERAL) ? P.probe_stride : 0u;

    while (true) {
        // ---- what to do next (warp-uniform)
        // Every S, P or F entry becomes at most one entry downstream (S -> P -> F -> S|P), so their
        // total only grows when input is taken, and input is taken only while it is <= 32: no queue
        // ever holds more than kQCap entries.
        enum : int { T_CHAIN_V, T_CHAIN_S, T_NARROW_TODO, T_NARROW_INPUT, T_PROBE, T_EXIT };
        int task;
        bool take_input = false;
        if (nV >= 32) task = T_CHAIN_V;
        else if (STREAMING && nS >= 32 && nT <= 32) task = T_CHAIN_S;
        else if (nT > 0) task = T_NARROW_TODO;
        else if (STREAMING && nP >= 16) task = T_PROBE;
        else if (STREAMING && nF >= 32) task = T_NARROW_INPUT;
        else if (!input_done && (!STREAMING || nS + nP + nF <= 32)) {
            if (in_next >= in_end) {
                uint32_t g = 0;
                if (lane == 0) g = (uint32_t)atomicAdd(P.cursor, (unsigned long long)kGrab);
                g = __shfl_sync(FULL, g, 0);
                if (g >= n_items) { input_done = true; continue; }
                in_next = g;
                in_end = min(g + kGrab, n_items);
            }
            task = T_NARROW_INPUT;
            take_input = true;
        } else if (STREAMING && nP > 0) task = T_PROBE;
        else if (STREAMING && nF > 0) task = T_NARROW_INPUT;
        else if (STREAMING && nS > 0) task = T_CHAIN_S;
        else if (nV > 0) task = T_CHAIN_V;
        else task = T_EXIT;
        if (task == T_EXIT) break;

        if (task == T_CHAIN_V || task == T_CHAIN_S) {
            // ================================================================ CHAIN
            const bool sb = STREAMING && task == T_CHAIN_S;
            uint32_t* const qb = sb ? Q.s_base : Q.v_base;
            uint32_t* const qo = sb ? Q.s_out : Q.v_out;
            uint32_t* const qm = sb ? Q.s_meta : Q.v_meta;
            pos_t* const qc = sb ? Q.s_col : Q.v_col;
            const uint32_t nq = sb ? nS : nV;
            const uint32_t m = min(32u, nq), q0 = nq - m;
            if (sb) nS = q0; else nV = q0;
            bool act = (uint32_t)lane < m;
            uint32_t o = 0, j = 0, rem = 0, pos = 0, cw = 0, nx = 0;
            pos_t col = 0;
            if (act) {
                const uint32_t b = qb[q0 + lane], meta = qm[q0 + lane];
                o = qo[q0 + lane];
                col = qc[q0 + lane];
                j = meta & 0xFFu;
                rem = meta >> 8;
                pos = b + j; // next base to consume
                cw = __ldg(P.codes + (pos >> 4));
                nx = __ldg(P.codes + (pos >> 4) + 1);
            }
            __syncwarp();
            bool fs = false; // the lane is past its first k-mer: its steps are streaming steps (SBWT.hh:561-575)
            uint32_t gs = o; // first result staged and not yet written
            // result number x of this lane -> stage; a completed sector goes out in one store
            auto emit = [&](uint32_t x, int64_t v) {
                if (!sb) { // one result per lane: nothing to collect
                    store_result<OUT32>(P, x, v);
                    gs = x + 1u;
                    return;
                }
                const uint32_t sl = (x + oph) & (OR - 1u);
                OS.v[sl][lane] = (typename OutStage<OUT32>::val_t)v;
                if (sl == OR - 1u) {
                    if (x - gs == OR - 1u) {
                        if (!(kDebugKnobs && P.debug_no_store)) {
                            uint32_t wv[8];
                            if (OUT32) {
#pragma unroll
                                for (int i = 0; i < 8; i++) wv[i] = (uint32_t)OS.v[i & (OR - 1u)][lane];
                                st_sector_cs(P.out32 + (x - 7u), wv);
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    const unsigned long long q = (unsigned long long)OS.v[i & (OR - 1u)][lane];
                                    wv[2 * i] = (uint32_t)q;
                                    wv[2 * i + 1] = (uint32_t)(q >> 32);
                                }
                                st_sector_cs(P.out + (x - 3u), wv);
                            }
                        }
                    } else {
                        for (uint32_t y = gs; y <= x; y++) store_result<OUT32>(P, y, (int64_t)OS.v[(y + oph) & (OR - 1u)][lane]);
                    }
                    gs = x + 1u;
                }
            };
            auto drain = [&](uint32_t end) { // the lane's run is over: write what is still staged, [gs, end)
                for (uint32_t y = gs; y < end; y++) store_result<OUT32>(P, y, (int64_t)OS.v[(y + oph) & (OR - 1u)][lane]);
                gs = end;
            };
            while (true) {
                if (act && j == k) { // a k-mer's interval is a singleton (SBWT.hh:410-413): its column is the answer
                    emit(o, (int64_t)col);
                    if (COUNT) { st_lookups++; st_hits++; }
                    o++;
                    rem--;
                    if (rem == 0) { act = false; drain(o); }
                    else { j = k - 1; fs = true; }
                }
                if (!__any_sync(FULL, act)) break;
                // ---- steady state of a streaming chain: every active lane sits on a found k-mer (SBWT.hh:561-575 with
                // l == r) and the next one is found as well. Those iterations need none of the bookkeeping below (pending
                // results, misses, queue pushes, walk-backs), so they run in a loop of their own: step, and if EVERY active
                // lane found its next k-mer, commit and emit at once. The first iteration in which some lane misses, meets a
                // flagged csector or needs anything else is left untouched (nothing committed) to the general code below.
                // (COMPACT kernels only: an index on classic sectors is HBM-bound, not issue-bound, and measured 7 % slower
                // with the extra loop: profiles/r01k_fast_chain.txt)
                if (STREAMING && !LITERAL && COMPACT && kFastChain && sb && __all_sync(FULL, !act || (fs && j == k - 1u))) {
                    while (true) {
                        bool ok = true;
                        pos_t ncol = 0;
                        if (act) {
                            const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                            if (COMPACT) {
                                const uint32_t cb = (uint32_t)col / (uint32_t)kCBlockCols, coff = (uint32_t)col - cb * (uint32_t)kCBlockCols;
                                const uint32_t base = ld_cbase(cbase, cb, c);
                                const Sector s = ld_sector(cmp_base + cb, pol);
                                const CompactRank cr = compact_rank(compact_match(s, c), base, coff);
                                ncol = (pos_t)cr.value;
                                ok = cr.bit != 0 && !(csector_flagged(s) | (base == 0xFFFFFFFFu)); // (base: see below)
                            } else {
                                const BlockPos bp = split_pos<WIDE>((int64_t)col);
                                const Sector s = ld_sector(sector_ptr<WIDE>(sec_base, bp.blk, c), pol_cl);
                                const SectorPrefix pf = sector_prefix(s);
                                const uint32_t f = bp.off >> 5, rm = bp.off & 31u;
                                const uint32_t w = sector_word(s, f);
                                ncol = (pos_t)(s.w[0] + __byte_perm(pf.X, pf.Y, f) + __popc(w & ((1u << rm) - 1u)));
                                if (WIDE) ncol += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (bp.blk >> ix.sb_shift));
                                ok = ((w >> rm) & 1u) != 0;
                            }
                        }
                        if (!__all_sync(FULL, ok)) break;
                        if (act) {
                            if (COUNT) { st_ranks += 2; st_sectors += 1; st_lookups++; st_hits++; }
                            col = ncol;
                            pos++;
                            if ((pos & 15u) == 0) { cw = nx; nx = __ldg(P.codes + (pos >> 4) + 1); }
                            emit(o, (int64_t)col);
                            o++;
                            rem--;
                            if (rem == 0) { act = false; drain(o); }
                        }
                        if (!__any_sync(FULL, act)) break;
                    }
                    if (!__any_sync(FULL, act)) break;
                }
                bool miss = false;
                if (act) {
                    const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                    pos_t ncol = 0;
                    int64_t cblk = -1; // classic block of col, when it was read
                    bool classic = true;
                    if (COMPACT) {
                        const uint32_t cb = (uint32_t)col / (uint32_t)kCBlockCols, coff = (uint32_t)col - cb * (uint32_t)kCBlockCols;
                        const uint32_t base = ld_cbase(cbase, cb, c); // in flight together with the csector
                        const Sector s = ld_sector(cmp_base + cb, pol);
                        // (base is never all ones; testing it here makes the branch wait for it, so ptxas cannot sink its load
                        // behind the branch, where it would add a second dependent memory latency to every step)
                        if (!(csector_flagged(s) | (base == 0xFFFFFFFFu))) {
                            const CompactRank cr = compact_rank(compact_match(s, c), base, coff);
                            ncol = (pos_t)cr.value;
                            miss = cr.bit == 0;
                            classic = false;
                        }
                    }
                    if (classic) {
                        const BlockPos bp = split_pos<WIDE>((int64_t)col);
                        cblk = bp.blk;
                        const Sector s = ld_sector(sector_ptr<WIDE>(sec_base, bp.blk, c), pol_cl);
                        const SectorPrefix pf = sector_prefix(s);
                        const uint32_t f = bp.off >> 5, rm = bp.off & 31u;
                        const uint32_t w = sector_word(s, f);
                        ncol = (pos_t)(s.w[0] + __byte_perm(pf.X, pf.Y, f) + __popc(w & ((1u << rm) - 1u)));
                        if (WIDE) ncol += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (bp.blk >> ix.sb_shift));
                        miss = ((w >> rm) & 1u) == 0; // [col, col] -> empty interval (SBWT.hh:433) / l != r (SBWT.hh:574)
                    }
                    if (COUNT) { st_ranks += 2; st_sectors += 1; }
                    bool bad = false; // LITERAL: the chain can run into a base outside ACGT (SBWT.hh:565-568)
                    if (LITERAL && fs) bad = ((__ldg(P.invalid + (pos >> 5)) >> (pos & 31u)) & 1u) != 0;
                    if (STREAMING && fs && !bad && (miss || !ix.edges_at_starts)) {
                        // literal form (SBWT.hh:562-563): the step starts from the suffix-group start of col
                        int64_t g = (int64_t)col;
                        while (true) {
                            const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
                            if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
                            g = (g & ~31ll) - 1;
                        }
                        if (COUNT) st_sectors++;
                        if (g != (int64_t)col) {
                            const BlockPos bs = split_pos<WIDE>(g);
                            const Sector ss = ld_sector(sector_ptr<WIDE>(sec_base, bs.blk, c), pol_cl);
                            if (COUNT) st_sectors += bs.blk != cblk;
                            miss = sector_bit(ss, bs.off) == 0;
                            ncol = (pos_t)lf_value<WIDE>(ix, ss, bs.blk, bs.off, c);
                        }
                    }
                    if (bad) miss = true;
                    if (!miss) {
                        col = ncol;
                        j++;
                        pos++;
                        if ((pos & 15u) == 0) { cw = nx; nx = __ldg(P.codes + (pos >> 4) + 1); }
                    }
                }
                const bool ended = act && miss;
                if (ended) {
                    emit(o, -1);
                    if (COUNT) st_lookups++;
                    o++;
                    rem--;
                    act = false;
                    drain(o);
                }
                if (STREAMING) { // the k-mers after a miss are searched from scratch (SBWT.hh:557-559), one lane each
                    const bool push = ended && rem > 0;
                    const unsigned pm = __ballot_sync(FULL, push);
                    if (pm) { // (the chain only covers k-mers free of invalid bases, so the range may be probed)
                        const uint32_t at = (D ? nP : nT) + __popc(pm & lt_mask);
                        if (push) {
                            if (D) { Q.p_base[at] = pos - j + 1; Q.p_out[at] = o; Q.p_cnt[at] = rem; }
                            else { Q.t_base[at] = pos - j + 1; Q.t_out[at] = o; Q.t_cnt[at] = rem; }
                        }
                        if (D) nP += __popc(pm); else nT += __popc(pm);
                        __syncwarp();
                    }
                }
            }
            continue;
        }

        // ==================================================================== NARROW
        // lanes = first k-mers of up to 32 work items or fresh items (LITERAL: also the next k-mer of a TODO
        // range, alone); or 32 k-mers of a TODO range / search chunk; or the probes of a few P ranges
        const bool probe = STREAMING && task == T_PROBE;
        const bool first = STREAMING && (task == T_NARROW_INPUT || LITERAL);
        bool act = false;
        uint32_t b = 0, o = 0, cnt = 1, nvalid = 1;
        uint32_t pr_tb = 0, pr_to = 0, pr_tc = 0, pr_np = 0, pr_S = 0, G = 0, seg_lo = 0, xm = 0; // PROBE
        if (probe) {
            // lanes 0 .. G-1 hold the top G ranges of P (as many as give <= 32 probes), then every lane takes one probe
            const uint32_t n_take = min(nP, 32u);
            if ((uint32_t)lane < n_take) {
                const uint32_t t = nP - 1 - lane;
                pr_tb = Q.p_base[t]; pr_to = Q.p_out[t]; pr_tc = Q.p_cnt[t];
                pr_np = min((pr_tc + D - 1) / D, 32u);
            }
            pr_S = pr_np;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, pr_S, s);
                if (lane >= s) pr_S += t;
            }
            G = __popc(__ballot_sync(FULL, (uint32_t)lane < n_take && pr_S <= 32u));
            const uint32_t total = __shfl_sync(FULL, pr_S, G - 1);
            __syncwarp();
            if ((uint32_t)lane < G)
                for (uint32_t t = 0; t < pr_np; t++) Q.own[pr_S - pr_np + t] = (uint32_t)lane | (t << 8);
            nP -= G;
            __syncwarp();
            act = (uint32_t)lane < total;
            const uint32_t ow = act ? Q.own[lane] : 0u;
            const uint32_t tb = __shfl_sync(FULL, pr_tb, ow & 0xFFu), to = __shfl_sync(FULL, pr_to, ow & 0xFFu);
            const uint32_t tc = __shfl_sync(FULL, pr_tc, ow & 0xFFu);
            seg_lo = (ow >> 8) * D;
            xm = min(seg_lo + D, tc) - 1u; // the probe is the last k-mer of its segment
            b = tb + xm;
            o = to + xm;
        } else if (task == T_NARROW_TODO) {
            const uint32_t t = nT - 1;
            const uint32_t tb = Q.t_base[t], to = Q.t_out[t], tc = Q.t_cnt[t];
            __syncwarp();
            if (LITERAL) { // one k-mer; its survivor streams on through the rest of the range
                act = lane == 0;
                b = tb; o = to; cnt = tc; nvalid = tc;
                nT = t;
            } else {
                act = (uint32_t)lane < tc;
                b = tb + lane;
                o = to + lane;
                if (tc > 32) {
                    if (lane == 0) { Q.t_base[t] = tb + 32; Q.t_out[t] = to + 32; Q.t_cnt[t] = tc - 32; }
                } else nT = t;
            }
            __syncwarp();
        } else if (!STREAMING) {
            const uint4 it = __ldg(reinterpret_cast<const uint4*>(P.items) + in_next);
            in_next++;
            act = (uint32_t)lane < it.z;
            b = it.x + lane;
            o = it.y + lane;
        } else {
            // fresh items first, then new work items
            const uint32_t n_f = min(nF, 32u);
            bool have = false;
            if ((uint32_t)lane < n_f) {
                const uint32_t t = nF - 1 - lane;
                b = Q.f_base[t]; o = Q.f_out[t]; cnt = Q.f_cnt[t]; nvalid = cnt;
                have = true;
            }
            nF -= n_f;
            if (take_input) {
                const uint32_t m = min(32u - n_f, in_end - in_next);
                if ((uint32_t)lane >= n_f && (uint32_t)lane - n_f < m) {
                    const uint4 it = __ldg(reinterpret_cast<const uint4*>(P.items) + in_next + ((uint32_t)lane - n_f));
                    b = it.x; o = it.y; cnt = it.z; nvalid = it.w;
                    have = true;
                }
                in_next += m;
            }
            act = have && nvalid > 0;
            // an item whose first k-mer covers an invalid base is answered one lane per k-mer
            const bool push = have && nvalid == 0;
            const unsigned pm = __ballot_sync(FULL, push);
            if (pm) {
                const uint32_t at = nT + __popc(pm & lt_mask);
                if (push) { Q.t_base[at] = b; Q.t_out[at] = o; Q.t_cnt[at] = cnt; }
                nT += __popc(pm);
            }
        }

        KmerWin<KW> win;
#pragma unroll
        for (int i = 0; i < 2 * KW; i++) win.w[i] = 0;
        bool alive = act;
        if (act) {
            win = load_win<KW>(P.codes, b);
            if (!probe && (!first || (LITERAL && task == T_NARROW_TODO))) alive = !kmer_invalid<KW>(P.invalid, b, (int)k);
        }
        pos_t l = 0, r = last_col;
        if (p != 0 && alive) {
            // first character = least significant digit of the table index (SBWT.hh:396-401)
            const TableRow<WIDE> row = TableRow<WIDE>::load(ix.table, win.w[0] & pmask, pol_t);
            l = (pos_t)row.l;
            r = (pos_t)row.r;
            if (row.absent()) alive = false;
            if (COUNT) st_sectors++;
        }
        // jl: characters consumed; when the walk dies, the string of the first jl + 1 characters is absent
        // (a row missing from the table: its p characters are)
        uint32_t jl = (p != 0 && act && !alive) ? p - 1u : p, single = (alive && l == r) ? 1u : 0u;
        while (true) {
            const bool go = alive && jl < k && single <= kSingleHold;
            if (!__any_sync(FULL, go)) break;
            if (go) {
                const int c = (int)win_char<KW>(win, jl);
                pos_t nl = 0, nr = 0;
                bool two = false, classic = true;
                if (COMPACT) {
                    const uint32_t p0 = (uint32_t)l, p1 = (uint32_t)r + 1u;
                    const uint32_t cb0 = p0 / (uint32_t)kCBlockCols, cb1 = p1 / (uint32_t)kCBlockCols;
                    two = cb1 != cb0;
                    const uint32_t base0 = ld_cbase(cbase, cb0, c); // in flight together with the csectors
                    uint32_t base1 = base0;
                    if ((cb1 >> kCSbShift) != (cb0 >> kCSbShift)) base1 = ld_cbase(cbase, cb1, c);
                    const Sector s0 = ld_sector(cmp_base + cb0, pol);
                    Sector s1;
                    if (two) s1 = ld_sector(cmp_base + cb1, pol);
                    if (!(csector_flagged(s0) | (two && csector_flagged(s1)) | ((base0 & base1) == 0xFFFFFFFFu))) { // (bases: see CHAIN)
                        const CompactMatch m0 = compact_match(s0, c);
                        nl = (pos_t)compact_rank(m0, base0, p0 - cb0 * (uint32_t)kCBlockCols).value;
                        if (two) nr = (pos_t)compact_rank(compact_match(s1, c), base1, p1 - cb1 * (uint32_t)kCBlockCols).value;
                        else nr = (pos_t)compact_rank(m0, base0, p1 - cb0 * (uint32_t)kCBlockCols).value;
                        classic = false;
                    }
                }
                if (classic) {
                    const BlockPos b0 = split_pos<WIDE>((int64_t)l), b1 = split_pos<WIDE>((int64_t)r + 1);
                    two = b1.blk != b0.blk;
                    const Sector s0 = ld_sector(sector_ptr<WIDE>(sec_base, b0.blk, c), pol_cl);
                    Sector s1;
                    if (two) s1 = ld_sector(sector_ptr<WIDE>(sec_base, b1.blk, c), pol_cl);
                    const SectorPrefix pf0 = sector_prefix(s0);
                    nl = (pos_t)sector_rank_fast(s0, pf0, b0.off);
                    if (two) {
                        const SectorPrefix pf1 = sector_prefix(s1);
                        nr = (pos_t)sector_rank_fast(s1, pf1, b1.off);
                    } else {
                        nr = (pos_t)sector_rank_fast(s0, pf0, b1.off);
                    }
                    if (WIDE) {
                        nl += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b0.blk >> ix.sb_shift));
                        nr += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (b1.blk >> ix.sb_shift));
                    }
                }
                nr -= 1;
                if (COUNT) { st_ranks += 2; st_sectors += two ? 2 : 1; }
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
        if (probe) {
            // a dead probe proves the k-mers [xm + jl + 1 - k, xm] of its range absent; "covered": its whole segment
            const bool covered = dead && (int)(xm + jl + 1u) - (int)k <= (int)seg_lo;
            const unsigned badmask = __ballot_sync(FULL, act && !covered);
            uint32_t nfill = 0;
            bool cont = false, fresh = false;
            if ((uint32_t)lane < G) {
                const uint32_t lo = pr_S - pr_np;
                const unsigned bm = badmask & ((pr_np >= 32u ? FULL : ((1u << pr_np) - 1u)) << lo);
                if (bm) { // restart the read at the first segment that is not proven absent
                    nfill = (uint32_t)(__ffs(bm) - 1 - (int)lo) * D;
                    cont = fresh = true;
                } else { // all probed segments are absent; a range longer than 32 segments goes on being probed
                    nfill = min(pr_tc, pr_np * D);
                    cont = nfill < pr_tc;
                }
            }
            const unsigned fm = __ballot_sync(FULL, cont && fresh), rm = __ballot_sync(FULL, cont && !fresh);
            if (fm) {
                const uint32_t at = nF + __popc(fm & lt_mask);
                if (cont && fresh) { Q.f_base[at] = pr_tb + nfill; Q.f_out[at] = pr_to + nfill; Q.f_cnt[at] = pr_tc - nfill; }
                nF += __popc(fm);
            }
            if (rm) {
                const uint32_t at = nP + __popc(rm & lt_mask);
                if (cont && !fresh) { Q.p_base[at] = pr_tb + nfill; Q.p_out[at] = pr_to + nfill; Q.p_cnt[at] = pr_tc - nfill; }
                nP += __popc(rm);
            }
            for (uint32_t g = 0; g < G; g++) { // the proven misses, 32 consecutive results per store
                const uint32_t to = __shfl_sync(FULL, pr_to, g), nf = __shfl_sync(FULL, nfill, g);
                for (uint32_t x = lane; x < nf; x += 32) {
                    store_result<OUT32>(P, to + x, -1);
                    if (COUNT) st_lookups++;
                }
            }
            __syncwarp();
            continue;
        }
        if (dead) {
            store_result<OUT32>(P, o, -1);
            if (COUNT) st_lookups++;
        }
        if (!first) {
            if (alive && jl == k) { // complete: the interval is a singleton (SBWT.hh:410-413)
                store_result<OUT32>(P, o, (int64_t)l);
                if (COUNT) { st_lookups++; st_hits++; }
            }
            const bool surv = alive && jl < k;
            const unsigned sm = __ballot_sync(FULL, surv);
            if (sm) {
                const uint32_t at = nV + __popc(sm & lt_mask);
                if (surv) { Q.v_base[at] = b; Q.v_out[at] = o; Q.v_meta[at] = jl | (1u << 8); Q.v_col[at] = l; }
                nV += __popc(sm);
            }
        } else {
            // what streaming cannot reach is answered one lane per k-mer: everything after a first
            // k-mer that is absent, and everything from the first k-mer that covers an invalid base on
            // (with probing: the k-mers up to the first invalid base are probed, the rest is answered per k-mer)
            const bool pd = dead && cnt > 1 && D == 0, pa = (alive || (dead && D != 0)) && nvalid < cnt;
            const unsigned tm = __ballot_sync(FULL, pd || pa);
            if (tm) {
                const uint32_t at = nT + __popc(tm & lt_mask);
                const uint32_t skip = pd ? 1u : nvalid;
                if (pd || pa) { Q.t_base[at] = b + skip; Q.t_out[at] = o + skip; Q.t_cnt[at] = cnt - skip; }
                nT += __popc(tm);
            }
            const bool pp = dead && D != 0 && nvalid > 1;
            const unsigned ppm = __ballot_sync(FULL, pp);
            if (ppm) {
                const uint32_t at = nP + __popc(ppm & lt_mask);
                if (pp) { Q.p_base[at] = b + 1; Q.p_out[at] = o + 1; Q.p_cnt[at] = nvalid - 1; }
                nP += __popc(ppm);
            }
            const unsigned sm = __ballot_sync(FULL, alive);
            if (sm) {
                const uint32_t at = nS + __popc(sm & lt_mask);
                if (alive) { Q.s_base[at] = b; Q.s_out[at] = o; Q.s_meta[at] = jl | (nvalid << 8); Q.s_col[at] = l; }
                nS += __popc(sm);
            }
        }
        __syncwarp();
    }

    if (COUNT) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            st_lookups += __shfl_xor_sync(FULL, st_lookups, s);
            st_hits += __shfl_xor_sync(FULL, st_hits, s);
            st_ranks += __shfl_xor_sync(FULL, st_ranks, s);
            st_sectors += __shfl_xor_sync(FULL, st_sectors, s);
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
