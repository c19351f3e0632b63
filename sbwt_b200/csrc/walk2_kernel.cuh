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
#define SBWT_B200_W2_MINBLOCKS 6
#endif
#ifndef SBWT_B200_SINGLE_HOLD
#define SBWT_B200_SINGLE_HOLD 2
#endif
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
    if (P.debug_no_store) return;
    if (OUT32) asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(P.out32 + o), "r"((int32_t)v) : "memory");
    else asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(P.out + o), "l"(v) : "memory");
}

// LITERAL (streaming mode on an index that violates "only suffix-group starts carry edges", i.e. a
// hand-made file): streaming answers may then differ from search() answers, so the reference's
// control flow is followed to the letter -- after a miss the k-mers are searched one at a time and
// streaming resumes from the first one found (SBWT.hh:556-576); invalid bases are met by the chain.
template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32, int KW, bool LITERAL>
__global__ void __launch_bounds__(kW2Threads, WIDE ? 3 : SBWT_B200_W2_MINBLOCKS) walk2_kernel(const WalkParams P) {
    static_assert(STREAMING || !LITERAL, "LITERAL is a streaming-mode variant");
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    __shared__ W2Queues<WIDE> queues[kW2Warps];
    W2Queues<WIDE>& Q = queues[threadIdx.x >> 5];
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t n_items = (uint32_t)*P.n_items;
    const uint32_t k = (uint32_t)ix.k, p = (uint32_t)ix.tp; // p: characters answered by the search table
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const Sector* const sec_base = ix.sectors;
    const uint64_t pol = make_l2_policy(P.index_evict_last != 0);
    const pos_t last_col = (pos_t)(ix.n_nodes - 1);
    constexpr uint32_t kGrab = STREAMING ? 32u : 8u; // items (streaming) or chunks (search) per cursor bump

    uint32_t nV = 0, nS = 0, nT = 0;     // queue fill, warp-uniform
    uint32_t in_next = 0, in_end = 0;    // grabbed input range
    bool input_done = false;
    unsigned long long st_lookups = 0, st_hits = 0, st_ranks = 0, st_sectors = 0;

    while (true) {
        // ---- what to do next (warp-uniform)
        enum : int { T_CHAIN_V, T_CHAIN_S, T_NARROW_TODO, T_NARROW_INPUT, T_EXIT };
        int task;
        if (nV >= 32) task = T_CHAIN_V;
        else if (STREAMING && nS >= 32 && nT <= 32) task = T_CHAIN_S;
        else if (nT > 0) task = T_NARROW_TODO;
        else if (!input_done) {
            if (in_next >= in_end) {
                uint32_t g = 0;
                if (lane == 0) g = (uint32_t)atomicAdd(P.cursor, (unsigned long long)kGrab);
                g = __shfl_sync(FULL, g, 0);
                if (g >= n_items) { input_done = true; continue; }
                in_next = g;
                in_end = min(g + kGrab, n_items);
            }
            task = T_NARROW_INPUT;
        } else if (STREAMING && nS > 0) task = T_CHAIN_S;
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
            while (true) {
                if (act && j == k) { // a k-mer's interval is a singleton (SBWT.hh:410-413): its column is the answer
                    store_result<OUT32>(P, o, (int64_t)col);
                    if (COUNT) { st_lookups++; st_hits++; }
                    o++;
                    rem--;
                    if (rem == 0) act = false;
                    else { j = k - 1; fs = true; }
                }
                if (!__any_sync(FULL, act)) break;
                bool miss = false;
                if (act) {
                    const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                    const BlockPos bp = split_pos<WIDE>((int64_t)col);
                    const Sector s = ld_sector(sector_ptr<WIDE>(sec_base, bp.blk, c), pol);
                    const SectorPrefix pf = sector_prefix(s);
                    const uint32_t f = bp.off >> 5, rm = bp.off & 31u;
                    const uint32_t w = sector_word(s, f);
                    pos_t ncol = (pos_t)(s.w[0] + __byte_perm(pf.X, pf.Y, f) + __popc(w & ((1u << rm) - 1u)));
                    if (WIDE) ncol += (pos_t)__ldg(ix.sbbase + (int64_t)c * ix.n_sb + (bp.blk >> ix.sb_shift));
                    miss = ((w >> rm) & 1u) == 0; // [col, col] -> empty interval (SBWT.hh:433) / l != r (SBWT.hh:574)
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
                            const Sector ss = ld_sector(sector_ptr<WIDE>(sec_base, bs.blk, c), pol);
                            if (COUNT) st_sectors += bs.blk != bp.blk;
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
                    store_result<OUT32>(P, o, -1);
                    if (COUNT) st_lookups++;
                    o++;
                    rem--;
                    act = false;
                }
                if (STREAMING) { // the k-mers after a miss are searched from scratch (SBWT.hh:557-559), one lane each
                    const bool push = ended && rem > 0;
                    const unsigned pm = __ballot_sync(FULL, push);
                    if (pm) {
                        const uint32_t at = nT + __popc(pm & lt_mask);
                        if (push) { Q.t_base[at] = pos - j + 1; Q.t_out[at] = o; Q.t_cnt[at] = rem; }
                        nT += __popc(pm);
                        __syncwarp();
                    }
                }
            }
            continue;
        }

        // ==================================================================== NARROW
        // lanes = first k-mers of 32 work items (LITERAL: also the next k-mer of a TODO range, alone)
        const bool first = STREAMING && (task == T_NARROW_INPUT || LITERAL);
        bool act = false;
        uint32_t b = 0, o = 0, cnt = 1, nvalid = 1;
        if (task == T_NARROW_TODO) {
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
            const uint32_t idx = in_next + lane;
            const bool have = idx < in_end;
            in_next = in_end;
            if (have) {
                const uint4 it = __ldg(reinterpret_cast<const uint4*>(P.items) + idx);
                b = it.x; o = it.y; cnt = it.z; nvalid = it.w;
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
            if (!first || (LITERAL && task == T_NARROW_TODO)) alive = !kmer_invalid<KW>(P.invalid, b, (int)k);
        }
        pos_t l = 0, r = last_col;
        if (p != 0 && alive) {
            // first character = least significant digit of the table index (SBWT.hh:396-401)
            const TableRow<WIDE> row = TableRow<WIDE>::load(ix.table, win.w[0] & pmask, pol);
            l = (pos_t)row.l;
            r = (pos_t)row.r;
            if (row.absent()) alive = false;
            if (COUNT) st_sectors++;
        }
        uint32_t jl = p, single = (alive && l == r) ? 1u : 0u;
        while (true) {
            const bool go = alive && jl < k && single <= kSingleHold;
            if (!__any_sync(FULL, go)) break;
            if (go) {
                const int c = (int)win_char<KW>(win, jl);
                const BlockPos b0 = split_pos<WIDE>((int64_t)l), b1 = split_pos<WIDE>((int64_t)r + 1);
                const bool two = b1.blk != b0.blk;
                const Sector s0 = ld_sector(sector_ptr<WIDE>(sec_base, b0.blk, c), pol);
                Sector s1;
                if (two) s1 = ld_sector(sector_ptr<WIDE>(sec_base, b1.blk, c), pol);
                const SectorPrefix pf0 = sector_prefix(s0);
                pos_t nl = (pos_t)sector_rank_fast(s0, pf0, b0.off);
                pos_t nr;
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
            const bool pd = dead && cnt > 1, pa = alive && nvalid < cnt;
            const unsigned tm = __ballot_sync(FULL, pd || pa);
            if (tm) {
                const uint32_t at = nT + __popc(tm & lt_mask);
                const uint32_t skip = pd ? 1u : nvalid;
                if (pd || pa) { Q.t_base[at] = b + skip; Q.t_out[at] = o + skip; Q.t_cnt[at] = cnt - skip; }
                nT += __popc(tm);
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
