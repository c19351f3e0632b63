// walk_kernel.cuh -- the colex-interval walk on the device (sm_100a), phase-sorted.
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
// A warp whose lanes are in different phases pays for both on every trip, so the phases are run as
// separate lock-step loops, and every warp sorts its own work between them through small queues in
// shared memory (no global traffic, no second launch):
//
//   NARROW  32 lanes = 32 k-mers (consecutive k-mers of one read; in streaming mode the first
//           k-mers of 32 work items). Each lane loads its k-mer into registers, takes its table
//           row and runs general interval steps until its interval is empty (result -1), the
//           k-mer is complete, or the interval has been a singleton for kSingleHold + 1 steps.
//           Survivors go to a queue.
//   CHAIN   singleton steps only. Search mode (and the stragglers of streaming mode): 32 lanes = 32
//           survivors finishing their own k-mer (SBWT.hh:425-436 on a singleton interval).
//           Streaming mode: every lane owns kChains survivors at once (their loads are in flight
//           together) and follows each through the rest of its own k-mer and then through the
//           following k-mers of its work item, one step and one result each (SBWT.hh:561-575).
//           The lanes first bring their chains to a common output phase (a few single stores), then
//           run a steady-state loop in which every chain steps, the results of one output sector
//           collect in registers and leave as one 32-byte store, and nothing else happens; a clear
//           bit (a miss, or a column that is not its suffix group's start) or a flagged csector
//           drops to a slow path that takes the literal walk-back over suffix_group_starts
//           (SBWT.hh:562-563). When a chain ends in a miss the rest of the item goes to PROBE.
//   PROBE   (streaming mode) runs of absent k-mers are not searched one by one. If the walk over
//           S = read[m .. m+j] dies at character j, no node's label ends with S, and because the
//           SBWT holds every prefix of every k-mer as (a suffix of) some node, no indexed k-mer
//           contains S anywhere: every k-mer of the read that covers [m, m+j] is absent, i.e. the
//           k-mers starting in [m+j-k+1, m]. So a range of k-mers left behind by a miss is probed at
//           every D-th k-mer only (D = P.probe_stride, about k - log4(n) - 2); probes that die early
//           enough prove their whole segment absent and the -1s are written with coalesced stores.
//           The first segment that is not proven absent restarts the read there as a fresh item
//           (one full search, then the streaming chain again), exactly the control flow of
//           SBWT.hh:556-576 -- only the proofs of absence are cheaper. Results are identical.
//
// Work (32-k-mer chunks in search mode, work items in streaming mode) is handed out through one
// global cursor, so the grid is persistent and self-balancing.
#pragma once

#include <type_traits>

#include "device_index.cuh"

namespace sbwt_b200 {

struct WalkParams {
    DeviceIndexView ix;
    const uint32_t* codes;    // 2-bit bases, 16 per u32 word
    const uint32_t* invalid;  // 1 bit per base, 32 per u32 word
    const WalkItem* items;
    const int64_t* n_items;   // device scalar
    int64_t* out;             // results as int64 ...
    int32_t* out32;           // ... or as int32 (OUT32 kernels; callers use them only when n_nodes < 2^31)
    unsigned long long* stats; // [lookups, hits, rank_ops, sectors] (COUNT only)
    unsigned long long* cursor; // next unclaimed work item (zeroed before every launch)
    int table_streams;         // 1: the search table is far larger than L2 -> its rows are read with evict_first
    uint32_t probe_stride;     // streaming mode: distance in k-mers between the probes of a range of presumed
                               // misses (0 = every k-mer after a miss is searched on its own)
};

// one row {l, r} of the search table; absent rows are all ones
template <bool WIDE>
struct TableRow;
template <>
struct TableRow<false> {
    uint32_t l, r;
    __device__ __forceinline__ bool absent() const { return l == 0xFFFFFFFFu; }
    static __device__ __forceinline__ TableRow load(const void* table, uint32_t idx, uint64_t pol) {
        TableRow t;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                     : "=r"(t.l), "=r"(t.r)
                     : "l"(reinterpret_cast<const uint2*>(table) + idx), "l"(pol));
        return t;
    }
};
template <>
struct TableRow<true> {
    int64_t l, r;
    __device__ __forceinline__ bool absent() const { return l < 0; }
    static __device__ __forceinline__ TableRow load(const void* table, uint32_t idx, uint64_t pol) {
        TableRow t;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
                     : "=l"(t.l), "=l"(t.r)
                     : "l"(reinterpret_cast<const longlong2*>(table) + idx), "l"(pol));
        return t;
    }
};

constexpr int kWalkThreads = 256;
constexpr int kWalkWarps = kWalkThreads / 32;
// resident blocks per SM the kernels are compiled for: 4 (64 registers) on the classic sectors, whose walks are bound by
// DRAM line fetches and want warps; 3 (80 registers, no spills) for the streaming walk on the one-hot layouts
// (profiles/r02f: c2 9.13 -> 8.72 ms on csector64, 11.0 -> 8.40 ms on csector96; c4s 17.7 vs 19.9 ms the other way)
#ifndef SBWT_B200_WALK_MINBLOCKS
#define SBWT_B200_WALK_MINBLOCKS 4
#endif
#ifndef SBWT_B200_WALK_MINBLOCKS_COMPACT
#define SBWT_B200_WALK_MINBLOCKS_COMPACT 3
#endif
template <bool STREAMING, bool WIDE, int LAY>
struct WalkBlocks {
    static constexpr int value = WIDE ? 3 : ((STREAMING && LAY != LAY_CLASSIC) ? SBWT_B200_WALK_MINBLOCKS_COMPACT : SBWT_B200_WALK_MINBLOCKS);
};
#ifndef SBWT_B200_SINGLE_HOLD
#define SBWT_B200_SINGLE_HOLD 2
#endif
#ifndef SBWT_B200_CHAINS
#define SBWT_B200_CHAINS 1
#endif
constexpr uint32_t kSingleHold = SBWT_B200_SINGLE_HOLD; // extra NARROW steps on a singleton interval before it is queued (drops most chance survivors)

// chains per lane of the streaming CHAIN phase (wide indexes keep one: their columns and results are 64-bit registers)
template <bool STREAMING, bool WIDE, bool LITERAL>
struct ChainCount {
    static constexpr int value = (STREAMING && !LITERAL && !WIDE) ? SBWT_B200_CHAINS : 1;
};

template <bool WIDE, int NCH>
struct WalkQueues {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    static constexpr int kCapV = 64;            // at most 32 are waiting when up to 32 more are pushed
    static constexpr int kCap = 32 + 32 * NCH;  // S + P + F never exceed this together (see the scheduler)
    // survivors that only have their own k-mer left (search mode, and the TODO ranges of streaming mode)
    uint32_t v_base[kCapV], v_out[kCapV], v_meta[kCapV];
    pos_t v_col[kCapV];
    // survivors that go on streaming through their work item
    uint32_t s_base[kCap], s_out[kCap], s_meta[kCap];
    pos_t s_col[kCap];
    // ranges of k-mers to be answered one lane per k-mer
    uint32_t t_base[kCap], t_out[kCap], t_cnt[kCap];
    // streaming mode: ranges (free of invalid bases) to be probed, and fresh items (restarts inside a read)
    uint32_t p_base[kCap], p_out[kCap], p_cnt[kCap];
    uint32_t f_base[kCap], f_out[kCap], f_cnt[kCap];
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
#ifdef SBWT_B200_DEBUG_NOSTORE // measurement builds only (tools/build_variants.sh): what the result stream costs
    if (v != -12345) return;
#endif
    if (OUT32) asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(P.out32 + o), "r"((int32_t)v) : "memory");
    else asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(P.out + o), "l"(v) : "memory");
}

// One whole 32-byte sector of results (4 x int64 or 8 x int32) in one store (STG.E.256).
__device__ __forceinline__ void st_sector_cs(void* p, const uint32_t (&w)[8]) {
#ifdef SBWT_B200_DEBUG_NOSTORE
    if (w[0] != 0xFFFFFFF7u) return;
#endif
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// The LITERAL chain writes one result per lane per step, each lane into its own read's slice of the output: as
// single stores those are 32 partial-sector writes per warp and step, which L2 has to merge (and
// fill from DRAM when the sector is gone again before its last part arrives: ncu r01d, 463 M
// write-lookup misses and 17 GB of extra DRAM reads per 9.6 GB of results). So every lane collects
// the results of one output sector in shared memory ([slot][lane]: conflict-free) and writes the
// sector with one 256-bit store; only the ragged ends of a lane's run are written one by one.
// (The non-LITERAL streaming chain keeps them in registers instead, see below.)
template <bool OUT32>
struct OutStage {
    typedef typename std::conditional<OUT32, int32_t, int64_t>::type val_t;
    static constexpr uint32_t R = OUT32 ? 8u : 4u; // results per sector
    val_t v[R][32];
};
struct NoStage {};

// Stage of the streaming CHAIN: the results of one round (kRound steps), one row per chain, 32 rows per chain slot of
// the lanes. Rows are padded to an odd number of 64-bit units, so that both the per-step writes (32 lanes, one column)
// and the flush (a few rows, consecutive columns) spread over the banks.
template <bool WIDE, int NCH>
struct ChainStage {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
#ifndef SBWT_B200_ROUND
#define SBWT_B200_ROUND 16
#endif
    static constexpr uint32_t kRound = WIDE ? 8u : (uint32_t)SBWT_B200_ROUND; // results per chain and round: one 128-byte line of int64 (narrow) / 64 bytes
    static constexpr uint32_t kStride = kRound + (WIDE ? 1u : 2u);
    pos_t v[NCH][32][kStride];
    uint32_t obase[NCH][32]; // index of the result that slot 0 of the row stands for
    uint32_t mask[NCH][32];  // slots of the row that hold a result of this round

    // all 32 rows of chain slot i -> global memory; lanes take consecutive results of a row, so a row leaves in one piece
    template <bool OUT32>
    static __device__ __forceinline__ void flush(ChainStage& S, int i, int lane, int64_t* out, int32_t* out32) {
        if (WIDE) { // 8 int64 per row: 8 lanes per row, 4 rows per pass
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int row = 4 * j + (lane >> 3), e = lane & 7;
                if ((S.mask[i][row] >> e) & 1u)
                    asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(out + (S.obase[i][row] + (uint32_t)e)), "l"((int64_t)S.v[i][row][e]) : "memory");
            }
        } else { // kRound u32 per row, two per lane: kRound / 2 lanes per row
            constexpr int LPR = (int)kRound / 2, RPP = 32 / LPR; // lanes per row, rows per pass
#pragma unroll
            for (int j = 0; j < LPR; j++) {
                const int row = RPP * j + lane / LPR, e = (lane % LPR) * 2;
                const uint32_t m2 = (S.mask[i][row] >> e) & 3u;
                if (m2) {
                    const uint32_t a = (uint32_t)S.v[i][row][e], b = (uint32_t)S.v[i][row][e + 1];
                    const uint32_t x = S.obase[i][row] + (uint32_t)e;
                    if (OUT32) {
                        if (m2 == 3u) asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(out32 + x), "r"(a), "r"(b) : "memory");
                        else if (m2 == 1u) asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(out32 + x), "r"(a) : "memory");
                        else asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(out32 + x + 1), "r"(b) : "memory");
                    } else { // int64 results of a narrow index: a column (< 2^32) or -1
                        const uint32_t ah = a == 0xFFFFFFFFu ? 0xFFFFFFFFu : 0u, bh = b == 0xFFFFFFFFu ? 0xFFFFFFFFu : 0u;
                        if (m2 == 3u) asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(out + x), "r"(a), "r"(ah), "r"(b), "r"(bh) : "memory");
                        else if (m2 == 1u) asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(out + x), "r"(a), "r"(ah) : "memory");
                        else asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(out + x + 1), "r"(b), "r"(bh) : "memory");
                    }
                }
            }
        }
    }
};

// ------------------------------------------------------------------ interval steps on the three layouts
//
// issue()/eval() split one singleton step into its memory request and its arithmetic, so that a lane can have the
// requests of several chains in flight before it touches the first answer. eval() returns false when the answer has
// to come from the classic sectors instead (a flagged csector); classic_step() gives it.
template <bool WIDE, int LAY>
struct Stepper {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    const Sector* sec;     // classic sectors
    const Sector* cmp;     // csectors (LAY_C96 / LAY_C64)
    const uint32_t* cbase; // LAY_C96
    const int64_t* sbbase; // WIDE
    int64_t n_sb;
    int sb_shift;
    uint64_t pol;          // L2 policy of the structure the walk lives on
    uint64_t pol_cold;     // classic sectors of a compact index (the rarely read structure)

    struct Load {
        Sector s;
        uint32_t base; // LAY_C96: cbase word
        int64_t wbase; // WIDE: superblock base
    };

    __device__ __forceinline__ void issue(Load& L, pos_t col, int c) const {
        if (LAY == LAY_C64) {
            L.s = ld_sector(cmp + ((uint32_t)col >> 6), pol);
        } else if (LAY == LAY_C96) {
            const uint32_t cb = (uint32_t)col / (uint32_t)kCBlockCols;
            L.base = ld_cbase(cbase, cb, c); // in flight together with the csector
            L.s = ld_sector(cmp + cb, pol);
        } else {
            const BlockPos bp = split_pos<WIDE>((int64_t)col);
            if (WIDE) L.wbase = __ldg(sbbase + (int64_t)c * n_sb + (bp.blk >> sb_shift));
            L.s = ld_sector(sector_ptr<WIDE>(sec, bp.blk, c), pol);
        }
    }

    __device__ __forceinline__ bool eval(const Load& L, pos_t col, int c, pos_t& ncol, bool& hit) const {
        if (LAY == LAY_C64) {
            const CompactRank cr = c64_rank(L.s, c, (uint32_t)col & 63u);
            ncol = (pos_t)cr.value;
            hit = cr.bit != 0;
            return !c64_flagged(L.s);
        } else if (LAY == LAY_C96) {
            const uint32_t cb = (uint32_t)col / (uint32_t)kCBlockCols, coff = (uint32_t)col - cb * (uint32_t)kCBlockCols;
            const CompactRank cr = compact_rank(compact_match(L.s, c), L.base, coff);
            ncol = (pos_t)cr.value;
            hit = cr.bit != 0;
            // (base is never all ones; testing it makes the flag test wait for it, so ptxas cannot sink its load behind a
            // branch on the csector, where it would add a second dependent memory latency to every step)
            return !(csector_flagged(L.s) | (L.base == 0xFFFFFFFFu));
        } else {
            const BlockPos bp = split_pos<WIDE>((int64_t)col);
            const SectorPrefix pf = sector_prefix(L.s);
            const uint32_t f = bp.off >> 5, rm = bp.off & 31u;
            const uint32_t w = sector_word(L.s, f);
            ncol = (pos_t)(L.s.w[0] + __byte_perm(pf.X, pf.Y, f) + __popc(w & ((1u << rm) - 1u)));
            if (WIDE) ncol += (pos_t)L.wbase;
            hit = ((w >> rm) & 1u) != 0;
            return true;
        }
    }

    // the same step from the classic sectors (every layout has them); blk = the block that was read
    __device__ __forceinline__ void classic_step(pos_t col, int c, pos_t& ncol, bool& hit, int64_t& blk) const {
        const BlockPos bp = split_pos<WIDE>((int64_t)col);
        blk = bp.blk;
        const Sector s = ld_sector(sector_ptr<WIDE>(sec, bp.blk, c), LAY == LAY_CLASSIC ? pol : pol_cold);
        int64_t v = (int64_t)sector_rank(s, bp.off);
        if (WIDE) v += __ldg(sbbase + (int64_t)c * n_sb + (bp.blk >> sb_shift));
        ncol = (pos_t)v;
        hit = sector_bit(s, bp.off) != 0;
    }

    // general step [l, r] -> [C[c] + rank_c(l), C[c] + rank_c(r + 1) - 1]; returns the sectors it had to read (1 or 2)
    __device__ __forceinline__ uint32_t narrow(pos_t l, pos_t r, int c, pos_t& nl, pos_t& nr) const {
        bool two = false, classic = true;
        if (LAY == LAY_C64) {
            const uint32_t p0 = (uint32_t)l, p1 = (uint32_t)r + 1u;
            const uint32_t cb0 = p0 >> 6, cb1 = p1 >> 6;
            two = cb1 != cb0;
            const Sector s0 = ld_sector(cmp + cb0, pol);
            Sector s1;
            if (two) s1 = ld_sector(cmp + cb1, pol);
            if (!(c64_flagged(s0) | (two && c64_flagged(s1)))) {
                nl = (pos_t)c64_rank(s0, c, p0 & 63u).value;
                nr = (pos_t)(two ? c64_rank(s1, c, p1 & 63u).value : c64_rank(s0, c, p1 & 63u).value);
                classic = false;
            }
        } else if (LAY == LAY_C96) {
            const uint32_t p0 = (uint32_t)l, p1 = (uint32_t)r + 1u;
            const uint32_t cb0 = p0 / (uint32_t)kCBlockCols, cb1 = p1 / (uint32_t)kCBlockCols;
            two = cb1 != cb0;
            const uint32_t base0 = ld_cbase(cbase, cb0, c); // in flight together with the csectors
            uint32_t base1 = base0;
            if ((cb1 >> kCSbShift) != (cb0 >> kCSbShift)) base1 = ld_cbase(cbase, cb1, c);
            const Sector s0 = ld_sector(cmp + cb0, pol);
            Sector s1;
            if (two) s1 = ld_sector(cmp + cb1, pol);
            if (!(csector_flagged(s0) | (two && csector_flagged(s1)) | ((base0 & base1) == 0xFFFFFFFFu))) { // (bases: see eval)
                const CompactMatch m0 = compact_match(s0, c);
                nl = (pos_t)compact_rank(m0, base0, p0 - cb0 * (uint32_t)kCBlockCols).value;
                if (two) nr = (pos_t)compact_rank(compact_match(s1, c), base1, p1 - cb1 * (uint32_t)kCBlockCols).value;
                else nr = (pos_t)compact_rank(m0, base0, p1 - cb0 * (uint32_t)kCBlockCols).value;
                classic = false;
            }
        }
        if (classic) {
            const uint64_t pc = LAY == LAY_CLASSIC ? pol : pol_cold;
            const BlockPos b0 = split_pos<WIDE>((int64_t)l), b1 = split_pos<WIDE>((int64_t)r + 1);
            two = b1.blk != b0.blk;
            const Sector s0 = ld_sector(sector_ptr<WIDE>(sec, b0.blk, c), pc);
            Sector s1;
            if (two) s1 = ld_sector(sector_ptr<WIDE>(sec, b1.blk, c), pc);
            const SectorPrefix pf0 = sector_prefix(s0);
            nl = (pos_t)sector_rank_fast(s0, pf0, b0.off);
            if (two) {
                const SectorPrefix pf1 = sector_prefix(s1);
                nr = (pos_t)sector_rank_fast(s1, pf1, b1.off);
            } else {
                nr = (pos_t)sector_rank_fast(s0, pf0, b1.off);
            }
            if (WIDE) {
                nl += (pos_t)__ldg(sbbase + (int64_t)c * n_sb + (b0.blk >> sb_shift));
                nr += (pos_t)__ldg(sbbase + (int64_t)c * n_sb + (b1.blk >> sb_shift));
            }
        }
        nr -= 1;
        return two ? 2u : 1u;
    }
};

// the block's shared memory: per warp its queues and the stage its streaming chain collects results in
template <bool STREAMING, bool WIDE, bool OUT32, bool LITERAL>
struct WalkShared {
    static constexpr int NCH = ChainCount<STREAMING, WIDE, LITERAL>::value;
    typedef typename std::conditional<LITERAL, OutStage<OUT32>,
                                      typename std::conditional<STREAMING, ChainStage<WIDE, NCH>, NoStage>::type>::type StageT;
    WalkQueues<WIDE, NCH> queues[kWalkWarps];
    StageT stage[kWalkWarps];
};

// LITERAL (streaming mode on an index that violates "only suffix-group starts carry edges", i.e. a
// hand-made file): streaming answers may then differ from search() answers, so the reference's
// control flow is followed to the letter -- after a miss the k-mers are searched one at a time and
// streaming resumes from the first one found (SBWT.hh:556-576); invalid bases are met by the chain.
// LAY: the layout rank steps are answered from (device_index.cuh); the compact ones serve narrow, non-LITERAL kernels,
// and a block flagged there (some column with no edge or several) is answered from the classic sectors.
template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32, int KW, bool LITERAL, int LAY>
// (the COUNT instantiations -- one counted launch per bench run, never on the product path -- get 128 registers: no spills)
__global__ void __launch_bounds__(kWalkThreads, COUNT ? 2 : WalkBlocks<STREAMING, WIDE, LAY>::value) walk_kernel(const WalkParams P) {
    static_assert(STREAMING || !LITERAL, "LITERAL is a streaming-mode variant");
    static_assert(LAY == LAY_CLASSIC || (!WIDE && !LITERAL), "the compact layouts serve narrow indexes that keep the edge invariant");
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t;
    constexpr int NCH = ChainCount<STREAMING, WIDE, LITERAL>::value;
    typedef WalkQueues<WIDE, NCH> QT;
    typedef WalkShared<STREAMING, WIDE, OUT32, LITERAL> SH;
    extern __shared__ __align__(16) unsigned char walk_smem[];
    SH& shm = *reinterpret_cast<SH*>(walk_smem);
    QT& Q = shm.queues[threadIdx.x >> 5];
    void* const stage_mem = &shm.stage[threadIdx.x >> 5];
    constexpr uint32_t OR = OUT32 ? 8u : 4u; // results per 32-byte output sector
    // phase of result 0 inside its sector
    const uint32_t oph = OUT32 ? (uint32_t)(((uintptr_t)P.out32 >> 2) & 7u) : (uint32_t)(((uintptr_t)P.out >> 3) & 3u);
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t n_items = (uint32_t)*P.n_items;
    const uint32_t k = (uint32_t)ix.k, p = (uint32_t)ix.tp; // p: characters answered by the search table
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    Stepper<WIDE, LAY> ST;
    ST.sec = ix.sectors; ST.cmp = ix.compact; ST.cbase = ix.cbase; ST.sbbase = ix.sbbase; ST.n_sb = ix.n_sb; ST.sb_shift = ix.sb_shift;
    ST.pol = make_l2_policy(true);       // the index stays in L2 while reads and results stream through it
    ST.pol_cold = make_l2_policy(false);
    uint64_t pol_t = ST.pol;             // table rows
    if (P.table_streams) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_t));
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
        // total only grows when input is taken, and input is taken only while it is <= 32 NCH: none of
        // those queues ever holds more than 32 + 32 NCH entries. T receives at most 32 NCH entries from a
        // CHAIN (run only while T holds <= 32) and at most 32 from a NARROW (run only while T is empty).
        enum : int { T_CHAIN_V, T_CHAIN_S, T_NARROW_TODO, T_NARROW_INPUT, T_PROBE, T_EXIT };
        int task;
        bool take_input = false;
        if (nV >= 32) task = T_CHAIN_V;
        else if (STREAMING && nS >= 32u * NCH && nT <= 32) task = T_CHAIN_S;
        else if (nT > 0) task = T_NARROW_TODO;
        else if (STREAMING && nP >= 16) task = T_PROBE;
        else if (STREAMING && nF >= 32) task = T_NARROW_INPUT;
        else if (!input_done && (!STREAMING || nS + nP + nF <= 32u * NCH)) {
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

        if (STREAMING && !LITERAL && task == T_CHAIN_S) {
            // ================================================================ CHAIN, streaming (SBWT.hh:561-575)
            // Every lane owns up to NCH survivors. A chain first finishes its own k-mer (`quiet` steps without a result,
            // then the step that completes it), then every further step answers the next k-mer of its work item.
            // Results are produced in ROUNDS of kRound steps: a chain's result number x goes to slot (x + phase) mod kRound
            // of its row of the warp's stage in shared memory, and after the round the warp writes all rows out together,
            // every row as one contiguous piece of its read's results -- for int64 results a whole 128-byte line per chain
            // (the round-1 kernel wrote 32-byte sectors one at a time: four partial-line writes per line, each chain
            // holding a half-written line in L2; the result stream cost 40 % of the kernel, profiles/r02e).
            typedef ChainStage<WIDE, NCH> CS;
            constexpr uint32_t R = CS::kRound;
            CS& STG = *reinterpret_cast<CS*>(stage_mem);
            const uint32_t ophR = OUT32 ? (uint32_t)(((uintptr_t)P.out32 >> 2) & (R - 1u)) : (uint32_t)(((uintptr_t)P.out >> 3) & (R - 1u));
            const uint32_t m = min(32u * NCH, nS), q0 = nS - m;
            nS = q0;
            bool act[NCH];
            pos_t col[NCH];
            uint32_t pos[NCH], cw[NCH], nx[NCH], o[NCH], oend[NCH], quiet[NCH];
            bool fs[NCH]; // the chain is past its first k-mer: its steps are streaming steps
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                const uint32_t qi = (uint32_t)i * 32u + (uint32_t)lane;
                act[i] = qi < m;
                col[i] = 0; pos[i] = 0; cw[i] = 0; nx[i] = 0; o[i] = 0; oend[i] = 0; quiet[i] = 0; fs[i] = false;
                if (act[i]) {
                    const uint32_t b = Q.s_base[q0 + qi], meta = Q.s_meta[q0 + qi];
                    o[i] = Q.s_out[q0 + qi];
                    col[i] = Q.s_col[q0 + qi];
                    const uint32_t j = meta & 0xFFu;
                    oend[i] = o[i] + (meta >> 8);
                    pos[i] = b + j; // next base to consume
                    cw[i] = __ldg(P.codes + (pos[i] >> 4));
                    nx[i] = __ldg(P.codes + (pos[i] >> 4) + 1);
                    if (j == k) { // NARROW already completed the first k-mer: its column is the answer (SBWT.hh:410-413)
                        store_result<OUT32>(P, o[i], (int64_t)col[i]);
                        if (COUNT) { st_lookups++; st_hits++; }
                        o[i]++;
                        fs[i] = true;
                        if (o[i] == oend[i]) act[i] = false;
                    } else {
                        quiet[i] = k - j - 1u;
                    }
                }
            }
            __syncwarp();

            // the answer of one step from the classic sectors; a clear bit on a streaming step takes the literal walk-back
            // (SBWT.hh:562-563): the step starts from the suffix-group start of col (which alone carries the group's edges)
            auto resolve = [&](pos_t colv, int c, bool streaming_step, pos_t& ncol, bool& hit) {
                int64_t cblk;
                ST.classic_step(colv, c, ncol, hit, cblk);
                if (!hit && streaming_step) {
                    int64_t g = (int64_t)colv;
                    while (true) {
                        const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
                        if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
                        g = (g & ~31ll) - 1;
                    }
                    if (COUNT) st_sectors++;
                    if (g != (int64_t)colv) {
                        int64_t gblk;
                        ST.classic_step((pos_t)g, c, ncol, hit, gblk);
                        if (COUNT) st_sectors += gblk != cblk;
                    }
                }
            };
            // the rest of a chain's item after a miss is searched from scratch (SBWT.hh:557-559): probed (D) or one lane each
            auto push_rest = [&](bool want, uint32_t kstart, uint32_t ov, uint32_t oendv) { // (warp-uniform call)
                const unsigned pm = __ballot_sync(FULL, want);
                if (pm) { // (a chain only covers k-mers free of invalid bases, so the range may be probed)
                    const uint32_t at = (D ? nP : nT) + __popc(pm & lt_mask);
                    if (want) {
                        if (D) { Q.p_base[at] = kstart + 1u; Q.p_out[at] = ov; Q.p_cnt[at] = oendv - ov; }
                        else { Q.t_base[at] = kstart + 1u; Q.t_out[at] = ov; Q.t_cnt[at] = oendv - ov; }
                    }
                    if (D) nP += __popc(pm); else nT += __popc(pm);
                }
            };
            // The next code word, every 16 bases, is loaded IN PLACE under a predicate, and by the step BEFORE the one that
            // moves to a new word, right behind that step's sector load: the two loads are in flight together and the step's
            // one wait covers both. (Issued after the step -- as round 1 and the first round-2 kernels did -- the load is
            // still in flight at the top of the next step, and ptxas, which tracks it on the scoreboard of the sector load,
            // waits for it BEFORE issuing the next sector load: 11 % of the kernel's stall samples, profiles/r02t.)
            auto refill = [&](int i) { // (after the step's character has been taken from cw)
                const uint32_t np = pos[i] + 1u, ph = np & 15u;
                if (ph == 0) cw[i] = nx[i];
                asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
                             : "+r"(nx[i]) : "l"(P.codes + (np >> 4) + 1), "r"(ph));
            };
            // a chain's row of the stage as a shared-space address held in a register (left to itself the compiler rebuilds
            // the generic address from %tid and the shared window in every step: 11 of the step's 75 instructions)
            uint32_t row_sa[NCH];
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                row_sa[i] = (uint32_t)__cvta_generic_to_shared(&STG.v[i][lane][0]);
                asm volatile("mov.u32 %0, %0;" : "+r"(row_sa[i]));
            }
            auto stage_put = [&](int i, uint32_t sl, pos_t v) {
                if (WIDE) asm volatile("st.shared.b64 [%0], %1;" ::"r"(row_sa[i] + sl * 8u), "l"((int64_t)v) : "memory");
                else asm volatile("st.shared.b32 [%0], %1;" ::"r"(row_sa[i] + sl * 4u), "r"((uint32_t)v) : "memory");
            };
            auto advance = [&](int i) { pos[i]++; }; // one base consumed (a step that fails ends its chain: cw may be one ahead)

            while (true) {
                // ---- own k-mers: steps without a result. A miss here is the item's first k-mer being absent.
                while (true) {
                    bool run[NCH], any = false;
#pragma unroll
                    for (int i = 0; i < NCH; i++) { run[i] = act[i] && quiet[i] > 0; any |= run[i]; }
                    if (!__any_sync(FULL, any)) break;
                    typename Stepper<WIDE, LAY>::Load L[NCH];
                    int c[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        c[i] = (int)((cw[i] >> ((pos[i] & 15u) * 2u)) & 3u);
                        if (run[i]) { ST.issue(L[i], col[i], c[i]); refill(i); }
                    }
                    uint32_t wantm = 0; // bit i: chain i ended in a miss and leaves k-mers behind
                    uint32_t kstart[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        kstart[i] = 0;
                        if (run[i]) {
                            pos_t ncol = 0;
                            bool hit = false;
                            if (!ST.eval(L[i], col[i], c[i], ncol, hit)) resolve(col[i], c[i], false, ncol, hit);
                            if (COUNT) { st_ranks += 2; st_sectors += 1; }
                            if (hit) {
                                col[i] = ncol;
                                advance(i);
                                quiet[i]--;
                            } else { // [col, col] -> empty interval (SBWT.hh:433)
                                kstart[i] = pos[i] - (k - 1u - quiet[i]);
                                store_result<OUT32>(P, o[i], -1);
                                if (COUNT) st_lookups++;
                                o[i]++;
                                act[i] = false;
                                if (o[i] < oend[i]) wantm |= 1u << i;
                            }
                        }
                    }
                    if (__any_sync(FULL, wantm != 0)) {
#pragma unroll
                        for (int i = 0; i < NCH; i++) push_rest((wantm >> i) & 1u, kstart[i], o[i], oend[i]);
                        __syncwarp();
                    }
                }

                // ---- one round: chain i produces the results of slots [lo, lo + width) of its row
                uint32_t lo[NCH], width[NCH];
                uint32_t s_first = R, s_last = 0;
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    lo[i] = 0; width[i] = 0;
                    if (act[i]) {
                        lo[i] = (o[i] + ophR) & (R - 1u);
                        width[i] = min(R - lo[i], oend[i] - o[i]);
                        s_first = min(s_first, lo[i]);
                        s_last = max(s_last, lo[i] + width[i]);
                    }
                }
                s_first = __reduce_min_sync(FULL, s_first);
                s_last = __reduce_max_sync(FULL, s_last);
                if (s_last == 0) break; // no chain left
                uint32_t endm = 0; // bit i: chain i ended in a miss during this round
                for (uint32_t sl = s_first; sl < s_last; sl++) {
                    bool run[NCH];
                    typename Stepper<WIDE, LAY>::Load L[NCH];
                    int c[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        run[i] = (sl - lo[i]) < width[i]; // (unsigned: also false below lo; width is 0 for a chain that is not running)
                        c[i] = (int)((cw[i] >> ((pos[i] & 15u) * 2u)) & 3u);
                        if (run[i]) { ST.issue(L[i], col[i], c[i]); refill(i); }
                    }
                    pos_t ncol[NCH];
                    bool ok[NCH], okall = true;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        ncol[i] = 0; ok[i] = true;
                        if (run[i]) {
                            bool hit = false;
                            const bool fast = ST.eval(L[i], col[i], c[i], ncol[i], hit);
                            ok[i] = fast && hit;
                            okall &= ok[i];
                        }
                    }
                    if (__all_sync(FULL, okall)) { // the usual case: every running chain found its next k-mer
#pragma unroll
                        for (int i = 0; i < NCH; i++) {
                            if (run[i]) {
                                stage_put(i, sl, ncol[i]);
                                col[i] = ncol[i];
                                advance(i);
                            }
                        }
                    } else {
                        uint32_t wantm = 0;
                        uint32_t kstart[NCH];
#pragma unroll
                        for (int i = 0; i < NCH; i++) {
                            kstart[i] = 0;
                            if (run[i]) {
                                bool hit = ok[i];
                                // (a chain's first result may be its own k-mer's: that step is not a streaming step)
                                if (!hit) resolve(col[i], c[i], fs[i] || sl > lo[i], ncol[i], hit);
                                if (hit) {
                                    stage_put(i, sl, ncol[i]);
                                    col[i] = ncol[i];
                                    advance(i);
                                } else { // the k-mer is absent: [col, col] -> empty (SBWT.hh:433) / l != r (SBWT.hh:574); the chain ends
                                    stage_put(i, sl, (pos_t)-1);
                                    kstart[i] = pos[i] - (k - 1u);
                                    width[i] = sl + 1u - lo[i];
                                    endm |= 1u << i;
                                    if (o[i] + width[i] < oend[i]) wantm |= 1u << i;
                                    if (COUNT) st_hits--; // (counted below with the round)
                                }
                            }
                        }
                        if (__any_sync(FULL, wantm != 0)) {
#pragma unroll
                            for (int i = 0; i < NCH; i++) push_rest((wantm >> i) & 1u, kstart[i], o[i] + width[i], oend[i]);
                            __syncwarp();
                        }
                    }
                }
                // ---- the round's results leave the stage: every row as one contiguous piece
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    STG.obase[i][lane] = o[i] - lo[i];
                    STG.mask[i][lane] = width[i] ? (((width[i] >= 32u ? 0u : (1u << width[i])) - 1u) << lo[i]) : 0u;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NCH; i++) CS::template flush<OUT32>(STG, i, lane, P.out, P.out32);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    if (act[i]) {
                        if (COUNT) { st_ranks += 2ull * width[i]; st_sectors += width[i]; st_lookups += width[i]; st_hits += width[i]; }
                        if (width[i]) fs[i] = true;
                        o[i] += width[i];
                        if (((endm >> i) & 1u) || o[i] == oend[i]) act[i] = false;
                    }
                }
            }
            __syncwarp();
            continue;
        }

        if (task == T_CHAIN_V || task == T_CHAIN_S) {
            // ================================================================ CHAIN: survivors' own k-mers; LITERAL streaming
            const bool sb = STREAMING && LITERAL && task == T_CHAIN_S;
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
                if (!LITERAL || !sb) { // one result per lane: nothing to collect
                    store_result<OUT32>(P, x, v);
                    gs = x + 1u;
                    return;
                }
                if constexpr (LITERAL) {
                    OutStage<OUT32>& OS = *reinterpret_cast<OutStage<OUT32>*>(stage_mem);
                    const uint32_t sl = (x + oph) & (OR - 1u);
                    OS.v[sl][lane] = (typename OutStage<OUT32>::val_t)v;
                    if (sl == OR - 1u) {
                        if (x - gs == OR - 1u) {
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
                        } else {
                            for (uint32_t y = gs; y <= x; y++) store_result<OUT32>(P, y, (int64_t)OS.v[(y + oph) & (OR - 1u)][lane]);
                        }
                        gs = x + 1u;
                    }
                }
            };
            auto drain = [&](uint32_t end) { // the lane's run is over: write what is still staged, [gs, end)
                if constexpr (LITERAL) {
                    OutStage<OUT32>& OS = *reinterpret_cast<OutStage<OUT32>*>(stage_mem);
                    for (uint32_t y = gs; y < end; y++) store_result<OUT32>(P, y, (int64_t)OS.v[(y + oph) & (OR - 1u)][lane]);
                }
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
                bool miss = false;
                if (act) {
                    const int c = (int)((cw >> ((pos & 15u) * 2u)) & 3u);
                    pos_t ncol = 0;
                    int64_t cblk = -1; // classic block of col, when it was read
                    bool hit = false;
                    typename Stepper<WIDE, LAY>::Load L;
                    ST.issue(L, col, c);
                    { // the next code word, in place and behind the sector load (see the streaming CHAIN: a plain conditional
                      // load after the step cost 11 % of the search-mode kernel's stall samples, profiles/r03m)
                        const uint32_t np = pos + 1u, ph = np & 15u;
                        if (ph == 0) cw = nx;
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
                                     : "+r"(nx) : "l"(P.codes + (np >> 4) + 1), "r"(ph));
                    }
                    if (!ST.eval(L, col, c, ncol, hit)) ST.classic_step(col, c, ncol, hit, cblk);
                    else if (LAY == LAY_CLASSIC) cblk = split_pos<WIDE>((int64_t)col).blk;
                    miss = !hit; // [col, col] -> empty interval (SBWT.hh:433) / l != r (SBWT.hh:574)
                    if (COUNT) { st_ranks += 2; st_sectors += 1; }
                    bool bad = false; // LITERAL: the chain can run into a base outside ACGT (SBWT.hh:565-568)
                    if (LITERAL && fs) bad = ((__ldg(P.invalid + (pos >> 5)) >> (pos & 31u)) & 1u) != 0;
                    if (STREAMING && LITERAL && fs && !bad) {
                        // literal form (SBWT.hh:562-563): the step starts from the suffix-group start of col
                        int64_t g = (int64_t)col;
                        while (true) {
                            const uint32_t sw = __ldg(ix.sgs + (g >> 5)) & (0xFFFFFFFFu >> (31 - (int)(g & 31)));
                            if (sw) { g = (g & ~31ll) + (31 - __clz(sw)); break; }
                            g = (g & ~31ll) - 1;
                        }
                        if (COUNT) st_sectors++;
                        if (g != (int64_t)col) {
                            int64_t gblk;
                            ST.classic_step((pos_t)g, c, ncol, hit, gblk);
                            if (COUNT) st_sectors += gblk != cblk;
                            miss = !hit;
                        }
                    }
                    if (bad) miss = true;
                    if (!miss) {
                        col = ncol;
                        j++;
                        pos++; // (a step that fails ends the lane's run: cw may be one word ahead)
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
                pr_np = min((pr_tc + D - 1) / max(D, 1u), 32u);
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
                const uint32_t ns = ST.narrow(l, r, c, nl, nr);
                if (COUNT) { st_ranks += 2; st_sectors += ns; }
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
