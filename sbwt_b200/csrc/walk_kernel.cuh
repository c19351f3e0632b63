// walk_kernel.cuh -- the colex-interval walk on the device (sm_100a).
//
// Replaces SBWT::search (include/sbwt/SBWT.hh:390-415), SBWT::update_sbwt_interval
// (SBWT.hh:423-437) and SBWT::streaming_search (SBWT.hh:545-581) together with the per-read
// loops of src/CLI/sbwt_search.cpp:45-91.
//
// Work item = up to `window` consecutive k-mers of one read (a whole read when it is short).
// One LANE owns one item at a time and is a small state machine. Every trip of the warp loop
// has exactly ONE point where the warp waits for memory: the (one or two) index sectors of the
// lane's interval step. Everything else a lane needs is already on chip when it is used:
//
//   * the packed read lives in a per-lane ring in shared memory (256 bases of 2-bit codes and
//     invalid flags), filled by cp.async one 64-base chunk ahead of the walk;
//   * the row of the search table (kmer_prefix_precalc, SBWT.hh:404) for the NEXT k-mer is
//     requested while the current from-scratch k-mer is being walked, so a run of absent
//     k-mers (the expensive case: every one of them is a fresh search) never waits for it.
//
//   STEP   one interval step  [l,r] -> [C[c]+rank_c(l), C[c]+rank_c(r+1)-1]  (two sectors, one
//          when l and r+1 fall into the same 224-column block). A streaming step (previous k-mer
//          found at column col, SBWT.hh:561-575) is the same step on [col, col] with the new
//          character: in a reference-built index only suffix-group starts carry edges
//          (SURVEY.md section 8(a) note 7), so a set bit_c(col) already proves col is the group
//          start; a clear bit (or an index violating the invariant) takes the literal
//          walk-back over suffix_group_starts on a slow path.
//   ADVANCE  runs when a lane finishes a k-mer: one result is stored, the lane moves to the next
//          k-mer of its item and either parks in STEP again or answers on the spot the k-mers
//          that need no walk (a non-ACGT byte inside, SBWT.hh:399,428,568; first p characters
//          absent from the table, SBWT.hh:424; p == k).
//
// Finished lanes refill from the warp's own contiguous item range (ballot/popc).
#pragma once

#include <type_traits>

#include "device_index.cuh"

namespace sbwt_b200 {

struct WalkParams {
    DeviceIndexView ix;
    const uint32_t* codes;    // 2-bit bases, 16 per u32 word (64 bases = one 16-byte chunk)
    const uint32_t* invalid;  // 1 bit per base, 32 per u32 word
    uint32_t n_chunks;        // 64-base chunks the two arrays hold (allocation, not batch size)
    const WalkItem* items;
    const int64_t* n_items;   // device scalar
    int64_t* out;             // results as int64 ...
    int32_t* out32;           // ... or as int32 (OUT32 kernels; callers use them only when n_nodes < 2^31)
    unsigned long long* stats; // [lookups, hits, rank_ops, sectors] (COUNT only)
    unsigned long long* cursor; // next unclaimed work item (walk2_kernel; zeroed before every launch)
    int index_evict_last;      // L2 policy of the index / table loads (1 = evict_last; 2, 3: experiments with l2_frac)
    float l2_frac;
    int debug_no_store;        // measurement only (SBWT_B200_DEBUG_NOSTORE): results are not written
    uint32_t probe_stride;     // walk2_kernel, streaming mode: distance in k-mers between the probes of a range of
                               // presumed misses (0 = every k-mer after a miss is searched on its own)
};

constexpr int kRingCodeWords = 16; // u32 words of codes per lane: 4 chunks of 64 bases
constexpr int kRingFlagWords = 8;  // u32 words of invalid flags per lane
constexpr int kWalkThreads = 256;

__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// one row {l, r} of the search table; absent rows are all ones
template <bool WIDE>
struct TableRow;
template <>
struct TableRow<false> {
    uint32_t l, r;
    __device__ __forceinline__ bool absent() const { return l == 0xFFFFFFFFu; }
    static __device__ __forceinline__ TableRow load(const void* table, uint32_t idx, uint64_t pol) {
        TableRow t;
        asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
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
        asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
                     : "=l"(t.l), "=l"(t.r)
                     : "l"(reinterpret_cast<const longlong2*>(table) + idx), "l"(pol));
        return t;
    }
};

template <bool STREAMING, bool WIDE, bool COUNT, bool OUT32>
__global__ void __launch_bounds__(kWalkThreads, WIDE ? 4 : 5) walk_kernel(const WalkParams P) {
    typedef typename std::conditional<WIDE, int64_t, uint32_t>::type pos_t; // columns fit 32 bits in narrow mode
    __shared__ uint32_t ring[(kRingCodeWords + kRingFlagWords) * kWalkThreads];
    const DeviceIndexView& ix = P.ix;
    const unsigned FULL = 0xFFFFFFFFu;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t gw = (blockIdx.x * kWalkThreads + tid) >> 5, nwarps = (gridDim.x * kWalkThreads) >> 5;
    const uint32_t n_items = (uint32_t)*P.n_items;
    // contiguous item range of this warp
    uint32_t next = (uint32_t)(((uint64_t)n_items * gw) / nwarps);
    const uint32_t end = (uint32_t)(((uint64_t)n_items * (gw + 1)) / nwarps);

    const int k = ix.k, p = ix.tp; // p: characters answered by the search table
    const uint32_t pmask = p ? (uint32_t)((1ull << (2 * p)) - 1ull) : 0u;
    const Sector* const sec_base = ix.sectors;
    const uint64_t pol = make_l2_policy(P.index_evict_last != 0);
    const pos_t last_col = (pos_t)(ix.n_nodes - 1);
    // per-lane ring: code word W (absolute index) at ring[(W & 15) * 256 + tid], flag word V at ring[(16 + (V & 7)) * 256 + tid]
    uint32_t ring_t = (uint32_t)__cvta_generic_to_shared(ring) + (uint32_t)tid * 4u;
    asm volatile("" : "+r"(ring_t)); // keep it in a register (the compiler otherwise rebuilds it from %tid every trip)
    constexpr uint32_t kFlagOff = kRingCodeWords * kWalkThreads * 4;

    auto load_chunk = [&](uint32_t ch) { // 64 bases: 4 code words + 2 flag words, asynchronously
        if (ch < P.n_chunks) {
            const uint32_t* gc = P.codes + (size_t)ch * 4;
            const uint32_t* gv = P.invalid + (size_t)ch * 2;
            const uint32_t dc = ring_t + ((ch & 3u) << 12);
            cp_async_4(dc, gc);
            cp_async_4(dc + 1024, gc + 1);
            cp_async_4(dc + 2048, gc + 2);
            cp_async_4(dc + 3072, gc + 3);
            const uint32_t dv = ring_t + kFlagOff + ((ch & 3u) << 11);
            cp_async_4(dv, gv);
            cp_async_4(dv + 1024, gv + 1);
        }
    };

    uint32_t cur = 0;        // first base of the current k-mer (index in the packed batch)
    uint32_t oidx = 0;       // next result slot
    uint32_t remaining = 0;  // k-mers left in the item, current one included; 0 = the lane needs an item
    uint32_t vfrom = 0;      // k-mers starting before this base cover an invalid base
    pos_t l = 0, r = 0;      // current interval
    uint32_t j = 0;          // characters of the current k-mer already consumed
    uint32_t fs = 0;         // this STEP is a streaming step (previous k-mer was found at column l)
    uint32_t pf_valid = 0;   // trow holds the table row of k-mer cur + 1
    // what ADVANCE has to do for this lane: nothing (the lane is parked in a STEP or idle), store -1,
    // store l (the k-mer was found at column l), or set up the first k-mer of a new item
    enum : uint32_t { A_NONE = 0, A_MISS = 1, A_HIT = 2, A_FRESH = 3 };
    uint32_t pend = A_NONE;
    TableRow<WIDE> trow;
    trow.l = 0; trow.r = 0;
    unsigned long long st_lookups = 0, st_hits = 0, st_ranks = 0, st_sectors = 0;

    while (true) {
        // ---- refill finished lanes from the warp's range
        const unsigned need = __ballot_sync(FULL, remaining == 0);
        if (need) {
            if (next >= end) {
                if (need == FULL) break;
            } else {
                const uint32_t mine = next + __popc(need & lt_mask);
                next += __popc(need);
                if (remaining == 0 && mine < end) {
                    const uint4 it = __ldg(reinterpret_cast<const uint4*>(P.items) + mine);
                    cur = it.x; oidx = it.y; remaining = it.z; vfrom = it.w;
                    const uint32_t ch = cur >> 6;
                    cp_async_wait_all(); // the previous item's read-ahead may still be landing in the ring
                    load_chunk(ch);
                    load_chunk(ch + 1);
                    load_chunk(ch + 2);
                    cp_async_wait_all();
                    pend = A_FRESH; pf_valid = 0;
                }
            }
        }

        // ---- STEP: the same instructions for every lane parked in a step
        const bool step = remaining != 0 && pend == A_NONE;
        const uint32_t q = cur + j;
        const int c = (int)((lds_u32(ring_t + ((q << 6) & 0x3C00u)) >> ((q & 15u) * 2u)) & 3u);
        const BlockPos b0 = split_pos<WIDE>((int64_t)l), b1 = split_pos<WIDE>((int64_t)r + 1);
        const bool two = step && (b1.blk != b0.blk);
        Sector s0, s1;
        if (step) s0 = ld_sector(sector_ptr<WIDE>(sec_base, b0.blk, c), pol);
        if (two) s1 = ld_sector(sector_ptr<WIDE>(sec_base, b1.blk, c), pol);
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
                const Sector ss = ld_sector(sector_ptr<WIDE>(sec_base, bs.blk, c), pol);
                if (COUNT) st_sectors += bs.blk != b0.blk;
                miss = sector_bit(ss, bs.off) == 0;
                nl = (pos_t)lf_value<WIDE>(ix, ss, bs.blk, bs.off, c);
                nr = nl;
            }
        }
        // the lane's state after the step, as selects (no control flow: every lane runs the same code)
        if (step) {
            l = nl;
            r = nr;
            j++;
            fs = 0;
            // a k-mer interval is a singleton (SBWT.hh:410-413)
            pend = miss ? (uint32_t)A_MISS : (j == (uint32_t)k ? (uint32_t)A_HIT : (uint32_t)A_NONE);
        }
        __syncwarp();

        // ---- ADVANCE, straight-line and predicated: store one result and set up the lane's next k-mer.
        // A k-mer that needs no walk (a non-ACGT byte inside, its first p characters absent from the
        // table, p == k) leaves `pend` set, and the lane comes back here on the next trip without a STEP.
        {
            const bool adv = pend != A_NONE;
            const bool hit = pend == A_HIT;
            const bool moved = adv && pend != A_FRESH; // a result to store, one k-mer forward
            const bool stored = moved && !P.debug_no_store;
            // one predicated streaming store (written as PTX: the compiler otherwise builds a jump table on `pend`)
            if (OUT32) {
                asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.cs.s32 [%1], %2; }" ::"r"((uint32_t)stored),
                             "l"(P.out32 + oidx), "r"(hit ? (int32_t)l : -1)
                             : "memory");
            } else {
                asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.cs.s64 [%1], %2; }" ::"r"((uint32_t)stored),
                             "l"(P.out + oidx), "l"(hit ? (int64_t)l : (int64_t)-1)
                             : "memory");
            }
            if (COUNT && moved) { st_lookups++; st_hits += hit; }
            const uint32_t inc = moved ? 1u : 0u;
            oidx += inc;
            cur += inc;
            remaining -= inc;
            if (moved && (cur & 63u) == 0) { // entering a new chunk: the one after it was requested 64 k-mers ago
                cp_async_wait_all();
                load_chunk((cur >> 6) + 2);
            }
            const bool su = adv && remaining != 0; // a k-mer to set up
            // the only base of this k-mer not seen by its predecessor
            const uint32_t qn = cur + (uint32_t)k - 1u;
            const uint32_t flag = (lds_u32(ring_t + kFlagOff + ((qn << 5) & 0x1C00u)) >> (qn & 31u)) & 1u;
            if (su && flag) vfrom = qn + 1u;
            const bool kvalid = cur >= vfrom;
            const bool stream = STREAMING && su && kvalid && hit;
            const bool scratch = su && kvalid && !stream;
            TableRow<WIDE> row;
            row.l = 0;
            row.r = last_col;
            bool row_absent = false;
            uint32_t new_pf = 0;
            if (p != 0) {
                // the first 16 bases of the k-mer; first character = least significant digit (SBWT.hh:396-401)
                const uint32_t w0 = lds_u32(ring_t + ((cur << 6) & 0x3C00u));
                const uint32_t w1 = lds_u32(ring_t + (((cur + 16u) << 6) & 0x3C00u));
                const uint32_t E = __funnelshift_r(w0, w1, (cur & 15u) * 2u);
                row = trow;
                if (scratch && !pf_valid) row = TableRow<WIDE>::load(ix.table, E & pmask, pol);
                if (COUNT && scratch) st_sectors++; // one table sector per from-scratch k-mer, however it was fetched
                if (scratch && remaining > 1) { // the next k-mer's row, in flight while this one is walked
                    trow = TableRow<WIDE>::load(ix.table, (E >> 2) & pmask, pol);
                    new_pf = 1;
                }
                row_absent = row.absent();
            }
            if (adv) pf_valid = new_pf;
            if (stream) { r = l; j = (uint32_t)k - 1u; }
            if (scratch) { l = (pos_t)row.l; r = (pos_t)row.r; j = (uint32_t)p; }
            fs = stream ? 1u : 0u; // (a lane that did not advance is mid-walk: fs was cleared by its step)
            pend = (su && !kvalid) || (scratch && row_absent) ? (uint32_t)A_MISS
                 : (scratch && p == k)                       ? (uint32_t)A_HIT // the row is the answer (a singleton, SBWT.hh:410-413)
                                                             : (uint32_t)A_NONE;
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
