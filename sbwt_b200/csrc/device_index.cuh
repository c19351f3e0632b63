// device_index.cuh -- the device-resident layout of the four SBWT bit vectors and the rank primitive.
//
// Replaces sdsl::bit_vector + rank_support_v5<1,1> under SubsetMatrixRank
// (reference: include/sbwt/SubsetMatrixRank.hh:19-37,
//  sdsl-lite/include/sdsl/rank_support_v5.hpp:65-134). The reference keeps, per
// character, a bit array and a separate 2048-bit-superblock directory, so one
// rank touches two arrays. Here one 32-byte SECTOR holds everything one rank
// needs:
//
//   sector(block b, char c) = { u32 count ; 7 x u32 payload }      32 bytes, 32-byte aligned
//     payload = bits [224 b, 224 b + 224) of vector c, LSB-first
//     count   = C[c] + rank_c(224 b)            (narrow: n_nodes < 2^32)
//             = rank_c(224 b) - rank_c(224 * first block of b's superblock)   (wide),
//               with sbbase[c][superblock] = C[c] + rank_c(superblock start) as int64
//   sectors are interleaved [b][c]: address = base + 32 * (4 b + c)
//
// so  C[c] + rank_c(pos) = count + popcount(payload bits below pos % 224)  -- the LF-mapping
// value the interval walk needs (SBWT.hh:430-431) from exactly one sector, fetched with one
// 256-bit load (LDG.E.256 on sm_100a). n_blocks = n_nodes / 224 + 1, so pos == n_nodes is
// addressable (reference: rank(size()) touches the padding word, memory_management.hpp:351-368).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace sbwt_b200 {

constexpr int kBlockCols = 224;       // columns per sector
constexpr int kPayloadWords = 7;      // u32 payload words per sector
constexpr int kDefaultSbShift = 23;   // 2^23 blocks (1.88e9 columns) per superblock in wide mode

struct __align__(32) Sector {
    uint32_t w[8]; // w[0] = count, w[1..7] = payload
};

struct DeviceIndexView {
    const Sector* sectors;     // [n_blocks][4]
    const int64_t* sbbase;     // [4][n_sb], wide mode only
    const int64_t* precalc;    // the file's table: 4^p pairs (l, r) as int64; nullptr if p == 0
    const void* table;         // search table over the first tp characters (see walk_kernel.cuh):
                               //   narrow: 4^tp x {u32 l, u32 r}, 4 rows per sector, absent = {0xFFFFFFFF, 0xFFFFFFFF}
                               //   wide:   4^tp x {i64 l, i64 r}, 2 rows per sector, absent = {-1, -1}
    int tp;                    // 0 = no table
    const uint32_t* sgs;       // suffix_group_starts as u32 words (padded), nullptr if absent
    int64_t n_nodes;
    int64_t n_blocks;
    int64_t n_sb;
    int k;
    int p;
    int sb_shift;
    int wide;                  // 1 if n_nodes >= 2^32 (or forced for tests)
    int edges_at_starts;       // structural invariant (i) of SURVEY.md section 8(a) note 7 holds
    // compact (one-hot) layout, see below; nullptr when the index is not eligible
    const Sector* compact;     // [n_cblocks]: csectors of `layout` (96 or 64 columns each)
    const uint32_t* cbase;     // LAY_C96 only: [n_csb][4]: C[c] + rank_c(first column of the superblock)
    int64_t n_cblocks;
    int layout;                // LAY_CLASSIC / LAY_C96 / LAY_C64: what `compact` holds (LAY_CLASSIC: nothing)
};

constexpr int LAY_CLASSIC = 0; // 224 columns x 1 character per sector
constexpr int LAY_C96 = 1;     // 96 columns x 4 characters, 15-bit relative counts + cbase[]
constexpr int LAY_C64 = 2;     // 64 columns x 4 characters, absolute counts inline

// ------------------------------------------------------------------ compact (one-hot) layout
//
// In an SBWT of low-repeat sequence almost every column carries exactly ONE outgoing edge (config 2/3 of
// BASELINE.json: 99.999 % of the columns; E. coli: 99.4 %), so the four bit vectors are one-hot per column and
// two bits per column say everything. For such indexes a second array is kept next to the sectors above,
//
//   csector(block b) = { 4 x u16 rel_c ; 96 bits lo ; 96 bits hi }                32 bytes, ALL FOUR characters
//     (hi, lo)[j] = the character of the one edge of column 96 b + j  (A=00 C=01 G=10 T=11)
//     rel_c       = rank_c(96 b) - rank_c(first column of b's superblock)  (< 2^15; superblock = 256 blocks)
//     bit 15 of rel_A = FLAG: some column of the block has no edge or more than one -> the block is answered from
//                       the classic sectors instead (rare by the eligibility test at load)
//   C[c] + rank_c(pos) = cbase[superblock][c] + rel_c + #{ j < pos % 96 : (hi, lo)[j] == c }
//
// 0.333 B per column instead of 0.571: a 100 M-column index is 33 MB and stays L2-resident next to the read and
// result streams (the measured capacity for randomly read data on this chip is ~62 MB: profiles/r01g_l2_capacity.txt),
// and a rank is 2 LOP3 + 1 POPC per 32 columns over 3 words instead of 7.
constexpr int kCBlockCols = 96;
constexpr int kCSbShift = 8; // 256 blocks = 24576 columns per superblock

struct CompactRank {
    uint32_t value; // C[c] + rank_c(pos)
    uint32_t bit;   // bit_c(pos): does column pos have an edge labelled c
};

__device__ __forceinline__ bool csector_flagged(const Sector& s) { return (s.w[0] & 0x8000u) != 0; }

// the columns of a csector whose edge is c, 32 per word, and the block's relative count for c
struct CompactMatch {
    uint32_t m0, m1, m2, rel;
};

__device__ __forceinline__ CompactMatch compact_match(const Sector& s, int c) {
    const uint32_t cw = (c & 2) ? s.w[1] : s.w[0];
    const uint32_t X = (c & 1) ? 0u : 0xFFFFFFFFu, Y = (c & 2) ? 0u : 0xFFFFFFFFu;
    CompactMatch m;
    m.rel = ((c & 1) ? (cw >> 16) : cw) & 0x7FFFu;
    m.m0 = (s.w[2] ^ X) & (s.w[5] ^ Y);
    m.m1 = (s.w[3] ^ X) & (s.w[6] ^ Y);
    m.m2 = (s.w[4] ^ X) & (s.w[7] ^ Y);
    return m;
}

// 0xFFFFFFFF << n with PTX semantics (n >= 32 gives 0), so ~shl_clamp(n) is the mask of the n lowest bits for 0 <= n <= 32+
__device__ __forceinline__ uint32_t shl_ones_clamp(uint32_t n) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xFFFFFFFFu), "r"(n));
    return r;
}

// off = pos % 96, base = cbase[(pos / 96) >> kCSbShift][c]. Straight-line: three masked popcounts.
__device__ __forceinline__ CompactRank compact_rank(const CompactMatch& m, uint32_t base, uint32_t off) {
    const uint32_t n1 = (uint32_t)max((int)off - 32, 0), n2 = (uint32_t)max((int)off - 64, 0);
    const uint32_t c0 = __popc(m.m0 & ~shl_ones_clamp(off)), c1 = __popc(m.m1 & ~shl_ones_clamp(n1)), c2 = __popc(m.m2 & ~shl_ones_clamp(n2));
    const uint32_t f = off >> 5;
    const uint32_t mf = f == 0 ? m.m0 : (f == 1 ? m.m1 : m.m2);
    CompactRank r;
    r.value = base + m.rel + c0 + c1 + c2;
    r.bit = (mf >> (off & 31u)) & 1u;
    return r;
}

// ------------------------------------------------------------------ LAY_C64: self-contained one-hot csectors
//
//   csector64(block b) = { 4 x u32 cnt_c ; 64 bits lo ; 64 bits hi }              32 bytes, ALL FOUR characters of 64 columns
//     cnt_c       = C[c] + rank_c(64 b), absolute: one rank is exactly ONE memory request (no cbase[] word, no
//                   division: block = pos >> 6), two masked popcounts
//     (hi, lo)[j] = the character of the one edge of column 64 b + j
//     cnt_A == 0xFFFFFFFF: some column of the block has no edge or several -> answered from the classic sectors
//                   (cnt_A <= C[1] <= n_nodes < 2^32 - 256 otherwise, so the marker cannot be a count)
// 0.5 B per column (LAY_C96: 0.333): which of the two a narrow one-hot index gets is a measured choice of the loader.
constexpr int kC64Cols = 64;

__device__ __forceinline__ bool c64_flagged(const Sector& s) { return s.w[0] == 0xFFFFFFFFu; }

// value = C[c] + rank_c(pos), bit = bit_c(pos) for in-block offset off = pos & 63 (0 <= off < 64)
__device__ __forceinline__ CompactRank c64_rank(const Sector& s, int c, uint32_t off) {
    const bool c0 = (c & 1) != 0, c1 = (c & 2) != 0;
    const uint32_t X = c0 ? 0u : 0xFFFFFFFFu, Y = c1 ? 0u : 0xFFFFFFFFu;
    const uint32_t ca = c0 ? s.w[1] : s.w[0], cb = c0 ? s.w[3] : s.w[2];
    const uint32_t m0 = (s.w[4] ^ X) & (s.w[6] ^ Y), m1 = (s.w[5] ^ X) & (s.w[7] ^ Y);
    const uint32_t n1 = (uint32_t)max((int)off - 32, 0);
    CompactRank r;
    r.value = (c1 ? cb : ca) + __popc(m0 & ~shl_ones_clamp(off)) + __popc(m1 & ~shl_ones_clamp(n1));
    r.bit = (((off & 32u) ? m1 : m0) >> (off & 31u)) & 1u;
    return r;
}

// cbase entry of (csector block cb, character c); volatile so that it is issued where it is written -- ahead of the
// csector load it travels with -- instead of being sunk behind the branch on the csector's flag
__device__ __forceinline__ uint32_t ld_cbase(const uint32_t* __restrict__ cbase, uint32_t cb, int c) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(cbase + ((cb >> kCSbShift) << 2) + c));
    return v;
}

// 256-bit read-only load of one sector, not allocated in L1 (random access, no reuse there).
__device__ __forceinline__ Sector ld_sector(const Sector* p) {
    Sector s;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]),
                   "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p));
    return s;
}

// Same load with an L2 eviction policy (createpolicy): evict_last keeps the index resident in
// L2 while reads and results stream through it.
__device__ __forceinline__ uint64_t make_l2_policy(bool evict_last) {
    uint64_t pol;
    if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ Sector ld_sector(const Sector* p, uint64_t pol) {
    Sector s;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]),
                   "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p), "l"(pol));
    return s;
}

// count + number of payload bits below in-block offset o (0 <= o < 224).
__device__ __forceinline__ uint32_t sector_rank(const Sector& s, uint32_t o) {
    const uint32_t full = o >> 5, rem = o & 31u;
    uint32_t r = s.w[0];
#pragma unroll
    for (int i = 0; i < kPayloadWords; i++) {
        const uint32_t m = ((uint32_t)i < full) ? 0xFFFFFFFFu : (((uint32_t)i == full) ? ((1u << rem) - 1u) : 0u);
        r += __popc(s.w[i + 1] & m);
    }
    return r;
}

// The same quantity with the work moved off the integer-ALU pipe (the walk is ALU-bound, ncu
// r01): the popcounts of payload words 0..5 are packed into bytes and turned into exclusive
// prefix sums with two multiplies (IMAD runs on the otherwise idle FMA pipe); one PRMT then
// picks the prefix of the word that holds the position, and only that word is masked.
//   X = [0, s1, s2, s3]   Y = [s4, s5, s6, *]   with s_f = ones in payload words [0, f)
struct SectorPrefix {
    uint32_t X, Y;
};

__device__ __forceinline__ SectorPrefix sector_prefix(const Sector& s) {
    const uint32_t pc0 = __popc(s.w[1]), pc1 = __popc(s.w[2]), pc2 = __popc(s.w[3]), pc3 = __popc(s.w[4]);
    const uint32_t pc4 = __popc(s.w[5]), pc5 = __popc(s.w[6]);
    const uint32_t P = pc0 + pc1 * 0x100u + pc2 * 0x10000u + pc3 * 0x1000000u;
    const uint32_t Q = P * 0x01010101u; // byte j = pc0 + ... + pcj  (<= 128, no carries)
    SectorPrefix r;
    r.X = Q << 8;
    r.Y = (Q >> 24) * 0x010101u + pc4 * 0x010100u + pc5 * 0x010000u; // [s4, s4+pc4, s4+pc4+pc5]
    return r;
}

__device__ __forceinline__ uint32_t sector_word(const Sector& s, uint32_t f) { // payload word f (0..6)
    const bool b0 = f & 1u, b1 = f & 2u, b2 = f & 4u;
    const uint32_t t0 = b0 ? s.w[2] : s.w[1], t1 = b0 ? s.w[4] : s.w[3], t2 = b0 ? s.w[6] : s.w[5];
    const uint32_t u0 = b1 ? t1 : t0, u1 = b1 ? s.w[7] : t2;
    return b2 ? u1 : u0;
}

__device__ __forceinline__ uint32_t sector_rank_fast(const Sector& s, const SectorPrefix& pf, uint32_t o) {
    const uint32_t f = o >> 5, rem = o & 31u;
    const uint32_t w = sector_word(s, f);
    return s.w[0] + __byte_perm(pf.X, pf.Y, f) + __popc(w & ((1u << rem) - 1u));
}

// payload bit at in-block offset o.
__device__ __forceinline__ uint32_t sector_bit(const Sector& s, uint32_t o) {
    const uint32_t wi = o >> 5;
    uint32_t w = s.w[1];
#pragma unroll
    for (int i = 1; i < kPayloadWords; i++) w = (wi == (uint32_t)i) ? s.w[i + 1] : w;
    return (w >> (o & 31u)) & 1u;
}

// One unit of work of the walk kernel: up to `window` consecutive k-mers of one read.
struct __align__(16) WalkItem {
    uint32_t base;  // index in the packed batch of the first base of the item's first k-mer
    uint32_t out;   // index of its first result
    uint32_t cnt;   // k-mers in the item
    uint32_t vfrom; // one past the last invalid base among [base, base + k - 1), or `base` if there is none
};

struct BlockPos {
    int64_t blk;
    uint32_t off;
};

template <bool WIDE>
__device__ __forceinline__ BlockPos split_pos(int64_t pos) {
    BlockPos bp;
    if (WIDE) {
        bp.blk = (int64_t)((uint64_t)pos / (uint64_t)kBlockCols);
        bp.off = (uint32_t)((uint64_t)pos - (uint64_t)bp.blk * (uint64_t)kBlockCols);
    } else {
        const uint32_t p32 = (uint32_t)pos;
        const uint32_t b = p32 / (uint32_t)kBlockCols;
        bp.blk = b;
        bp.off = p32 - b * (uint32_t)kBlockCols;
    }
    return bp;
}

template <bool WIDE>
__device__ __forceinline__ const Sector* sector_addr(const DeviceIndexView& ix, int64_t blk, int c) {
    return ix.sectors + ((blk << 2) + c);
}

template <bool WIDE>
__device__ __forceinline__ const Sector* sector_ptr(const Sector* base, int64_t blk, int c) {
    if (WIDE) return base + ((blk << 2) + c);
    return base + (uint32_t)(((uint32_t)blk << 2) + (uint32_t)c); // < 2^32 sectors' worth of bytes offset: one IMAD.WIDE.U32
}

// C[c] + rank_c(pos) given the already loaded sector of pos's block.
template <bool WIDE>
__device__ __forceinline__ int64_t lf_value(const DeviceIndexView& ix, const Sector& s, int64_t blk, uint32_t off, int c) {
    int64_t v = (int64_t)sector_rank(s, off);
    if (WIDE) v += __ldg(ix.sbbase + (int64_t)c * ix.n_sb + (blk >> ix.sb_shift));
    return v;
}

} // namespace sbwt_b200
