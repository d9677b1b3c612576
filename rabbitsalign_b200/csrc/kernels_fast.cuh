// kernels_fast.cuh -- packed s16x2 DPX wavefront kernel for sm_100a (the bulk path).
//
// What it computes: the same DP and direction information as the reference's
// gasal_local_kernel<LOCAL, WITH_TB, FALSE> (GASAL2/src/kernels/local_kernel_template.h:45-60,118-430)
// for pairs over the alphabet {A,C,G,T,N} (any case).  How: nothing like the reference.
//
//   * inter-query packing: every 32-bit register holds the same DP quantity of TWO pairs of equal query
//     length (pair A in the low 16 bits, pair B in the high 16 bits), so one DPX instruction
//     (VIMNMX3.S16x2, VIADDMNMX.S16x2, VIMNMX.S16x2) advances two cells;
//   * L lanes ("group"; 4 for |q| <= 160, 8 up to 256, 16 beyond) sweep one such double-pair as an anti-diagonal
//     wavefront: lane l owns a contiguous run of C = ceil(|q|/L) query columns in registers (H of the previous row, E)
//     and is one target row behind lane l-1; H and F of a lane's last column reach the next lane with one shuffle each
//     per row; a warp carries 32/L groups;
//   * scores are kept BIASED (value + kBias in every half) so every half stays a non-negative 16-bit
//     number: plain 32-bit IADD3 then adds/subtracts both halves at once with no cross-half borrow, and a
//     "difference >= 1" predicate becomes one IADD3 with a constant that carries into a chosen bit;
//   * the substitution score is ONE byte-permute per double-cell: per target row the two pairs' 4-entry
//     score profiles sit in two registers (staged in shared memory once per group), and a per-column
//     selector picks the byte of the query base of A and of B;
//   * the four direction facts per cell (F opened, E opened, H != diagonal, max(F,E,0) != F) are gathered
//     with bit-selects into one nibble per cell per pair and streamed to a global scratch tile
//     (fast_layout.cuh); the match/mismatch bit is not stored (the traceback recomputes it from the bases);
//   * the end cell: every cell's key (H << B | 2^B-1 - column in lane), B = 5 or 6 column bits, is one IMAD; per lane and
//     half the kernel keeps the key of the first cell, in the reference's visiting order (8-row block, column, row in
//     block), among the cells with the largest H seen so far (exact, no fallback);
//   * symbols outside {A,C,G,T,N}: the pair is flagged and fully redone by the exact kernel.
#pragma once
#include <mutex>
#include "common.cuh"
#include "fast_cell.cuh"
#include "fast_layout.cuh"
#include "kernels_exact.cuh"
#include "kernels_tb.cuh"

namespace rsa {

constexpr int kFastMinQlen = 8;
constexpr int kFastMaxC = 40;                 // 4-lane groups: C <= 40 (160 bases); 8- and 16-lane groups: C <= 32
constexpr int kFastMaxQlen = 16 * 32;         // 512: 4 lanes x C<=40, 8 lanes x C<=32 up to 256 bases, 16 lanes x C<=32 beyond
constexpr int kFastMaxTlen = 2047;
constexpr int kFastStageUnits = 34;           // 64-base units of the longest window (2047 bases at any phase: 33) + 1
constexpr int kFastStageBytes = 24 * kFastStageUnits + 16;   // per warp: 16 B of codes + 8 B of flags per unit, one status word
constexpr int kFastLutBytes = 64;             // two 8-word profile tables (pair A, pair B) at the head of the shared memory
// 8-lane groups: 4 groups per warp, 4 warps per block; 16-lane groups: 2 groups per warp, 2 warps per block
// (their shared-memory ring is deeper and wider)
__host__ __device__ constexpr int fast_groups_per_warp(int L) { return 32 / L; }
__host__ __device__ constexpr int fast_warps_per_block(int L) { return L == 16 ? 2 : 4; }
__host__ __device__ constexpr int fast_ring_slots(int L) { return L == 4 ? 8 : (L == 8 ? 16 : 32); }  // >= L - 1 + 4, power of two
// (Tracing pairs back inside this kernel was tried in round 1 and dropped: 1.36 vs 1.50 TCUPS per step)

struct FastGroup {
    uint32_t a, b;      // pair indices in the chunk (b == a: lone pair; a == 0xFFFFFFFF: empty slot)
    uint64_t dir_off;   // byte offset of the group's direction tile in the scratch
    uint16_t qlen;
    uint16_t rows;      // max(|t_a|, |t_b|)
};

static_assert(sizeof(FastGroup) == sizeof(FastGroupRef), "FastGroupRef must mirror FastGroup");

// Per-chunk redo bookkeeping living in the metadata blob (host zeroes it before the upload).
struct RedoHeader {
    unsigned int count;             // number of entries in list[]
    unsigned int pad;
    unsigned long long scratch_used; // bytes handed out from the redo scratch region
};

// Can the packed kernel represent this scoring?  (biased halves must stay positive; the signed profile bytes
// score - gap_oe [- 1] must all be negative and fit a byte: fast_cell.cuh)
__host__ inline bool fast_scoring_ok(const Scoring& sc) {
    return sc.match > 0 && sc.mismatch > 0 && sc.gap_ext >= 0 && sc.gap_oe >= sc.gap_ext && sc.gap_oe > sc.match &&
           sc.mismatch + sc.gap_oe < kBias - 2 && sc.match + sc.mismatch < 128 &&
           sc.match * kFastMinQlen < 1024;  // per-pair bound match*|q| <= 1023 is applied by the planner
}

// nibble (ascii & 0xF) -> code: A(1)->0 C(3)->1 G(7)->2 T(4)->3 N(0xE)->4, everything else 0xF
// (a 16 x 4-bit table in one 64-bit constant: shift + mask instead of five compare/select pairs -- the staging of a group
// converts ~200 bytes per lane, and staging was 12 % of the kernel's warp time, profiles/r2_kernels.md)
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t nib) {
    constexpr unsigned long long kTable = 0xF4FFFFFF2FF31F0Full;   // entry n = bits 4n .. 4n+3
    // A(1)->0 C(3)->1 G(7)->2 T(4)->3 N(0xE)->4, every other nibble 0xF
    static_assert(kTable == ((0xFull << 0) | (0x0ull << 4) | (0xFull << 8) | (0x1ull << 12) | (0x3ull << 16) | (0xFull << 20) | (0xFull << 24) |
                             (0x2ull << 28) | (0xFull << 32) | (0xFull << 36) | (0xFull << 40) | (0xFull << 44) | (0xFull << 48) |
                             (0xFull << 52) | (0x4ull << 56) | (0xFull << 60)), "nibble -> base code table");
    return (uint32_t)(kTable >> (4u * (nib & 0xFu))) & 0xFu;
}

// The resident reference in the packed form the staging above reads (north_star: 2-bit-packed windows, vectorised loads):
// plane 1, 2 bits per base: A C G T -> 0..3; N -> 0 and every other symbol -> 1 with plane 2, 1 bit per base ("not ACGT"),
// set.  One thread packs 32 bases (two 128-bit loads of ASCII -> two code words + one flag word).  The ASCII copy stays
// resident beside it: the traceback, the exact kernel (symbols outside ACGTN compare by nibble) and explicit windows use it.
__global__ void __launch_bounds__(256) pack_reference_kernel(const uint8_t* __restrict__ ref, long long len, long long n_words32,
                                                             uint32_t* __restrict__ codes, uint32_t* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words32) return;
    const long long base = i * 32;
    uint32_t bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (base + 32 <= len && (reinterpret_cast<uintptr_t>(ref) & 15u) == 0) {
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(ref + base)), v1 = __ldg(reinterpret_cast<const uint4*>(ref + base) + 1);
        bytes[0] = v0.x; bytes[1] = v0.y; bytes[2] = v0.z; bytes[3] = v0.w; bytes[4] = v1.x; bytes[5] = v1.y; bytes[6] = v1.z; bytes[7] = v1.w;
    } else {
        for (int j = 0; j < 32 && base + j < len; ++j) bytes[j >> 2] |= (uint32_t)ref[base + j] << (8 * (j & 3));
    }
    uint32_t c0 = 0, c1 = 0, f = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint32_t code = base_code(nibble_of((uint8_t)(bytes[j >> 2] >> (8 * (j & 3)))));
        const uint32_t two = code < 4u ? code : (code == 4u ? 0u : 1u);
        if (base + j < len) {
            if (j < 16) c0 |= two << (2 * j); else c1 |= two << (2 * (j - 16));
            f |= (code >= 4u ? 1u : 0u) << j;
        }
    }
    codes[2 * i] = c0;
    codes[2 * i + 1] = c1;
    flags[i] = f;
}

__device__ __forceinline__ int half_s(uint32_t v, int h) { return (int)(int16_t)(h ? (v >> 16) : (v & 0xFFFFu)); }

// Column keys: the running maximum is tracked on key = (H << 5) | (31 - column_in_lane), still one 16-bit half
// per pair (H = unbiased score < 1024, i.e. match * |q| <= 1023: checked per pair by the planner).  The larger key wins, i.e. the larger H and, for equal H in one row,
// the smaller column -- the reference's order inside a row (SURVEY.md 8a rule 3).  Building the key is one
// IMAD (FMA pipe, otherwise idle); no per-column maximum registers are needed.
// 4-lane groups own up to 40 columns per lane: 6 column bits, H < 512 (match * |q| <= 511, |q| <= 160).
__host__ __device__ constexpr int fast_col_bits(int L) { return L == 4 ? 6 : 5; }
__host__ __device__ inline bool fast_key_ok(int qlen, int match) {
    return match * qlen <= (fast_lanes_for(qlen) == 4 ? 511 : 1023);
}

// FLEX = false: every group of the launch has exactly C = ceil(|q|/L) columns per lane (the widest lanes; the others C - 1).
// FLEX = true ("merged classes", variable-length batches): groups with ceil(|q|/L) in C-3 .. C share one launch; a lane
// owns C-4 .. C columns, the last four column slots are conditional and a group's tile keeps its own (narrower) geometry.
constexpr int kFlexSlack = 3;   // a flex launch of width C takes groups of width C - kFlexSlack .. C

template <int L, int C, bool HASN, bool FLEX>
struct FastDp {
    // One group's sweep.  All 32 lanes of the warp call this together (shuffles inside).
    //
    // Per step (= one target row per lane) the work is ordered so that nothing at the head of the step waits:
    //   phase 0  issue the two shuffles (H and F of the left neighbour's last column) and the shared-memory
    //            loads of the NEXT step's target profile;
    //   phase 1  S[c] <- H(r-1, c-1) + sub(r, c) for all columns, in place and right-to-left: independent of
    //            the shuffles, covers their latency;
    //   phase 2  the F-dependent chain left-to-right (H, E', F', direction flags, keys).
    __device__ static __forceinline__ void run(const FastConsts& k, const uint8_t* __restrict__ tcodes,
                                               const uint32_t* __restrict__ lut, int rows, int nsteps, int gl,
                                               int ncols, int nw2g, const uint32_t (&qsel)[C], uint32_t* __restrict__ dir,
                                               uint32_t* __restrict__ ring, int lane, uint32_t& bestkey_out,
                                               int (&firstrow)[2]) {
        uint32_t S[C], E[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { S[c] = k.zero; E[c] = k.zero; }
        uint32_t Hlast = k.zero, Fout = k.zero, Hl_prev = k.zero;
        uint32_t bestkey = 0;  // key of "no positive cell yet"
        firstrow[0] = firstrow[1] = 0;
        constexpr int NW = (C + 3) / 4;
        constexpr int NW2 = (NW + 1) & ~1;      // stored words per lane and row (fast_layout.cuh); FLEX: the group's own, nw2g
        constexpr int D = FLEX ? kFlexSlack + 1 : 1;   // a lane owns C - D .. C columns
        const bool wide = ncols == C;
        if (!FLEX) nw2g = NW2;
        constexpr int RS = fast_ring_slots(L);
        constexpr int CB = fast_col_bits(L);
        constexpr uint32_t kColMask = (uint32_t)((1 << CB) - 1) * 0x00010001u;
        const uint32_t kmul = CB == 6 ? k.k64 : k.k32;
        const int rmax = rows > 0 ? rows - 1 : 0;
        uint2 pr;
        {
            const uint32_t tc = tcodes[0];
            pr.x = lut[tc & 0xFu];
            pr.y = lut[8u + (tc >> 4)];
        }
        for (int s = 0; s < nsteps; ++s) {
            // ---- phase 0
            uint32_t Hl = __shfl_up_sync(0xFFFFFFFFu, Hlast, 1, L);
            uint32_t Fl = __shfl_up_sync(0xFFFFFFFFu, Fout, 1, L);
            const int r = s - gl;
            const uint32_t tcn = tcodes[min(max(r + 1, 0), rmax)];
            uint2 prn;
            prn.x = lut[tcn & 0xFu];
            prn.y = lut[8u + (tcn >> 4)];
            if (r >= 0 && r < rows) {
                // ---- phase 1
#pragma unroll
                for (int c = C - 1; c >= 0; --c) {
                    uint32_t sub = prmt(pr.x, pr.y, qsel[c]);
                    if (HASN) sub = bitsel(prmt(0xFFFFFFFFu, 0u, qsel[c] >> 16), sub, k.sub_n);
                    S[c] = (c == 0 ? Hl_prev : S[c - 1]) + sub;
                }
                // ---- phase 2
                uint32_t F = (gl == 0) ? k.zero : Fl;
                uint32_t rowkey = 0, nib_even = 0, p_lo = 0, Fsave = F, key_prev = 0;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (!FLEX && c == C - 1) Fsave = F;  // F entering the last (conditional) column
                    uint32_t nib = 0;           // (absent column of a narrow lane: zero nibble)
                    if (c < C - D || c < ncols) {
                        uint32_t h, fn, en, key;
                        fast_cell(k, S[c], F, E[c], key_colconst<CB>(c), kmul, h, fn, en, nib, key);
                        if (c & 1) rowkey = __vimax3_s16x2(rowkey, key_prev, key);
                        else if (c == C - 1) rowkey = __vmaxs2(rowkey, key);
                        key_prev = key;
                        S[c] = h;
                        E[c] = en;
                        F = fn;
                        if (FLEX && c >= C - D - 1 && c == ncols - 1) { Hlast = h; Fout = fn; }   // the lane's last real column
                    } else if (c & 1) {
                        rowkey = __vmaxs2(rowkey, key_prev);  // (an even last column has no pending key)
                    }
                    // direction words of this lane and row (fast_cell.cuh: dir_pair / dir_word): parked in this lane's
                    // shared-memory ring as soon as a word is complete, so that it does not occupy a register, until every
                    // lane of the group has reached the same row
                    uint32_t* const slot = ring + ((((s & (RS - 1)) * NW + (c >> 2)) << 5) + lane);
                    if ((c & 3) == 0) { nib_even = nib; if (c == C - 1) *slot = nib; }
                    else if ((c & 3) == 1) { p_lo = dir_pair(k, nib_even, nib); if (c == C - 1) *slot = p_lo; }
                    else if ((c & 3) == 2) { nib_even = nib; if (c == C - 1) *slot = dir_word(p_lo, nib); }
                    else *slot = dir_word(p_lo, dir_pair(k, nib_even, nib));
                }
                if (!FLEX) {
                    Hlast = wide ? S[C - 1] : S[(C >= 2) ? C - 2 : 0];
                    Fout = wide ? F : Fsave;
                }
                // Running maximum of this lane, per half, kept as the FIRST cell in the reference's visiting order
                // (8-row block, column, row in block) among the cells seen so far with the largest H:
                //   row H > best H                      -> this row's key wins;
                //   row H == best H, same 8-row block and smaller column (= larger key) -> this row's key wins;
                //   otherwise the earlier cell stays (earlier block, or same block and smaller-or-equal column).
                bool ge_lo, ge_hi, g2_lo, g2_hi;
                (void)__vibmax_s16x2(bestkey | kColMask, rowkey, &ge_hi, &ge_lo);  // ge: best H >= row H
                (void)__vibmax_s16x2(bestkey, rowkey, &g2_hi, &g2_lo);             // g2: best key >= row key
                const int rb = r >> 3;
                const bool up_lo = !g2_lo && (!ge_lo || rb == (firstrow[0] >> 3));
                const bool up_hi = !g2_hi && (!ge_hi || rb == (firstrow[1] >> 3));
                if (up_lo) firstrow[0] = r;
                if (up_hi) firstrow[1] = r;
                bestkey = bitsel((up_lo ? 0x0000FFFFu : 0u) | (up_hi ? 0xFFFF0000u : 0u), rowkey, bestkey);
            }
            Hl_prev = (gl == 0) ? k.zero : Hl;
            pr = prn;
            // De-skewed, row-blocked store: lane gl computed row r at step r+gl and parked its words in ring slot
            // (r+gl)&(RS-1).  When the group's LAST lane has finished the last row of a 4-row block (step = 4b+3+L-1),
            // every lane stores ITS OWN words of rows 4b..4b+3, two words (4 rows x 8 columns, both pairs) per 256-bit
            // store = one full 32-byte sector; the sectors of row blocks 2u and 2u+1 alternate in memory (64-byte units of
            // 8 rows x 8 columns, fast_layout.cuh).
            const int rl = s - (L - 1);                   // row the last lane finished in this step
            if (rl >= 0 && (rl & 3) == 3 && rl - 3 < rows) {
                const int rb = rl >> 2;
                // sector of words (wv, wv + 1) of this lane and row block: unit ((rb >> 1) * L + gl) * (W / 2) + wv / 2, half rb & 1
                uint4* dblk = reinterpret_cast<uint4*>(dir) + (((size_t)(rb >> 1) * L + gl) * nw2g + (rb & 1)) * 2;
                const uint32_t* r0 = ring + ((((rl - 3 + gl) & (RS - 1)) * NW) << 5) + lane;
                const uint32_t* r1 = ring + ((((rl - 2 + gl) & (RS - 1)) * NW) << 5) + lane;
                const uint32_t* r2 = ring + ((((rl - 1 + gl) & (RS - 1)) * NW) << 5) + lane;
                const uint32_t* r3 = ring + ((((rl + gl) & (RS - 1)) * NW) << 5) + lane;
#pragma unroll
                for (int wv = 0; wv < NW2; wv += 2) {
                    if (FLEX && wv >= nw2g) break;   // a narrower group's tile has fewer words per lane and row
                    uint4 v, w = make_uint4(0u, 0u, 0u, 0u);
                    v.x = r0[wv << 5]; v.y = r1[wv << 5]; v.z = r2[wv << 5]; v.w = r3[wv << 5];
                    if (wv + 1 < NW) { w.x = r0[(wv + 1) << 5]; w.y = r1[(wv + 1) << 5]; w.z = r2[(wv + 1) << 5]; w.w = r3[(wv + 1) << 5]; }
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                 :: "l"(dblk + 2 * wv), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                }
            }
        }
        bestkey_out = bestkey;
    }
};

template <int L, int C, bool FLEX>
// (4-lane groups with C = 28..40 columns would take ~250 registers: two blocks per SM.  Capped at 168 registers for three
// blocks they spill ~30 words per thread and measure slightly faster with the clamp-fact recipe: DP phase 1983 vs 1971
// GCUPS, value 1914 vs 1891.  A/B builds that did not pay, numbers in DESIGN.md 4.1: row loop unrolled by two (1951 /
// 1894); per-column selectors re-loaded from shared memory each row instead of held in C registers (168 registers, no
// spills, DP phase 2037 -- but 74 KB of shared memory per block leave no room for the traceback kernel beside it: 1732).)
#ifndef RSA_FAST_WIDE_BLOCKS
#define RSA_FAST_WIDE_BLOCKS 3   // (A/B builds) 4-lane groups, C > 27
#endif
#ifndef RSA_FAST_WIDE_BLOCKS8
#define RSA_FAST_WIDE_BLOCKS8 2  // (A/B builds) 8-lane groups, C > 27: 255 registers; capped at 168 the 250-bp batch measured
                                 // DP phase 1848 vs 1824 GCUPS but value 1784 vs 1859 (less room beside the traceback)
#endif
__global__ void __launch_bounds__(32 * fast_warps_per_block(L), (L == 16 ? (C <= 27 ? 6 : 4) : (C <= 20 ? 4 : (C <= 27 ? 3 : (L == 4 ? RSA_FAST_WIDE_BLOCKS : RSA_FAST_WIDE_BLOCKS8)))))
fast_dp_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
               const PairMeta* __restrict__ meta, const FastGroup* __restrict__ groups, int n_groups,
               uint8_t* __restrict__ scratch, DpEnd* __restrict__ ends, RedoHeader* __restrict__ redo,
               uint32_t* __restrict__ redo_list, FastConsts k, int rows_pad, const uint4* __restrict__ tpack,
               const uint2* __restrict__ tflag) {
    extern __shared__ uint8_t fast_smem[];
    uint32_t* lut = reinterpret_cast<uint32_t*>(fast_smem);  // 8 profile words for pair A (low halves), 8 for pair B
    if (threadIdx.x < 16) lut[threadIdx.x] = profile_word(threadIdx.x & 7, k, threadIdx.x >> 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int GPW = fast_groups_per_warp(L), WPB = fast_warps_per_block(L);
    const int gi = lane / L, gl = lane % L;
    // the planner orders a class by increasing window length; blocks are handed out in index order, so walk the
    // class from its long end: the last, partially filled wave then holds the shortest tasks
    const int g_index = (int)(gridDim.x * WPB * GPW) - 1 - ((blockIdx.x * WPB + warp) * GPW + gi);
    FastGroup grp;
    grp.a = 0xFFFFFFFFu; grp.b = 0xFFFFFFFFu; grp.dir_off = 0; grp.qlen = 0; grp.rows = 0;
    if (g_index < n_groups) grp = groups[g_index];
    const bool live = grp.a != 0xFFFFFFFFu;
    uint8_t* tcodes = fast_smem + kFastLutBytes + (size_t)(warp * GPW + gi) * rows_pad;
    constexpr int kRingWords = fast_ring_slots(L) * ((C + 3) / 4) * 32;  // per warp: row slots x words x lanes
    uint32_t* ring = reinterpret_cast<uint32_t*>(fast_smem + kFastLutBytes + (size_t)WPB * GPW * rows_pad) + warp * kRingWords;

    // ---- staging: target profiles into shared memory, query selectors into registers ----------------
    // Two forms of the targets.  (1) ASCII windows (explicit windows of a batch): every lane of a group converts the rows
    // r = gl, gl + L, ... with byte loads.  (2) Windows of the resident reference in its packed form (2 bits per base +
    // a "not ACGT" bit plane, pack_reference_kernel): the WARP stages one window after the other -- its 64-base units
    // arrive with one 128-bit and one 64-bit load per lane, coalesced, land in a small per-warp buffer, and the 32 lanes
    // unpack the rows of that window into the group's row codes.
    bool bad_a = false, bad_b = false;
    int rows = 0, tlen_a = 0, tlen_b = 0, qlen = 0;
    uint32_t toff_a = 0, toff_b = 0;
    const uint8_t *qa = nullptr, *qb = nullptr;
    FastGeom geo = fast_geom(L * C);
    if (live) {
        const PairMeta ma = meta[grp.a], mb = meta[grp.b];
        qlen = grp.qlen;
        geo = fast_geom(qlen);
        rows = grp.rows;
        tlen_a = ma.tlen; tlen_b = mb.tlen;
        toff_a = ma.toff; toff_b = mb.toff;
        qa = qbuf + ma.qoff; qb = qbuf + mb.qoff;
        if (!tpack) {
            const uint8_t* ta = tbuf + ma.toff;
            const uint8_t* tb = tbuf + mb.toff;
            // kStageBatch rows of both windows are loaded before the first one is used: the loads of a batch overlap
            // (one exposed memory latency per batch instead of one per row -- the profile showed the staging warps
            // waiting on the long scoreboard for half of their time)
            constexpr int kStageBatch = 8;
            for (int r0 = gl; r0 < rows; r0 += L * kStageBatch) {
                uint32_t xa[kStageBatch], xb[kStageBatch];
#pragma unroll
                for (int j = 0; j < kStageBatch; ++j) {
                    const int r = r0 + j * L;
                    xa[j] = (r < tlen_a) ? (uint32_t)ta[r] : 0u;
                    xb[j] = (r < tlen_b) ? (uint32_t)tb[r] : 0u;
                }
#pragma unroll
                for (int j = 0; j < kStageBatch; ++j) {
                    const int r = r0 + j * L;
                    if (r < rows) {
                        uint32_t ca = 5u, cb = 5u;  // rows past a pair's own window
                        if (r < tlen_a) { ca = base_code(nibble_of((uint8_t)xa[j])); bad_a |= (ca == 0xFu); ca = min(ca, 5u); }
                        if (r < tlen_b) { cb = base_code(nibble_of((uint8_t)xb[j])); bad_b |= (cb == 0xFu); cb = min(cb, 5u); }
                        tcodes[r] = (uint8_t)(ca | (cb << 4));
                    }
                }
            }
        }
    }
    if (tpack) {  // (uniform over the launch)
        uint8_t* stage = fast_smem + kFastLutBytes + (size_t)WPB * GPW * rows_pad + (size_t)WPB * kRingWords * 4 + (size_t)warp * kFastStageBytes;
        uint4* wcode = reinterpret_cast<uint4*>(stage);
        uint2* wflag = reinterpret_cast<uint2*>(stage + 16 * kFastStageUnits);
        uint32_t* wbad = reinterpret_cast<uint32_t*>(stage + 24 * kFastStageUnits);
        if (lane == 0) *wbad = 0u;
        for (int w = 0; w < 2 * GPW; ++w) {
            const int src = (w >> 1) * L;  // first lane of the group
            const uint32_t o = __shfl_sync(0xFFFFFFFFu, (w & 1) ? toff_b : toff_a, src);
            const int tl = __shfl_sync(0xFFFFFFFFu, (w & 1) ? tlen_b : tlen_a, src);
            const int rws = __shfl_sync(0xFFFFFFFFu, rows, src);
            if (!__shfl_sync(0xFFFFFFFFu, live ? 1 : 0, src)) continue;
            const uint32_t u0 = o >> 6;                                             // first 64-base unit of the window
            const int nu = (int)(((o & 63u) + (uint32_t)tl + 63u) >> 6);           // units it touches (<= kFastStageUnits)
            __syncwarp();
            for (int u = lane; u < nu; u += 32) { wcode[u] = __ldg(tpack + u0 + u); wflag[u] = __ldg(tflag + u0 + u); }
            __syncwarp();
            uint8_t* tc = fast_smem + kFastLutBytes + (size_t)(warp * GPW + (w >> 1)) * rows_pad;
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(wcode);
            const uint32_t* fw = reinterpret_cast<const uint32_t*>(wflag);
            bool bad = false;
            for (int r = lane; r < rws; r += 32) {
                uint32_t c = 5u;  // rows past this pair's own window
                if (r < tl) {
                    const uint32_t pos = (o & 63u) + (uint32_t)r;
                    const uint32_t code = (cw[pos >> 4] >> (2u * (pos & 15u))) & 3u;
                    const bool other = (fw[pos >> 5] >> (pos & 31u)) & 1u;        // N (code 0) or a symbol outside ACGTN (code 1)
                    c = other ? (code == 0u ? 4u : 5u) : code;
                    bad |= other && code != 0u;
                }
                if (w & 1) tc[r] |= (uint8_t)(c << 4); else tc[r] = (uint8_t)c;    // (row r is this lane's in both passes)
            }
            if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) *wbad |= 1u << w;
        }
        __syncwarp();
        const uint32_t wb = *wbad;
        bad_a |= (wb >> (2 * gi)) & 1u;
        bad_b |= (wb >> (2 * gi + 1)) & 1u;
    }
    const int ncols = live ? ((gl < geo.rem) ? geo.C : geo.C - 1) : 0;
    const int col0 = live ? fast_lane_col0(geo, gl) : 0;
    uint32_t qsel[C];
    bool has_n = false;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        uint32_t ca = 0, cb = 0, msel = 0x4444u;  // mask selector: bytes of 0xFFFFFFFF (keep) / of 0 (N: replace)
        if (c < ncols) {
            ca = base_code(nibble_of(qa[col0 + c]));
            cb = base_code(nibble_of(qb[col0 + c]));
            bad_a |= (ca == 0xFu);
            bad_b |= (cb == 0xFu);
            if (ca >= 4u) { msel &= 0xFF00u; ca = 0; }
            if (cb >= 4u) { msel &= 0x00FFu; cb = 0; }
            has_n |= (msel != 0x4444u);
        }
        // sub bytes: [0] = X[ca], [1] = 0 (sign of a non-negative byte), [2] = Y[cb], [3] = 0; the upper half
        // carries the PRMT selector that expands to the per-half "query base is not N" mask
        qsel[c] = ca | ((8u | ca) << 4) | ((4u + cb) << 8) | ((12u + cb) << 12) | ((msel ^ 0x4444u) << 16);
    }
    const unsigned gmask = ((1u << L) - 1u) << (L * gi);
    bad_a = (__ballot_sync(0xFFFFFFFFu, bad_a) & gmask) != 0;
    bad_b = (__ballot_sync(0xFFFFFFFFu, bad_b) & gmask) != 0;
    const bool warp_has_n = __any_sync(0xFFFFFFFFu, has_n);
    int nsteps = ((rows + 3) & ~3) + L - 1;  // rows rounded up: the last row block is flushed in-loop
    if (!live) nsteps = 0;
#pragma unroll
    for (int off = 16; off >= L; off >>= 1) nsteps = max(nsteps, __shfl_xor_sync(0xFFFFFFFFu, nsteps, off));
    __syncthreads();  // lut + this warp's target codes

    uint32_t* dir = reinterpret_cast<uint32_t*>(scratch + grp.dir_off);
    uint32_t bestkey;
    int firstrow[2];
    if (warp_has_n) FastDp<L, C, true, FLEX>::run(k, tcodes, lut, rows, nsteps, gl, ncols, geo.W, qsel, dir, ring, lane, bestkey, firstrow);
    else FastDp<L, C, false, FLEX>::run(k, tcodes, lut, rows, nsteps, gl, ncols, geo.W, qsel, dir, ring, lane, bestkey, firstrow);

    // ---- end cell per pair (half 0 = a, half 1 = b) --------------------------------------------------
    // Every lane holds its first-in-reference-order maximum cell; lanes own increasing column ranges, so among
    // the lanes reaching the pair's maximum the winner is the smallest (8-row block, lane).  Lane 0 of the group
    // then owns pair a and lane 1 pair b: they publish the DpEnd for the traceback kernel.
    DpEnd my_end;
    my_end.score = 0; my_end.qend = 0; my_end.tend = 0; my_end.flags = 0;
    int my_pair = -1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t pi = h ? grp.b : grp.a;
        const bool bad = h ? bad_b : bad_a;
        constexpr int CB = fast_col_bits(L);
        const int kh = half_s(bestkey, h);
        const int mine = kh >> CB;
        int S = mine;
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) S = max(S, __shfl_xor_sync(0xFFFFFFFFu, S, off));
        const bool cand = live && (mine == S) && (S > 0);
        int key = cand ? (((firstrow[h] >> 3) << 8) | gl) : 0x7FFFFFFF;
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off));
        const int wl = (S > 0) ? (key & 0xFF) : 0;  // winner lane of the group (S == 0: trackers stay at (0,0))
        const int qend = __shfl_sync(0xFFFFFFFFu, col0 + (((1 << CB) - 1) - (kh & ((1 << CB) - 1))), wl, L);
        const int tend = __shfl_sync(0xFFFFFFFFu, firstrow[h], wl, L);
        if (!live || (h == 1 && grp.b == grp.a) || gl != h) continue;
        my_pair = (int)pi;
        my_end.score = S;
        my_end.qend = S > 0 ? qend : 0;
        my_end.tend = S > 0 ? tend : 0;
        my_end.flags = bad ? DPF_NEED_EXACT : (DPF_DONE | DPF_LAYOUT_FAST);
    }
    if (my_pair >= 0) {
        ends[my_pair] = my_end;
        if (my_end.flags & DPF_NEED_EXACT) {
            // symbols outside {A,C,G,T,N}: full exact redo (own direction tile), traced by tb_kernel afterwards
            const unsigned int slot = atomicAdd(&redo->count, 1u);
            redo_list[slot] = (uint32_t)my_pair;
        }
    }
}

// ---- redo pass: exact kernel over the pairs the packed kernel flagged --------------------------------
//
// Warps pull entries from the redo list.  DPF_LAYOUT_FAST set: only the end cell is recomputed (the packed
// kernel's direction tile is valid); clear: the pair gets a fresh exact-layout tile from the redo scratch
// region and diroff[pi] is redirected to it.

template <int C, bool WRITE_DIR>
__device__ __forceinline__ void exact_dp_warp(const uint8_t* q, const uint8_t* tn, int qlen, int tlen, uint8_t* dir,
                                              const Scoring& sc, int lane, int& out_score, int& out_qend, int& out_tend) {
    const int c0 = lane * C;
    int qc[C], Hp[C], E[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int col = c0 + c;
        qc[c] = col < qlen ? (int)nibble_of(q[col]) : (int)kWildcard;
        Hp[c] = 0; E[c] = 0;
    }
    int Hlast = 0, Fout = 0, Hl_prev = 0, best = 0;
    uint32_t bestkey = 0xFFFFFFFFu;
    constexpr int kRowBytes = 16 * C;
    const int nsteps = tlen + 31;
    for (int s = 0; s < nsteps; ++s) {
        int Hl = __shfl_up_sync(0xFFFFFFFFu, Hlast, 1);
        int Fl = __shfl_up_sync(0xFFFFFFFFu, Fout, 1);
        if (lane == 0) { Hl = 0; Fl = 0; }
        const int r = s - lane;
        if (r >= 0 && r < tlen) {
            const int tb = tn[r];
            int diag = Hl_prev, F = Fl;
            uint32_t w[(C + 7) / 8];
#pragma unroll
            for (int kk = 0; kk < (C + 7) / 8; ++kk) w[kk] = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int qb = qc[c];
                int sub = (qb == tb) ? sc.match : -sc.mismatch;
                if (qb == (int)kWildcard || tb == (int)kWildcard) sub = 0;
                const int tmp = diag + sub;
                const int e = E[c];
                const int h = max(max(max(tmp, F), e), 0);
                const int tg = tmp - sc.gap_oe;
                if (WRITE_DIR) {
                    uint32_t d = (h == tmp) ? (tmp >= diag ? 0u : 1u) : (h == F ? 3u : 2u);
                    if (!(tg > F - sc.gap_ext)) d |= 8u;
                    if (!(tg > e - sc.gap_ext)) d |= 4u;
                    w[c >> 3] |= d << ((c & 7) * 4);
                }
                F = max(tg, F - sc.gap_ext);
                E[c] = max(tg, e - sc.gap_ext);
                const int col = c0 + c;
                if (col < qlen && h > 0) {
                    const uint32_t key = ((uint32_t)(r >> 3) << 12) | ((uint32_t)col << 3) | (uint32_t)(r & 7);
                    if (h > best) { best = h; bestkey = key; }
                    else if (h == best && key < bestkey) bestkey = key;
                }
                diag = Hp[c];
                Hp[c] = h;
            }
            Hlast = Hp[C - 1];
            Fout = F;
            if (WRITE_DIR) {
                uint8_t* row = dir + (size_t)r * kRowBytes + lane * (C / 2);
                if (C == 4) *reinterpret_cast<uint16_t*>(row) = (uint16_t)w[0];
                else {
#pragma unroll
                    for (int kk = 0; kk < (C + 7) / 8; ++kk) reinterpret_cast<uint32_t*>(row)[kk] = w[kk];
                }
            }
        }
        Hl_prev = Hl;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int ob = __shfl_down_sync(0xFFFFFFFFu, best, off);
        const uint32_t ok = __shfl_down_sync(0xFFFFFFFFu, bestkey, off);
        if (ob > best || (ob == best && ok < bestkey)) { best = ob; bestkey = ok; }
    }
    best = __shfl_sync(0xFFFFFFFFu, best, 0);
    bestkey = __shfl_sync(0xFFFFFFFFu, bestkey, 0);
    out_score = best;
    if (best == 0) { out_qend = 0; out_tend = 0; }
    else { out_qend = (int)((bestkey >> 3) & 0x1FFu); out_tend = (int)(((bestkey >> 12) << 3) | (bestkey & 7u)); }
}

constexpr int kRedoWarpsPerBlock = 4;

__global__ void __launch_bounds__(32 * kRedoWarpsPerBlock)
exact_redo_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
                  const PairMeta* __restrict__ meta, uint64_t* __restrict__ diroff, uint8_t* __restrict__ scratch,
                  DpEnd* __restrict__ ends, RedoHeader* __restrict__ redo, const uint32_t* __restrict__ redo_list,
                  Scoring sc, int tlen_pad, unsigned long long redo_base, unsigned long long scratch_cap,
                  const unsigned long long* __restrict__ redo_base_dev, unsigned long long* __restrict__ counters) {
    // the redo head-room starts behind the planned tiles: known to the host (host planner) or left in the chunk's
    // PlanHeader by plan_offsets (device planner)
    if (redo_base_dev) redo_base = *redo_base_dev;
    const unsigned long long redo_cap = scratch_cap > redo_base ? scratch_cap - redo_base : 0ull;
    extern __shared__ uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tn = smem + warp * tlen_pad;
    const unsigned int total = redo->count;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[2] = total;  // diagnostics: pairs re-run by this pass
    const unsigned int stride = gridDim.x * kRedoWarpsPerBlock;
    for (unsigned int it = blockIdx.x * kRedoWarpsPerBlock + warp; it < total; it += stride) {
        const uint32_t pi = redo_list[it];
        const PairMeta m = meta[pi];
        const uint8_t* q = qbuf + m.qoff;
        const uint8_t* t = tbuf + m.toff;
        const int qlen = m.qlen, tlen = m.tlen;
        const bool keep_fast_tile = (ends[pi].flags & DPF_LAYOUT_FAST) != 0;
        __syncwarp();
        for (int i = lane; i < tlen; i += 32) tn[i] = (uint8_t)nibble_of(t[i]);
        __syncwarp();
        uint8_t* dir = nullptr;
        bool ok = true;
        if (!keep_fast_tile) {
            unsigned long long off = 0;
            const unsigned long long need = ((unsigned long long)tlen * exact_row_bytes(qlen) + 15ull) & ~15ull;
            if (lane == 0) off = atomicAdd(&redo->scratch_used, need);
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
            ok = off + need <= redo_cap;
            if (ok) {
                dir = scratch + redo_base + off;
                if (lane == 0) diroff[pi] = redo_base + off;
            }
        }
        int score = 0, qend = 0, tend = 0;
        if (ok) {
            const int cls = exact_class_cols(qlen);
            if (keep_fast_tile) {
                if (cls == 4) exact_dp_warp<4, false>(q, tn, qlen, tlen, nullptr, sc, lane, score, qend, tend);
                else if (cls == 8) exact_dp_warp<8, false>(q, tn, qlen, tlen, nullptr, sc, lane, score, qend, tend);
                else exact_dp_warp<16, false>(q, tn, qlen, tlen, nullptr, sc, lane, score, qend, tend);
            } else {
                if (cls == 4) exact_dp_warp<4, true>(q, tn, qlen, tlen, dir, sc, lane, score, qend, tend);
                else if (cls == 8) exact_dp_warp<8, true>(q, tn, qlen, tlen, dir, sc, lane, score, qend, tend);
                else exact_dp_warp<16, true>(q, tn, qlen, tlen, dir, sc, lane, score, qend, tend);
            }
        }
        if (lane == 0) {
            DpEnd e;
            e.score = score; e.qend = qend; e.tend = tend;
            e.flags = ok ? (DPF_DONE | (keep_fast_tile ? DPF_LAYOUT_FAST : 0u)) : DPF_NO_SCRATCH;
            if (!ok) atomicAdd(&counters[1], 1ull);
            ends[pi] = e;
        }
    }
}

}  // namespace rsa

// ---- host-side launchers ------------------------------------------------------------------------------
namespace rsa {

template <int L, int C, bool FLEX = false>
inline void launch_fast_one(cudaStream_t st, const uint8_t* q, const uint8_t* t, const PairMeta* meta,
                            const FastGroup* groups, int n_groups, uint8_t* scratch, DpEnd* ends, RedoHeader* redo,
                            uint32_t* redo_list, const FastConsts& k, int max_rows, const uint4* tpack, const uint2* tflag) {
    const int rows_pad = (max_rows + 15) & ~15;
    constexpr int WPB = fast_warps_per_block(L);
    const int groups_per_block = WPB * fast_groups_per_warp(L);
    const int blocks = (n_groups + groups_per_block - 1) / groups_per_block;
    const size_t smem = kFastLutBytes + (size_t)groups_per_block * rows_pad + (size_t)WPB * fast_ring_slots(L) * ((C + 3) / 4) * 32 * 4 +
                        (tpack ? (size_t)WPB * kFastStageBytes : 0);
    if (smem > 48 * 1024) {
        // Long windows / wide lanes need the opt-in shared-memory limit.  The attribute belongs to the FUNCTION (per
        // device), not to the launch: setting it to this launch's size raced with other workers launching the same
        // instantiation with a larger size (their launch then failed and the chunk's pairs came back untouched).
        // Raise it once per device to the device maximum instead; occupancy depends on the launch's own size only.
        static std::mutex mu;
        static bool done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (dev >= 0 && dev < 64 && !done[dev]) {
            int optin = 0;
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            cudaFuncSetAttribute(fast_dp_kernel<L, C, FLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
            done[dev] = true;
        }
    }
    fast_dp_kernel<L, C, FLEX><<<blocks, 32 * WPB, smem, st>>>(q, t, meta, groups, n_groups, scratch, ends,
                                                                     redo, redo_list, k, rows_pad, tpack, tflag);
}

// Widths a merged ("flex") launch exists for: multiples of four columns per lane; a flex launch of width C takes groups of
// width C - kFlexSlack .. C (FastDp<.., FLEX = true>).
__host__ inline int fast_flex_width(int L, int C) {
    const int b = (C + 3) / 4 * 4;
    const int hi = L == 4 ? kFastMaxC : 32;
    return b > hi ? hi : b;
}

// returns 0, or -1 when C is outside the instantiated range
inline int launch_fast_class(cudaStream_t st, int L, int C, bool flex, const uint8_t* q, const uint8_t* t, const PairMeta* meta,
                             const FastGroup* groups, int n_groups, uint8_t* scratch, DpEnd* ends, RedoHeader* redo,
                             uint32_t* redo_list, const FastConsts& k, int max_rows, const uint4* tpack = nullptr,
                             const uint2* tflag = nullptr) {
#define RSA_FAST_ARGS st, q, t, meta, groups, n_groups, scratch, ends, redo, redo_list, k, max_rows, tpack, tflag
#define RSA_FAST_CASE(l, c) case c: launch_fast_one<l, c, false>(RSA_FAST_ARGS); return 0;
#define RSA_FLEX_CASE(l, c) case c: launch_fast_one<l, c, true>(RSA_FAST_ARGS); return 0;
    if (flex) {
        if (L == 16) {
            switch (C) { RSA_FLEX_CASE(16, 20) RSA_FLEX_CASE(16, 24) RSA_FLEX_CASE(16, 28) RSA_FLEX_CASE(16, 32) default: return -1; }
        }
        if (L == 4) {
            switch (C) {
                RSA_FLEX_CASE(4, 4) RSA_FLEX_CASE(4, 8) RSA_FLEX_CASE(4, 12) RSA_FLEX_CASE(4, 16) RSA_FLEX_CASE(4, 20)
                RSA_FLEX_CASE(4, 24) RSA_FLEX_CASE(4, 28) RSA_FLEX_CASE(4, 32) RSA_FLEX_CASE(4, 36) RSA_FLEX_CASE(4, 40)
                default: return -1;
            }
        }
        switch (C) { RSA_FLEX_CASE(8, 24) RSA_FLEX_CASE(8, 28) RSA_FLEX_CASE(8, 32) default: return -1; }
    }
    if (L == 16) {
        switch (C) {
            RSA_FAST_CASE(16, 17) RSA_FAST_CASE(16, 18) RSA_FAST_CASE(16, 19) RSA_FAST_CASE(16, 20) RSA_FAST_CASE(16, 21)
            RSA_FAST_CASE(16, 22) RSA_FAST_CASE(16, 23) RSA_FAST_CASE(16, 24) RSA_FAST_CASE(16, 25) RSA_FAST_CASE(16, 26)
            RSA_FAST_CASE(16, 27) RSA_FAST_CASE(16, 28) RSA_FAST_CASE(16, 29) RSA_FAST_CASE(16, 30) RSA_FAST_CASE(16, 31)
            RSA_FAST_CASE(16, 32)
            default: return -1;
        }
    }
    if (L == 4) {
        switch (C) {
            RSA_FAST_CASE(4, 2) RSA_FAST_CASE(4, 3) RSA_FAST_CASE(4, 4) RSA_FAST_CASE(4, 5) RSA_FAST_CASE(4, 6) RSA_FAST_CASE(4, 7)
            RSA_FAST_CASE(4, 8) RSA_FAST_CASE(4, 9) RSA_FAST_CASE(4, 10) RSA_FAST_CASE(4, 11) RSA_FAST_CASE(4, 12) RSA_FAST_CASE(4, 13)
            RSA_FAST_CASE(4, 14) RSA_FAST_CASE(4, 15) RSA_FAST_CASE(4, 16) RSA_FAST_CASE(4, 17) RSA_FAST_CASE(4, 18) RSA_FAST_CASE(4, 19)
            RSA_FAST_CASE(4, 20) RSA_FAST_CASE(4, 21) RSA_FAST_CASE(4, 22) RSA_FAST_CASE(4, 23) RSA_FAST_CASE(4, 24) RSA_FAST_CASE(4, 25)
            RSA_FAST_CASE(4, 26) RSA_FAST_CASE(4, 27) RSA_FAST_CASE(4, 28) RSA_FAST_CASE(4, 29) RSA_FAST_CASE(4, 30) RSA_FAST_CASE(4, 31)
            RSA_FAST_CASE(4, 32) RSA_FAST_CASE(4, 33) RSA_FAST_CASE(4, 34) RSA_FAST_CASE(4, 35) RSA_FAST_CASE(4, 36) RSA_FAST_CASE(4, 37)
            RSA_FAST_CASE(4, 38) RSA_FAST_CASE(4, 39) RSA_FAST_CASE(4, 40)
            default: return -1;
        }
    }
    switch (C) {
#if RSA_FAST_L4_MAXQ < 160   // (A/B builds without 4-lane groups)
        RSA_FAST_CASE(8, 1) RSA_FAST_CASE(8, 2) RSA_FAST_CASE(8, 3) RSA_FAST_CASE(8, 4) RSA_FAST_CASE(8, 5) RSA_FAST_CASE(8, 6)
        RSA_FAST_CASE(8, 7) RSA_FAST_CASE(8, 8) RSA_FAST_CASE(8, 9) RSA_FAST_CASE(8, 10) RSA_FAST_CASE(8, 11) RSA_FAST_CASE(8, 12)
        RSA_FAST_CASE(8, 13) RSA_FAST_CASE(8, 14) RSA_FAST_CASE(8, 15) RSA_FAST_CASE(8, 16) RSA_FAST_CASE(8, 17) RSA_FAST_CASE(8, 18)
        RSA_FAST_CASE(8, 19) RSA_FAST_CASE(8, 20)
#endif
        RSA_FAST_CASE(8, 21) RSA_FAST_CASE(8, 22) RSA_FAST_CASE(8, 23) RSA_FAST_CASE(8, 24) RSA_FAST_CASE(8, 25) RSA_FAST_CASE(8, 26)
        RSA_FAST_CASE(8, 27) RSA_FAST_CASE(8, 28) RSA_FAST_CASE(8, 29) RSA_FAST_CASE(8, 30) RSA_FAST_CASE(8, 31) RSA_FAST_CASE(8, 32)
        default: return -1;
    }
#undef RSA_FAST_CASE
#undef RSA_FLEX_CASE
#undef RSA_FAST_ARGS
}

inline void launch_exact_redo(cudaStream_t st, const uint8_t* q, const uint8_t* t, const PairMeta* meta,
                              uint64_t* diroff, uint8_t* scratch, DpEnd* ends, RedoHeader* redo,
                              const uint32_t* redo_list, const Scoring& sc, int max_tlen, int n_sms,
                              unsigned long long redo_base, unsigned long long scratch_cap,
                              const unsigned long long* redo_base_dev, unsigned long long* counters) {
    const int tlen_pad = (max_tlen + 15) & ~15;
    exact_redo_kernel<<<n_sms * 2, 32 * kRedoWarpsPerBlock, (size_t)tlen_pad * kRedoWarpsPerBlock, st>>>(
        q, t, meta, diroff, scratch, ends, redo, redo_list, sc, tlen_pad, redo_base, scratch_cap, redo_base_dev, counters);
}

}  // namespace rsa
