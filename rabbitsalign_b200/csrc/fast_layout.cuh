// fast_layout.cuh -- geometry of the packed kernel's column ownership and direction-bit scratch.
//
// A "group" is L lanes (4 for |q| <= 160, 8 for 161..256, 16 for 257..512) working on TWO pairs of equal query length (pair A in the low 16-bit halves of every
// packed register, pair B in the high halves).  The query's columns are dealt to the 8 lanes in order:
// with C = ceil(qlen/L), the first `rem` lanes own C columns and the remaining lanes own C-1, so that
// every owned column is a real query base (no padding columns exist).
//
// Direction scratch of a group: 16-byte chunks holding a 4-row x 4-column cell tile (both pairs).  Two chunks of one lane
// side by side (4 rows x 8 columns) fill one 32-byte SECTOR, which the DP kernel writes with ONE 256-bit store per thread
// (never a partial sector: partial sectors are read-modify-written by the L2); the sectors of two consecutive row blocks
// sit side by side in one 64-byte UNIT (8 rows x 8 columns of one lane), the granularity of an HBM access, in which the
// traceback, which moves diagonally, finds the next ~8 path cells.  For target row r, lane l, word w (w = column-in-lane / 4),
// with W = ceil(C/4) rounded up to an even number and rows padded to a multiple of 8:
//     unit   = ((r>>3)*L + l)*(W/2) + (w>>1)
//     uint32 index = ((unit*2 + ((r>>2)&1))*2 + (w&1))*4 + (r&3)
// (Round 1 ordered the chunks [row block][word][lane]: a sector then held the same word of two neighbouring LANES, i.e.
// columns C apart, and every 16-byte chunk the traceback touched cost a 64-byte HBM access of which three quarters were
// useless: 6.2 KB read per pair; [row block][lane][word] brought 3.7 KB.)
// low half = pair A, high half = pair B; nibble k = (column-in-lane & 3) sits at bits [4k,4k+4) of its half:
//     bit3 F of the next column NOT opened from the diagonal (extended; = reference bit3)
//     bit2 E of the next row NOT opened from the diagonal (= reference bit2)
//     bit1 not_diag (H != diag+sub)
//     bit0 not_f    (max(F,E,0) != F, i.e. with not_diag: source is E -> code 2, else F -> code 3)
#pragma once
#include "common.cuh"

namespace rsa {

// lanes per group as a function of the query length (one rule for planner, kernels and traceback)
// 4 lanes up to 160 bases (C <= 40 columns per lane: the per-row overhead of a lane -- shuffles, profile loads, maximum
// tracking, loop control -- is spread over twice the columns of an 8-lane group and the wavefront skew shrinks from 7 to 3
// idle steps per window), 8 lanes up to 256 (C <= 32), 16 lanes beyond.
#ifndef RSA_FAST_L4_MAXQ
#define RSA_FAST_L4_MAXQ 160   // (A/B builds: 0 = no 4-lane groups)
#endif
__host__ __device__ inline int fast_lanes_for(int qlen) { return qlen <= RSA_FAST_L4_MAXQ ? 4 : (qlen <= 256 ? 8 : 16); }

struct FastGeom {
    int L;    // lanes per group
    int C;    // columns of the widest lanes
    int rem;  // number of lanes owning C columns (1..8); the rest own C-1
    int W;    // 32-bit direction words per lane and row as STORED (even: the last one may be padding)
};

__host__ __device__ inline FastGeom fast_geom(int qlen) {
    FastGeom g;
    g.L = fast_lanes_for(qlen);
    g.C = (qlen + g.L - 1) / g.L;
    g.rem = qlen - g.L * (g.C - 1);
    g.W = (((g.C + 3) / 4) + 1) & ~1;
    return g;
}

// rows of a group's tile: whole 8-row units
__host__ __device__ inline int fast_tile_rows(int rows) { return (rows + 7) & ~7; }

// bytes of direction scratch of one group (two pairs) with `rows` target rows
__host__ __device__ inline uint64_t fast_dir_bytes(const FastGeom& g, int rows) {
    return (uint64_t)fast_tile_rows(rows) * g.L * g.W * 4u;
}

// first column owned by lane l
__host__ __device__ inline int fast_lane_col0(const FastGeom& g, int l) {
    return l <= g.rem ? l * g.C : g.rem * g.C + (l - g.rem) * (g.C - 1);
}

__device__ __forceinline__ uint32_t fast_fetch_flags(const FastGeom& g, const uint8_t* dir, int i, int j,
                                                     int half) {
    int lane, cc;
    const int wide = g.rem * g.C;
    if (j < wide) { lane = j / g.C; cc = j - lane * g.C; }
    else { const int jj = j - wide; const int k = jj / (g.C - 1); lane = g.rem + k; cc = jj - k * (g.C - 1); }
    const uint32_t word = reinterpret_cast<const uint32_t*>(dir)[((((((size_t)(i >> 3) * g.L + lane) * (g.W >> 1) + (cc >> 3)) * 2 + ((i >> 2) & 1)) * 2 + ((cc >> 2) & 1)) << 2) + (i & 3)];
    const uint32_t h16 = half ? (word >> 16) : (word & 0xFFFFu);
    return (h16 >> (4 * (cc & 3))) & 0xFu;
}

}  // namespace rsa
