// fast_cell.cuh -- the packed (two pairs per register) Smith-Waterman cell recipe, shared verbatim by the
// product kernel (kernels_fast.cuh) and by tools/dpx_microbench.cu, whose "bare recipe" rate is the issue-rate
// ceiling bench.py reports the kernel against.
#pragma once
#include <cstdint>
#include "common.cuh"

namespace rsa {

constexpr int kBias = 64;  // every stored half = value + kBias; E,F >= -(mismatch+gap_oe) > -kBias

struct FastConsts {
    uint32_t zero;     // (kBias, kBias)
    uint32_t neg_x16;  // per-half s16 (-mismatch) for the fused add+max of H
    uint32_t neg_xoe;  // subtract mismatch + gap_oe
    uint32_t neg_e;    // per-half s16 (-gap_ext) for VIADDMNMX
    uint32_t k_f, k_e, k_d, k_n;
    uint32_t x_pair;   // (mismatch, mismatch): biased profile value of a zero-scoring cell
    uint32_t prof_match;  // match + mismatch (byte)
    uint32_t k32, k64, one, minus1;  // multipliers kept in registers so that the adds below stay IMADs (FMA pipe)
    int bias;
};

__host__ __device__ inline uint32_t pair16(int v) { return ((uint32_t)(v & 0xFFFF) << 16) | (uint32_t)(v & 0xFFFF); }

// per-column constant of the maximum-tracking key: (2^B - 1 - column) - bias * 2^B in both halves, as a ring constant.
// B = 5 column bits for 8- and 16-lane groups (C <= 32 columns per lane), 6 for 4-lane groups (C <= 40).
template <int B = 5>
__host__ __device__ constexpr uint32_t key_colconst(int c) { return (uint32_t)(((1 << B) - 1) - c - kBias * (1 << B)) * 0x00010001u; }

__host__ inline FastConsts make_fast_consts(const Scoring& sc) {
    FastConsts k;
    k.bias = kBias;
    k.zero = pair16(kBias);
    k.neg_x16 = pair16(-sc.mismatch);
    k.neg_xoe = (uint32_t)(0u - (uint32_t)(sc.mismatch + sc.gap_oe) * 0x00010001u);
    k.neg_e = pair16(-sc.gap_ext);
    k.k_f = pair16(sc.gap_ext + 0x7FFF);
    k.k_e = pair16(sc.gap_ext + 0x3FFF);
    k.k_d = pair16(0x1FFF + sc.mismatch);  // the flag is taken from h - s, and s carries +mismatch
    k.k_n = pair16(0x0FFF);
    k.x_pair = pair16(sc.mismatch);
    k.prof_match = (uint32_t)(sc.match + sc.mismatch);
    k.k32 = 32u;
    k.k64 = 64u;
    k.one = 1u;
    k.minus1 = 0xFFFFFFFFu;
    return k;
}

// PTX prmt in its default mode: selector nibble bit 3 replicates the sign of the chosen byte (the
// __byte_perm intrinsic masks that bit away).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t bitsel(uint32_t mask, uint32_t a, uint32_t b) {  // mask ? a : b, one LOP3
    return (a & mask) | (b & ~mask);
}

// a*b + c with b in a register: stays an IMAD (FMA pipe), which this ALU-bound recipe leaves idle otherwise
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One packed cell (pair A in the low halves, pair B in the high halves).
//   s      = H(r-1,c-1) + sub(r,c) + mismatch   (biased ring value, computed in phase 1)
//   F, e   = F(r,c), E(r,c) entering the cell
// out: h = H(r,c); fn, en = F(r,c+1), E(r+1,c); fl = direction facts in bits 15..12 of each half
//      (15 F opened, 14 E opened, 13 H != diagonal, 12 max(F,E,0) != F; bits 11..0 are garbage);
//      key = ((h - bias) << 5) | (31 - column) for the first-maximum tracking; `colconst` carries both the column
//      term and the -bias*32 correction (ring constant), so the key of an all-zero cell is just its column term.
// ALU pipe: VIMNMX3, 3x VIADDMNMX (H takes its "- mismatch" inside the fused add+max, so diag+sub is never
// materialised), 3x IADD3, 3x LOP3, one add.  FMA pipe: the adds below written as IMADs.
__device__ __forceinline__ void fast_cell(const FastConsts& k, uint32_t s, uint32_t F, uint32_t e, uint32_t colconst, uint32_t kmul,
                                          uint32_t& h, uint32_t& fn, uint32_t& en, uint32_t& fl, uint32_t& key) {
    const uint32_t tg = s + k.neg_xoe;
    const uint32_t u = __vimax3_s16x2(F, e, k.zero);
    h = __viaddmax_s16x2(s, k.neg_x16, u);                        // max(diag + sub, F, E, 0)
    fn = __viaddmax_s16x2(F, k.neg_e, tg);
    en = __viaddmax_s16x2(e, k.neg_e, tg);
    const uint32_t fo = fn - F + k.k_f;                           // bit15: F opened
    const uint32_t eo = en - e + k.k_e;                           // bit14: E opened
    const uint32_t nd = imad(s, k.minus1, imad(h, k.one, k.k_d));    // bit13: H != diagonal  (h - (s - x) + 0x1FFF, FMA pipe)
    const uint32_t nf = u - F + k.k_n;                            // bit12: max(F,E,0) != F
    fl = bitsel(0x80008000u, fo, eo);
    fl = bitsel(0xC000C000u, fl, nd);
    fl = bitsel(0xE000E000u, fl, nf);
    key = imad(h, kmul, colconst);                                // ((h-bias) << B) | (2^B - 1 - column), FMA pipe
}

}  // namespace rsa
