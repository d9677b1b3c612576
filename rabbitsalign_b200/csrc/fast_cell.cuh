// fast_cell.cuh -- the packed (two pairs per register) Smith-Waterman cell recipe, shared verbatim by the
// product kernel (kernels_fast.cuh) and by tools/dpx_microbench.cu, whose "bare recipe" rate is the own-recipe
// ceiling bench.py reports next to the kernel.
//
// Round-2 recipe ("clamp facts"): the ALU pipe (DPX, LOP3, IADD3, PRMT) binds this kernel, so everything that can run
// on the FMA pipe does.  Compared with the round-1 recipe (carry-trick flags: 3 IADD3 + 3 LOP3 + shift/insert per cell)
//   * the four direction facts are CLEAN 0/1 values per half, one VIADDMNMX.S16x2.RELU each:
//     clamp(a + (-b) + c, 0, 1); the negated operands are IMADs (FMA pipe);
//   * clean 0/1 halves need no masks: the facts are merged into a nibble, and two nibbles into a byte, with multiply-adds
//     x * 2^k + y on the packed register (IMAD, FMA pipe again: no LOP3/SHF).  HFMA2 does the same on the fp16 view (0/1
//     halves are denormals whose bit pattern is their integer value; exact below 2048) but measured slower: mixing fp16
//     and integer work on the FMA pipe costs issue slots (bare recipe 2544 vs 2798 GCUPS, kernel 1901 vs 1970);
//   * the substitution profile carries (score - gap_open_extend) as SIGNED bytes, so S = H(diag) + profile is already the
//     "open a gap from the diagonal" value both E' and F' need (no separate add), and H takes its "+ gap_oe" inside the
//     fused add+max.
// ALU pipe per cell pair: PRMT, VIMNMX3, 3 VIADDMNMX, 4 VIADDMNMX.RELU, 1/2 VIMNMX3 (row key), 1/4 PRMT (word).
// FMA pipe per cell pair: 8.5 IMAD (S, three negations, key, 3.5 merges).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include "common.cuh"

namespace rsa {

#ifndef RSA_CELL_IMAD_MERGE
#define RSA_CELL_IMAD_MERGE 1   // 0: merge the direction facts with HFMA2 on denormal bit patterns instead (A/B builds)
#endif

constexpr int kBias = 64;  // every stored half = value + kBias; E,F >= -(mismatch+gap_oe) > -kBias

struct FastConsts {
    uint32_t zero;     // (kBias, kBias)
    uint32_t oe16;     // per-half s16 (+gap_oe) for the fused add+max of H
    uint32_t neg_e;    // per-half s16 (-gap_ext) for VIADDMNMX
    uint32_t one16;    // (1, 1): upper clamp of a direction fact
    uint32_t c_b1;     // b1 = c_b1 - S  : per half  -S + 1 - gap_ext   (gap facts)
    uint32_t c_b2;     // b2 = c_b2 - S  : per half  -S - gap_oe        (H != diagonal)
    uint32_t c_b3;     // b3 = c_b3 - F  : per half  -F                 (max(F,E,0) != F)
    uint32_t sub_n;    // S increment of a query N: (-gap_oe, -gap_oe - 1)
    uint32_t i2, i4, i16;   // multipliers of the fact merges (in registers so that the products stay IMADs)
    uint32_t h2, h4, h16;   // (A/B builds: the same as fp16 pairs 2.0, 4.0, 16.0)
    uint32_t k32, k64, minus1;  // multipliers kept in registers so that the products below stay IMADs (FMA pipe)
    int match, mismatch, gap_oe;
    int bias;
};

__host__ __device__ inline uint32_t pair16(int v) { return ((uint32_t)(v & 0xFFFF) << 16) | (uint32_t)(v & 0xFFFF); }

// 32-bit constant whose halves are (hi, lo) as an ARITHMETIC sum hi * 65536 + lo: used where a 32-bit subtraction of a
// packed register with strictly positive halves always borrows from the high half (so `hi` carries a +1, see below)
__host__ __device__ inline uint32_t ring32(int hi, int lo) { return (uint32_t)hi * 65536u + (uint32_t)lo; }

// per-column constant of the maximum-tracking key: (2^B - 1 - column) - bias * 2^B in both halves, as a ring constant.
// B = 5 column bits for 8- and 16-lane groups (C <= 32 columns per lane), 6 for 4-lane groups (C <= 40).
template <int B = 5>
__host__ __device__ constexpr uint32_t key_colconst(int c) { return (uint32_t)(((1 << B) - 1) - c - kBias * (1 << B)) * 0x00010001u; }

__host__ inline FastConsts make_fast_consts(const Scoring& sc) {
    FastConsts k;
    k.bias = kBias;
    k.match = sc.match; k.mismatch = sc.mismatch; k.gap_oe = sc.gap_oe;
    k.zero = pair16(kBias);
    k.oe16 = pair16(sc.gap_oe);
    k.neg_e = pair16(-sc.gap_ext);
    k.one16 = pair16(1);
    // c - X over a packed register X whose halves are all > c: the low half goes negative and borrows exactly 1 from the
    // high half, always; the high constant is therefore one larger
    k.c_b1 = ring32(1 - sc.gap_ext + 1, 1 - sc.gap_ext);
    k.c_b2 = ring32(-sc.gap_oe + 1, -sc.gap_oe);
    k.c_b3 = ring32(1, 0);
    k.sub_n = ((uint32_t)((-sc.gap_oe - 1) & 0xFFFF) << 16) | (uint32_t)((-sc.gap_oe) & 0xFFFF);
    k.h2 = 0x40004000u;
    k.h4 = 0x44004400u;
    k.h16 = 0x4C004C00u;
    k.i2 = 2u; k.i4 = 4u; k.i16 = 16u;
    k.k32 = 32u;
    k.k64 = 64u;
    k.minus1 = 0xFFFFFFFFu;
    return k;
}

// 4-byte SIGNED profile of one target-row code: byte q = score(query code q, target) - gap_oe (pair A, low halves), and
// one less for pair B (high halves): the profile value is always negative, so the 32-bit add S = H + profile always
// carries out of the low half, and B's "- 1" takes that carry back.
// code 0..3 = base, 4 = N (scores 0 against everything), 5 = row past the pair's own window (everything mismatches, so
// H only decays there and can never reach the pair's maximum).
__host__ __device__ inline uint32_t profile_word(uint32_t code, const FastConsts& k, int half) {
    const int mm = -k.mismatch - k.gap_oe - half, ma = k.match - k.gap_oe - half, nn = -k.gap_oe - half;
    if (code < 4u) return ((uint32_t)(mm & 0xFF) * 0x01010101u & ~(0xFFu << (8u * code))) | ((uint32_t)(ma & 0xFF) << (8u * code));
    if (code == 4u) return (uint32_t)(nn & 0xFF) * 0x01010101u;
    return (uint32_t)(mm & 0xFF) * 0x01010101u;
}

// The three non-DPX instructions of the recipe.  The device versions are single PTX instructions; the host versions
// restate them so that tests/cell_host_check.cu can run the very same fast_cell() on the CPU (the DPX intrinsics have
// host implementations in the CUDA headers).
//
// prmt in its default mode: selector nibble bit 3 replicates the sign of the chosen byte (the __byte_perm intrinsic
// masks that bit away).
__host__ __device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (sel >> (4 * i)) & 0xFu;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xFFu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        d |= byte << (8 * i);
    }
    return d;
#endif
}

__host__ __device__ __forceinline__ uint32_t bitsel(uint32_t mask, uint32_t a, uint32_t b) {  // mask ? a : b, one LOP3
    return (a & mask) | (b & ~mask);
}

// a*b + c with b in a register: stays an IMAD (FMA pipe), which this ALU-bound recipe leaves idle otherwise
__host__ __device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a * b + c;
#endif
}

// a*b + c on fp16 pairs WITHOUT flush-to-zero: on halves holding small integers (denormal / first-binade bit patterns,
// value < 2048) and b = 2^k this is exact integer arithmetic on the bit patterns (FMA pipe)
__host__ __device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    uint32_t d = 0;
    for (int i = 0; i < 2; ++i) {
        __half_raw ra, rb, rc;
        ra.x = (unsigned short)(a >> (16 * i)); rb.x = (unsigned short)(b >> (16 * i)); rc.x = (unsigned short)(c >> (16 * i));
        const double v = (double)__half2float(__half(ra)) * (double)__half2float(__half(rb)) + (double)__half2float(__half(rc));
        const __half_raw rd = __float2half_rn((float)v);   // exact products and sums in this recipe: no double rounding
        d |= (uint32_t)rd.x << (16 * i);
    }
    return d;
#endif
}

// One packed cell (pair A in the low halves, pair B in the high halves).
//   s      = H(r-1,c-1) + sub(r,c) - gap_oe   (biased; computed in phase 1 as H + signed profile byte)
//   F, e   = F(r,c), E(r,c) entering the cell
// out: h = H(r,c); fn, en = F(r,c+1), E(r+1,c); nib = the cell's direction nibble in bits 3..0 of each half, in the
//      reference's polarity for the gap bits (fast_layout.cuh):
//        8  F NOT opened here (F - ext >= diag + sub - oe)     4  E NOT opened here
//        2  H != diagonal  (max(F,E,0) > diag + sub)           1  max(F,E,0) != F
//      key = ((h - bias) << B) | (2^B - 1 - column) for the first-maximum tracking; `colconst` carries both the column
//      term and the -bias*2^B correction (ring constant), so the key of an all-zero cell is just its column term.
__host__ __device__ __forceinline__ void fast_cell(const FastConsts& k, uint32_t s, uint32_t F, uint32_t e, uint32_t colconst, uint32_t kmul,
                                          uint32_t& h, uint32_t& fn, uint32_t& en, uint32_t& nib, uint32_t& key) {
    const uint32_t u = __vimax3_s16x2(F, e, k.zero);
    h = __viaddmax_s16x2(s, k.oe16, u);                          // max(diag + sub, F, E, 0)
    fn = __viaddmax_s16x2(F, k.neg_e, s);
    en = __viaddmax_s16x2(e, k.neg_e, s);
    const uint32_t b1 = imad(s, k.minus1, k.c_b1);               // per half: -s + 1 - ext   (FMA pipe)
    const uint32_t b2 = imad(s, k.minus1, k.c_b2);               //           -s - oe
    const uint32_t b3 = imad(F, k.minus1, k.c_b3);               //           -F
    const uint32_t xf = __viaddmin_s16x2_relu(F, b1, k.one16);   // clamp(F - ext - (s) + 1, 0, 1)
    const uint32_t xe = __viaddmin_s16x2_relu(e, b1, k.one16);
    const uint32_t nd = __viaddmin_s16x2_relu(u, b2, k.one16);   // clamp(u - (diag + sub), 0, 1)
    const uint32_t nf = __viaddmin_s16x2_relu(u, b3, k.one16);   // clamp(u - F, 0, 1)
#if RSA_CELL_IMAD_MERGE
    nib = imad(imad(xf, k.i2, xe), k.i4, imad(nd, k.i2, nf));    // 8 xf + 4 xe + 2 nd + nf   (FMA pipe)
#else
    nib = hfma2(hfma2(xf, k.h2, xe), k.h4, hfma2(nd, k.h2, nf));
#endif
    key = imad(h, kmul, colconst);                               // ((h-bias) << B) | (2^B - 1 - column), FMA pipe
}

// Gathering the direction nibbles of four columns into one word (nibble k of each half = column k of the word): two
// nibbles make a byte with one multiply-add (FMA pipe), one byte-permute puts the two bytes of each half in place.
__host__ __device__ __forceinline__ uint32_t dir_pair(const FastConsts& k, uint32_t nib_even, uint32_t nib_odd) {   // bits 7..0 of each half
#if RSA_CELL_IMAD_MERGE
    return imad(nib_odd, k.i16, nib_even);
#else
    return hfma2(nib_odd, k.h16, nib_even);
#endif
}
__host__ __device__ __forceinline__ uint32_t dir_word(uint32_t p01, uint32_t p23) { return prmt(p01, p23, 0x6240u); }

}  // namespace rsa
