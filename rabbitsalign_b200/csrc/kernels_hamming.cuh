// kernels_hamming.cuh -- SURVEY 8f rank 3, first half: the Hamming shortcut of the seed extension on the device.
//
// What the reference does on the host for every candidate site before it decides to run Smith-Waterman
// (extend_seed_part, src/aln.cpp:374-431): if the projected window has the read's length,
//   hamming_distance(query, window)                        src/aligner.hpp:54-67
//   (float) distance / |query| < 0.05  ?                   src/aln.cpp:395
//   hamming_align(query, window, match, mismatch, bonus)   src/aligner.cpp:254-302
//     highest_scoring_segment                              src/aligner.cpp:219-252  (Kadane scan, end bonus at both ends)
// and the pair never reaches the GPU.  Here: a warp takes 32 pairs per round.  For each pair in turn the 32 lanes compare 32
// bases per step with coalesced byte loads and a ballot turns every step into one word of that pair's mismatch bit mask
// (shared memory); then every lane takes one pair: the distance is a popcount, the sequential segment scan and the
// '=' / 'X' run encoding run over the lane's own bit mask, which is where the reference's tie rules live (strict '>' keeps the FIRST best segment, a negative running score restarts the
// segment behind the current base, the end bonus counts once per reached read end).
// Output: rsa_ext_alninfo_t (= AlignmentInfo, BAM-style ops) with status 0 when the shortcut applies, 1 when the pair
// needs the gapped path (distance too high, unequal lengths, empty read), 3 when the run list does not fit
// RSA_EXT_CIGAR_INLINE ops; the distance itself goes to `hamming` (-1 for unequal lengths).
#pragma once
#include "common.cuh"
#include "kernels_finish.cuh"

namespace rsa {

constexpr int kHamWarpsPerBlock = 4;
constexpr int kHamMaxWords = 16;   // reads up to 512 bases (the engine's packed limit; longer reads take the gapped path)

// A warp takes 32 pairs per round.  Stage A, cooperative: for each of the 32 pairs in turn the lanes compare 32 bases per
// step (coalesced byte loads) and a ballot makes one word of that pair's mismatch bit mask (shared memory, [pair][word]).
// Stage B, one lane per pair: distance test, segment scan and run encoding over the lane's own bit mask.
__global__ void __launch_bounds__(32 * kHamWarpsPerBlock)
hamming_kernel(const uint8_t* __restrict__ qbuf, const int64_t* __restrict__ qoff, const uint8_t* __restrict__ tbuf,
               const int64_t* __restrict__ toff, const int64_t* __restrict__ win_off, long long n, int match, int mismatch,
               int end_bonus, int32_t* __restrict__ hamming, rsa_ext_alninfo_t* __restrict__ out) {
    __shared__ uint32_t s_mask[kHamWarpsPerBlock][32][kHamMaxWords + 1];   // (+1: lanes walk their own rows, no bank conflicts)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rounds = (n + 31) / 32;
    for (long long rd = (long long)blockIdx.x * kHamWarpsPerBlock + warp; rd < rounds; rd += (long long)gridDim.x * kHamWarpsPerBlock) {
        const long long pi = rd * 32 + lane;     // this lane's pair in stage B
        long long q0 = 0, t0 = 0;
        int len = 0, tlen = -1;
        if (pi < n) {
            q0 = qoff[pi];
            len = (int)(qoff[pi + 1] - q0);
            if (win_off) { t0 = win_off[pi]; tlen = len; }
            else { t0 = toff[pi]; tlen = (int)(toff[pi + 1] - t0); }
        }
        const bool comparable = pi < n && tlen == len && len <= kHamMaxWords * 32;
        __syncwarp();
        // ---- stage A
        for (int p = 0; p < 32; ++p) {
            const int plen = __shfl_sync(0xFFFFFFFFu, comparable ? len : 0, p);
            if (plen == 0) continue;   // (uniform)
            const uint8_t* q = qbuf + __shfl_sync(0xFFFFFFFFu, q0, p);
            const uint8_t* t = tbuf + __shfl_sync(0xFFFFFFFFu, t0, p);
            for (int b = 0; b < plen; b += 32) {
                const int i = b + lane;
                const bool mis = i < plen && q[i] != t[i];
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, mis);
                if (lane == 0) s_mask[warp][p][b >> 5] = w;
            }
        }
        __syncwarp();
        // ---- stage B
        if (pi >= n) continue;
        const uint32_t* mask = s_mask[warp][lane];
        rsa_ext_alninfo_t a;
        a.sw_score = 0; a.edit_distance = 0; a.ref_start = 0; a.ref_end = 0; a.query_start = 0; a.query_end = 0;
        a.n_cigar = 0; a.status = 1;
#pragma unroll
        for (int k = 0; k < RSA_EXT_CIGAR_INLINE; ++k) a.cigar[k] = 0;
        int hd = -1;
        if (comparable) {
            hd = 0;
            for (int w = 0; w < (len + 31) >> 5; ++w) hd += __popc(mask[w]);
            // (float) hamming_dist / query.size() < 0.05 : float division, compared as double (src/aln.cpp:395)
            const bool pass = len > 0 && (double)((float)hd / (float)len) < 0.05;
            if (pass) {
                // highest_scoring_segment (src/aligner.cpp:219-252)
                int start = 0, score = end_bonus, best_start = 0, best_end = 0, best = 0;
                for (int i = 0; i < len; ++i) {
                    const bool mis = (mask[i >> 5] >> (i & 31)) & 1u;
                    score += mis ? -mismatch : match;
                    if (score < 0) { start = i + 1; score = 0; }
                    if (score > best) { best_start = start; best = score; best_end = i + 1; }
                }
                if (score + end_bonus > best) { best = score + end_bonus; best_end = len; best_start = start; }
                // hamming_align (src/aligner.cpp:254-302)
                OpList cg{a.cigar, 0, RSA_EXT_CIGAR_INLINE, false};
                if (best_start > 0) cg.push(CG_SOFT, (uint32_t)best_start);
                int mismatches = 0, counter = 0;
                bool prev_match = false, first = true;
                for (int i = best_start; i < best_end; ++i) {
                    const bool is_match = !((mask[i >> 5] >> (i & 31)) & 1u);
                    mismatches += is_match ? 0 : 1;
                    if (!first && is_match != prev_match) { cg.push(prev_match ? CG_EQ : CG_X, (uint32_t)counter); counter = 0; }
                    counter++;
                    prev_match = is_match;
                    first = false;
                }
                if (!first) cg.push(prev_match ? CG_EQ : CG_X, (uint32_t)counter);
                if (len - best_end > 0) cg.push(CG_SOFT, (uint32_t)(len - best_end));
                if (cg.overflow) {
                    a.status = 3;
#pragma unroll
                    for (int k = 0; k < RSA_EXT_CIGAR_INLINE; ++k) a.cigar[k] = 0;
                } else {
                    a.status = 0;
                    a.n_cigar = (int16_t)cg.n;
                    a.sw_score = best; a.edit_distance = mismatches;
                    a.ref_start = best_start; a.ref_end = best_end; a.query_start = best_start; a.query_end = best_end;
                }
            }
        }
        hamming[pi] = hd;
        out[pi] = a;
    }
}

}  // namespace rsa
