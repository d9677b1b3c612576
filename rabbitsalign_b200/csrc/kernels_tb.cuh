// kernels_tb.cuh -- traceback walker: direction nibbles -> run-length CIGAR bytes + start cell.
//
// One thread per pair, following gasal_get_tb<LOCAL> (reference GASAL2/src/kernels/get_tb.h:16-147)
// step for step: the two-variable mode machine (op_select/op_shift, :74-82), the 63-capped run bytes
// emitted end-to-start (:87-98,113-117), the running score that stops the walk on equality with the
// alignment score (:100-103, N cells are credited +match), the inclusive start cell (:142-145) and the
// byte count (:146).  The direction nibble itself comes from whichever DP kernel produced it; both
// layouts decode to the reference's 4-bit code (common.cuh).
#pragma once
#include "common.cuh"
#include "kernels_exact.cuh"
#include "fast_layout.cuh"

namespace rsa {

struct TbCtx {
    const uint8_t* q;
    const uint8_t* t;
    const uint8_t* dir;
    int qlen, tlen;
    int row_bytes;       // exact layout
    bool fast;
    FastGeom fg;         // fast layout
    int half;            // fast layout: which 16-bit half of the packed words
};

__device__ __forceinline__ uint32_t tb_fetch(const TbCtx& c, int i, int j) {
    if (!c.fast) {
        const uint8_t b = c.dir[(size_t)i * c.row_bytes + (j >> 1)];
        return (j & 1) ? (b >> 4) : (b & 0xFu);
    }
    // fast layout stores [F not opened, E not opened, not_diag, not_f] per cell (fast_layout.cuh): the two gap bits are the
    // reference's bits 3 and 2 as they are; rebuild the reference's 2-bit source code.
    const uint32_t f = fast_fetch_flags(c.fg, c.dir, i, j, c.half);
    uint32_t lo2;
    if (!(f & 2u)) {
        const uint32_t qb = nibble_of(c.q[j]), tb = nibble_of(c.t[i]);
        const bool mism = (qb != tb) && qb != kWildcard && tb != kWildcard;  // tmp < diag <=> sub < 0
        lo2 = mism ? 1u : 0u;
    } else {
        lo2 = (f & 1u) ? 2u : 3u;
    }
    return lo2 | (f & 0xCu);
}

// One walk.  Bytes 0..inline_cap-1 go to `inl`; if `full` != nullptr every byte also goes there.
__device__ inline int tb_walk(const TbCtx& c, int score, int tend, int qend, const Scoring& sc,
                              uint8_t* inl, int inline_cap, uint8_t* full, int* out_i, int* out_j) {
    int i = tend, j = qend;
    int sel = 3, sh = 0;
    uint32_t prev = 0;
    int count = 0, n_ops = 0, cur = 0;
    while (i >= 0 && j >= 0) {
        const uint32_t cell = tb_fetch(c, i, j);
        const uint32_t op = (cell >> sh) & (uint32_t)sel;
        const uint32_t out = (op == 0 || sel == 3) ? op : (uint32_t)sh;
        sel = (op == 0 || (op == 1 && sel == 3)) ? 3 : 1;
        sh = (op == 0 || (op == 1 && sel == 3)) ? 0 : ((op == 2 || op == 3) ? (int)op : sh);
        if (count < 63 && out == prev) {
            count++;
        } else {
            if (count > 0) {
                const uint8_t b = (uint8_t)(prev | (uint32_t)(count << 2));
                if (n_ops < inline_cap) inl[n_ops] = b;
                if (full) full[n_ops] = b;
                n_ops++;
            }
            count = 1;
        }
        if ((out == 2 || out == 3) && prev != out) cur -= sc.gap_oe;
        else if (out == 2 || out == 3) cur -= sc.gap_ext;
        else if (out == 1) cur -= sc.mismatch;
        else cur += sc.match;
        if (cur == score) break;
        prev = out;
        if (out != 3) i--;
        if (out != 2) j--;
    }
    const uint8_t b = (uint8_t)(prev | (uint32_t)(count << 2));
    if (n_ops < inline_cap) inl[n_ops] = b;
    if (full) full[n_ops] = b;
    n_ops++;
    *out_i = i;
    *out_j = j;
    return n_ops;
}

// Everything the walker needs besides the pair's own DpEnd.
struct TbArgs {
    const uint8_t* qbuf;
    const uint8_t* tbuf;
    const PairMeta* meta;
    const uint32_t* info;     // bit 0: which half of the packed words; bits 16..: status of pairs no kernel ran on
    const uint64_t* dir_off;
    const uint8_t* scratch;
    rsa_ext_result_t* res;
    Scoring sc;
    uint8_t* arena;
    unsigned long long* arena_used;
    unsigned long long arena_cap;
};

// Trace one pair back and write its 64-byte record.  Shared (not inlined) by the stand-alone traceback
// kernel and by the packed DP kernel, which traces its own pairs right after computing them while their
// direction tiles are still in L2.
//
// Long CIGARs (n_ops > RSA_EXT_RLE_INLINE): the record keeps the FIRST inline bytes (so records are byte-deterministic)
// and the full string goes to the chunk's arena; its arena offset (scheduling dependent, internal) is parked in
// ends[pi].qend/tend, which nothing reads after the traceback (the host fetches `ends` only when an arena was used).
__device__ __noinline__ void tb_one_pair(const TbArgs& a, int pi, DpEnd e, DpEnd* __restrict__ ends) {
    rsa_ext_result_t r;
#pragma unroll
    for (int k = 0; k < RSA_EXT_RLE_INLINE; ++k) r.rle[k] = 0;
    const PairMeta m = a.meta[pi];
    TbCtx c;
    c.q = a.qbuf + m.qoff;
    c.t = a.tbuf + m.toff;
    c.dir = a.scratch + a.dir_off[pi];
    c.qlen = m.qlen;
    c.tlen = m.tlen;
    c.fast = (e.flags & DPF_LAYOUT_FAST) != 0;
    c.row_bytes = exact_row_bytes(m.qlen);
    c.half = (int)(a.info[pi] & 1u);
    c.fg = fast_geom(m.qlen);
    int si, sj;
    int n_ops = tb_walk(c, e.score, e.tend, e.qend, a.sc, r.rle, RSA_EXT_RLE_INLINE, nullptr, &si, &sj);
    r.status = 0;
    if (n_ops > RSA_EXT_RLE_INLINE) {
        // rare: long CIGAR.  Reserve n_ops bytes in the chunk's arena and walk again writing all of them.
        const unsigned long long off = atomicAdd(a.arena_used, (unsigned long long)n_ops);
        if (off + (unsigned long long)n_ops <= a.arena_cap) {
            tb_walk(c, e.score, e.tend, e.qend, a.sc, r.rle, 0, a.arena + off, &si, &sj);
            ends[pi].qend = (int32_t)(uint32_t)(off & 0xFFFFFFFFull);
            ends[pi].tend = (int32_t)(uint32_t)(off >> 32);
        } else {
            r.status = 2;  // cannot happen: the arena is sized for the worst case of the chunk
        }
    }
    r.score = e.score;
    r.query_start = sj;
    r.query_end = e.qend;
    r.ref_start = si;
    r.ref_end = e.tend;
    r.n_ops = (int16_t)n_ops;
    a.res[pi] = r;
}

constexpr int kTbThreads = 128;

// Stand-alone traceback: every pair of the chunk that is not traced yet (exact-kernel pairs, redone pairs)
// and the failed records of pairs no kernel ran on.
__global__ void __launch_bounds__(kTbThreads) tb_kernel(TbArgs a, int n, DpEnd* __restrict__ ends) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n) return;
    const DpEnd e = ends[pi];
    if (e.flags & DPF_TRACED) return;
    if (!(e.flags & DPF_DONE)) {
        // not aligned: empty sequence / window longer than max_target_len (host decided), or no scratch (redo)
        rsa_ext_result_t r;
#pragma unroll
        for (int k = 0; k < RSA_EXT_RLE_INLINE; ++k) r.rle[k] = 0;
        r.score = 0; r.query_start = -1; r.query_end = -1; r.ref_start = -1; r.ref_end = -1;
        r.n_ops = 0; r.status = (e.flags & DPF_NO_SCRATCH) ? (int16_t)4 : (int16_t)(a.info[pi] >> 16);
        if (r.status == 0) {  // the host routed this pair to a kernel and no kernel wrote its DpEnd: a launch was lost
            r.status = 5;
            atomicAdd(a.arena_used + 3, 1ull);
        }
        a.res[pi] = r;
        return;
    }
    tb_one_pair(a, pi, e, ends);
}

// Traceback of the packed kernel's pairs in GROUP order: threads 2g and 2g+1 trace pairs a and b of group g.
// The two pairs share every direction word (low/high halves) and, with equal flanks, walk the same cells, so
// their loads hit the same 16-byte chunks: one HBM sector serves both walkers.
struct FastGroupRef { uint32_t a, b; uint64_t dir_off; uint16_t qlen, rows; };  // == FastGroup (kernels_fast.cuh)

__global__ void __launch_bounds__(kTbThreads) tb_groups_kernel(TbArgs a, const FastGroupRef* __restrict__ groups,
                                                               int n_groups, DpEnd* __restrict__ ends) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = t >> 1;
    if (g >= n_groups) return;
    const FastGroupRef grp = groups[g];
    if (grp.a == 0xFFFFFFFFu) return;
    if ((t & 1) && grp.b == grp.a) return;
    const int pi = (int)((t & 1) ? grp.b : grp.a);
    const DpEnd e = ends[pi];
    if ((e.flags & (DPF_DONE | DPF_LAYOUT_FAST | DPF_TRACED)) != (DPF_DONE | DPF_LAYOUT_FAST)) return;
    tb_one_pair(a, pi, e, ends);
    ends[pi].flags = e.flags | DPF_TRACED;
}

}  // namespace rsa
