// kernels_exact.cuh -- exact int32 wavefront kernel (one warp per pair).
//
// Computes what gasal_local_kernel<LOCAL, WITH_TB, FALSE> computes
// (reference GASAL2/src/kernels/local_kernel_template.h:45-60 cell, :118-430 loop nest) for ONE pair per
// warp, with a different traversal: lane l owns query columns [l*C, l*C+C) and walks down the target
// rows one anti-diagonal step behind lane l-1; H and F of a lane's last column travel to the next lane
// by shuffle.  Because the visiting order differs from the reference's (8-row block, column, row in
// block), the "first maximum" rule (:58-59,158,165) is reproduced with an explicit key
// (row>>3, column, row&7): the end cell is the cell with the largest H and, among those, the smallest key.
//
// This is the general path: any nibble alphabet (bases are compared on `ascii & 0xF`, 0xE scores 0:
// gasal_kernels.h:48-51), any |q| <= 512, any |t|.  The packed s16x2 DPX kernel (kernels_fast.cuh) takes
// the bulk of the pairs and hands the ones it declines to this kernel.
#pragma once
#include "common.cuh"

namespace rsa {

constexpr int kExactWarpsPerBlock = 4;

template <int C>
__global__ void __launch_bounds__(32 * kExactWarpsPerBlock)
exact_dp_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
                const PairMeta* __restrict__ meta, const uint32_t* __restrict__ list, int n_list,
                const uint64_t* __restrict__ dir_off, uint8_t* __restrict__ scratch,
                DpEnd* __restrict__ ends, Scoring sc, int tlen_pad) {
    extern __shared__ uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * kExactWarpsPerBlock + warp;
    if (task >= n_list) return;
    const uint32_t pi = list[task];
    const PairMeta m = meta[pi];
    const uint8_t* q = qbuf + m.qoff;
    const uint8_t* t = tbuf + m.toff;
    const int qlen = m.qlen, tlen = m.tlen;

    // target nibbles of this pair, staged once per warp
    uint8_t* tn = smem + warp * tlen_pad;
    for (int i = lane; i < tlen; i += 32) tn[i] = (uint8_t)nibble_of(t[i]);
    __syncwarp();

    const int c0 = lane * C;
    int qc[C], Hp[C], E[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int col = c0 + c;
        qc[c] = col < qlen ? (int)nibble_of(q[col]) : (int)kWildcard;  // pad = 'N' (host_batch.cpp:142-145)
        Hp[c] = 0;
        E[c] = 0;
    }
    int Hlast = 0, Fout = 0, Hl_prev = 0;
    int best = 0;
    uint32_t bestkey = 0xFFFFFFFFu;
    constexpr int kRowBytes = 16 * C;  // 32 lanes * C/2 bytes
    uint8_t* dir = scratch + dir_off[pi];
    const int nsteps = tlen + 31;
    for (int s = 0; s < nsteps; ++s) {
        int Hl = __shfl_up_sync(0xFFFFFFFFu, Hlast, 1);
        int Fl = __shfl_up_sync(0xFFFFFFFFu, Fout, 1);
        if (lane == 0) { Hl = 0; Fl = 0; }  // left border: H = 0, F = 0 (:123-127)
        const int r = s - lane;
        if (r >= 0 && r < tlen) {
            const int tb = tn[r];
            int diag = Hl_prev;
            int F = Fl;
            uint32_t w[(C + 7) / 8];
#pragma unroll
            for (int k = 0; k < (C + 7) / 8; ++k) w[k] = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int qb = qc[c];
                int sub = (qb == tb) ? sc.match : -sc.mismatch;
                if (qb == (int)kWildcard || tb == (int)kWildcard) sub = 0;
                const int tmp = diag + sub;
                const int e = E[c];
                const int h = max(max(max(tmp, F), e), 0);
                uint32_t d = (h == tmp) ? (tmp >= diag ? 0u : 1u) : (h == F ? 3u : 2u);
                const int tg = tmp - sc.gap_oe;
                if (!(tg > F - sc.gap_ext)) d |= 8u;
                F = max(tg, F - sc.gap_ext);
                if (!(tg > e - sc.gap_ext)) d |= 4u;
                E[c] = max(tg, e - sc.gap_ext);
                w[c >> 3] |= d << ((c & 7) * 4);
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
            uint8_t* row = dir + (size_t)r * kRowBytes + lane * (C / 2);
            if (C == 4) *reinterpret_cast<uint16_t*>(row) = (uint16_t)w[0];
            else {
#pragma unroll
                for (int k = 0; k < (C + 7) / 8; ++k) reinterpret_cast<uint32_t*>(row)[k] = w[k];
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
    if (lane == 0) {
        DpEnd e;
        e.score = best;
        if (best == 0) { e.qend = 0; e.tend = 0; }  // maxXY_x/y stay at their initial 0 (:80-84)
        else {
            e.qend = (int)((bestkey >> 3) & 0x1FFu);
            e.tend = (int)(((bestkey >> 12) << 3) | (bestkey & 7u));
        }
        e.flags = DPF_DONE;
        ends[pi] = e;
    }
}

// class of a query length for this kernel: columns per lane
__host__ __device__ inline int exact_class_cols(int qlen) { return qlen <= 128 ? 4 : (qlen <= 256 ? 8 : 16); }
__host__ __device__ inline int exact_row_bytes(int qlen) { return 16 * exact_class_cols(qlen); }

}  // namespace rsa
