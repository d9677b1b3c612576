// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the REFERENCE's own CUDA kernel sources for the host, untouched and from where they lie
// under /root/reference (never copied into this repo), so that the C restatement in sw_oracle.c and
// the CUDA product path can be checked against the reference's actual code:
//
//     GASAL2/src/kernels/pack_rc_seqs.h          gasal_pack_kernel
//     GASAL2/src/kernels/local_kernel_template.h gasal_local_kernel<LOCAL, WITH_TB, FALSE>
//     GASAL2/src/kernels/get_tb.h                gasal_get_tb<LOCAL>
//
// The three headers are plain C in CUDA clothing; this file supplies the handful of CUDA names they
// use (vector types, thread indices, qualifiers), the Int2Type/SameType helpers and enums that
// gasal_kernels.h:8-26 and gasal.h:37-65 declare, and a driver that follows
// src/gasal2_ssw.cpp:19-256 (pad with 'N' to a multiple of 8, byte offsets, one "thread" per pair,
// reverse RLE decode).  Build: oracle/Makefile -> oracle/_ref/libgasal_ref.so (git-ignored).
//
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <string>
#include <vector>

#ifndef MAX_QUERY_LEN
#define MAX_QUERY_LEN 500  // build.sh:23
#endif
#ifndef N_CODE
#define N_CODE 0x4E  // build.sh:23
#endif

// ---- CUDA names used by the three headers -------------------------------------------------------
struct uint4 { uint32_t x, y, z, w; };
struct short2 { short x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return short2{x, y}; }
struct dim3_ { unsigned x = 0, y = 0, z = 0; };
static thread_local dim3_ blockIdx, threadIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __constant__ static thread_local
#define register
using std::max;
using std::min;

// ---- declarations the kernels expect (gasal_kernels.h:8-56, gasal.h:37-65,86-95) ---------------
template <int Val> struct Int2Type { typedef enum { val_ = Val } val__; };
template <typename X, typename Y> struct SameType { enum { result = 0 }; };
template <typename T> struct SameType<T, T> { enum { result = 1 }; };
#define SAMETYPE(a, b) (SameType<a, b>::result)
enum comp_start { WITHOUT_START, WITH_START, WITH_TB };
enum Bool { FALSE, TRUE };
enum algo_type { UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW };
enum data_source { NONE, QUERY, TARGET, BOTH };
enum operation_on_seq { FORWARD_NATURAL, REVERSE_NATURAL, FORWARD_COMPLEMENT, REVERSE_COMPLEMENT };
struct gasal_res {
    int32_t *aln_score, *query_batch_end, *target_batch_end, *query_batch_start, *target_batch_start;
    uint8_t *cigar;
    uint32_t *n_cigar_ops;
};
typedef struct gasal_res gasal_res_t;
__constant__ int32_t _cudaGapO, _cudaGapOE, _cudaGapExtend, _cudaMatchScore, _cudaMismatchScore;
#define N_VALUE (N_CODE & 0xF)
// no N_PENALTY in the reference build (GASAL2/Makefile:42-50, build.sh:23): gasal_kernels.h:48-51
#define DEV_GET_SUB_SCORE_LOCAL(score, rbase, gbase) \
    score = (rbase == gbase) ? _cudaMatchScore : -_cudaMismatchScore; \
    score = ((rbase == N_VALUE) || (gbase == N_VALUE)) ? 0 : score;

// resolved through -I$(REF_ROOT)/GASAL2/src/kernels (oracle/Makefile); nothing is copied
#include "pack_rc_seqs.h"
#include "local_kernel_template.h"
#include "get_tb.h"

// ---- driver: what solve_ssw_on_gpu + gasal_aln_async do around the kernels ------------------------
extern "C" int gasal_ref_max_query_len() { return MAX_QUERY_LEN; }

extern "C" int gasal_ref_batch(int n, const char *qbuf, const int64_t *qoff, const char *tbuf,
                               const int64_t *toff, int match, int mismatch, int gap_open_m1,
                               int gap_ext, int32_t *score, int32_t *qs, int32_t *qe, int32_t *rs,
                               int32_t *re, int32_t *n_ops_out, char *cigar_pool, int64_t pool_cap,
                               int64_t *cigar_off) {
    // gasal_copy_subst_scores: gasal_align.cu:329-339
    _cudaMatchScore = match;
    _cudaMismatchScore = mismatch;
    _cudaGapO = gap_open_m1;
    _cudaGapExtend = gap_ext;
    _cudaGapOE = gap_open_m1 + gap_ext;

    // gasal_host_batch_fill: host_batch.cpp:79-153 (pad to x8 with N_CODE; byte offsets)
    std::vector<uint32_t> qo(n), to(n), ql(n), tl(n);
    std::vector<uint8_t> uq, ut;
    for (int i = 0; i < n; ++i) {
        qo[i] = (uint32_t)uq.size();
        to[i] = (uint32_t)ut.size();
        ql[i] = (uint32_t)(qoff[i + 1] - qoff[i]);
        tl[i] = (uint32_t)(toff[i + 1] - toff[i]);
        uq.insert(uq.end(), qbuf + qoff[i], qbuf + qoff[i + 1]);
        while (uq.size() % 8) uq.push_back(N_CODE);
        ut.insert(ut.end(), tbuf + toff[i], tbuf + toff[i + 1]);
        while (ut.size() % 8) ut.push_back(N_CODE);
    }
    if (uq.empty() || ut.empty()) return -1;  // gasal_align.cu:36-43 would exit()
    uint32_t qbytes = (uint32_t)uq.size(), tbytes = (uint32_t)ut.size();
    std::vector<uint32_t> pq(qbytes / 8 + 1), pt(tbytes / 8 + 1);

    // gasal_pack_kernel as one thread: gasal_align.cu:180-194
    blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 1; gridDim.x = 1;
    gasal_pack_kernel((uint32_t *)uq.data(), (uint32_t *)ut.data(), pq.data(), pt.data(),
                      (int)(qbytes / 8), (int)(tbytes / 8), qbytes / 4, tbytes / 4);

    std::vector<int32_t> a_score(n), a_qe(n), a_te(n), a_qs(n), a_ts(n);
    std::vector<uint32_t> a_nops(n);
    gasal_res_t res{a_score.data(), a_qe.data(), a_te.data(), a_qs.data(), a_ts.data(), nullptr, a_nops.data()};

    // direction tiles: tile index * n_tasks + tid (local_kernel_template.h:257, get_tb.h:57)
    uint64_t max_tiles = 0;
    for (int i = 0; i < n; ++i) {
        uint64_t q8 = (ql[i] + 7) & ~7u, t8 = (tl[i] + 7) & ~7u;
        max_tiles = std::max<uint64_t>(max_tiles, q8 * t8 / 32 + 1);
    }
    std::vector<uint4> tb((size_t)(max_tiles * (uint64_t)n));
    std::vector<uint32_t> qlens_dev(ql);  // get_tb overwrites this with n_ops (get_tb.h:146)

    blockDim.x = 1; threadIdx.x = 0; gridDim.x = (unsigned)n;
    for (int tid = 0; tid < n; ++tid) {
        blockIdx.x = (unsigned)tid;
        gasal_local_kernel<Int2Type<LOCAL>, Int2Type<WITH_TB>, Int2Type<FALSE>>(
            pq.data(), pt.data(), qlens_dev.data(), tl.data(), qo.data(), to.data(), &res, nullptr,
            tb.data(), n);
    }
    // CIGAR bytes land in the (re-used) unpacked query buffer at the query's byte offset
    // (gasal_align.h:50); give every pair extra head-room after its slot so that an over-long
    // walk cannot clobber its neighbour here (on the GPU that is a data race).
    std::vector<std::vector<uint8_t>> per_pair(n);
    for (int tid = 0; tid < n; ++tid) {
        blockIdx.x = (unsigned)tid;
        std::vector<uint8_t> slot(8192 + ql[tid] + 8);
        // run the walker on a private slot: same code, cigar_offset[tid] = 0
        std::vector<uint32_t> off1(n, 0);
        gasal_get_tb<Int2Type<LOCAL>>(slot.data(), qlens_dev.data(), tl.data(), off1.data(), tb.data(), &res, n);
        per_pair[tid].assign(slot.begin(), slot.begin() + qlens_dev[tid]);
    }

    // decode: src/gasal2_ssw.cpp:184-243
    int64_t pos = 0;
    for (int j = 0; j < n; ++j) {
        cigar_off[j] = pos;
        const std::vector<uint8_t> &c = per_pair[j];
        int n_cigar_ops = (int)qlens_dev[j];
        std::string s;
        int last_op = c[n_cigar_ops - 1] & 3;
        int count = c[n_cigar_ops - 1] >> 2;
        for (int u = n_cigar_ops - 2; u >= 0; u--) {
            int curr_op = c[u] & 3;
            if (curr_op == last_op) {
                count += c[u] >> 2;
            } else {
                s += std::to_string(count);
                s += "MXDI"[last_op];
                count = c[u] >> 2;
            }
            last_op = curr_op;
        }
        s += std::to_string(count);
        s += "MXDI"[last_op];
        if (pos + (int64_t)s.size() > pool_cap) return -2;
        memcpy(cigar_pool + pos, s.data(), s.size());
        pos += (int64_t)s.size();
        score[j] = a_score[j];
        qs[j] = a_qs[j];
        qe[j] = a_qe[j];
        rs[j] = a_ts[j];
        re[j] = a_te[j];
        n_ops_out[j] = n_cigar_ops;
    }
    cigar_off[n] = pos;
    return 0;
}
