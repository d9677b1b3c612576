// kernels_finish.cuh -- "next row" of the scope table: what the reference does on the
// HOST with every GPU record, done on the device:
//
//   gasal_fail                      src/pc.cpp:446-478      accept the record or send the pair to the CPU aligner
//   Aligner::Align_gpu              ext/ssw/ssw_cpp.cpp:396-429 -> my_ssw_align / parse_cigar_string (ssw.c:925-992)
//   ConvertAlignment                ext/ssw/ssw_cpp.cpp:54-90    leading / trailing soft clips
//   CalculateNumberMismatchOnly     ext/ssw/ssw_cpp.cpp:212-247  M -> '=', NM = X + I + D
//   Aligner::align_gpu              src/aligner.cpp:13-112       greedy ungapped extension to both read ends, end bonus
//
// One thread per pair turns the 64-byte record into an `rsa_ext_alninfo_t` = the fields of `AlignmentInfo`
// (src/aligner.hpp:20-30) with the CIGAR as BAM-style ops, so a pipeline that adopts it skips the CIGAR text
// round trip (text -> ints -> text -> Cigar) and the per-pair string streams of the reference.
#pragma once
#include "common.cuh"

namespace rsa {

// strobealign's op codes (src/cigar.hpp:11-21)
enum : uint32_t { CG_INS = 1, CG_DEL = 2, CG_SOFT = 4, CG_EQ = 7, CG_X = 8 };

struct OpList {
    uint32_t* ops;
    int n, cap;
    bool overflow;
    __device__ void push(uint32_t op, uint32_t len) {  // Cigar::push (src/cigar.hpp:51-58): merge equal neighbours
        if (n > 0 && (ops[n - 1] & 0xFu) == op) { ops[n - 1] += len << 4; return; }
        if (n < cap) ops[n++] = (len << 4) | op; else overflow = true;
    }
};

constexpr int kFinishThreads = 128;
constexpr int kFinishMaxOps = 96;

__global__ void __launch_bounds__(kFinishThreads)
finish_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf, const PairMeta* __restrict__ meta,
              const rsa_ext_result_t* __restrict__ res, int n, Scoring sc, int end_bonus,
              rsa_ext_alninfo_t* __restrict__ out) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n) return;
    const rsa_ext_result_t r = res[pi];
    const PairMeta m = meta[pi];
    const uint8_t* q = qbuf + m.qoff;
    const uint8_t* t = tbuf + m.toff;
    const int qlen = m.qlen, tlen = m.tlen;
    rsa_ext_alninfo_t a;
    a.sw_score = 0; a.edit_distance = 0; a.ref_start = 0; a.ref_end = 0; a.query_start = 0; a.query_end = 0;
    a.n_cigar = 0; a.status = 0;
#pragma unroll
    for (int k = 0; k < RSA_EXT_CIGAR_INLINE; ++k) a.cigar[k] = 0;
    if (r.status == 1) {  // window longer than 2000: the sentinel both aligners return (src/aligner.cpp:18-24)
        a.edit_distance = 100000; a.sw_score = -1000000; a.status = 2;
        out[pi] = a;
        return;
    }
    // ---- gasal_fail (src/pc.cpp:466-478)
    bool fail = r.status != 0 || r.n_ops <= 0 || r.score == 0 || r.query_start < 0 || r.query_end < 0 ||
                r.ref_start < 0 || r.ref_end < 0 || r.query_end >= qlen || r.ref_end >= tlen;
    if (!fail && r.n_ops > RSA_EXT_RLE_INLINE) { a.status = 3; out[pi] = a; return; }  // long CIGAR: record path
    uint32_t core_ops[kFinishMaxOps], ext_ops[kFinishMaxOps];
    OpList core{core_ops, 0, kFinishMaxOps, false};
    int edits = 0;
    if (!fail) {
        // the reference's decode order (gasal2_ssw.cpp:184-243): bytes last-to-first, equal neighbours merged;
        // 'M' becomes '=' and NM = X + I + D (CalculateNumberMismatchOnly)
        long qsum = 0;
        for (int u = r.n_ops - 1; u >= 0; --u) {
            const uint32_t op = r.rle[u] & 3u, len = r.rle[u] >> 2;
            if (op == 0) { core.push(CG_EQ, len); qsum += len; }
            else if (op == 1) { core.push(CG_X, len); qsum += len; edits += (int)len; }
            else if (op == 2) { core.push(CG_DEL, len); edits += (int)len; }
            else { core.push(CG_INS, len); qsum += len; edits += (int)len; }
        }
        if (qsum != (long)(r.query_end - r.query_start + 1)) fail = true;  // calculate_cigar_length check
    }
    if (fail) { a.status = 1; out[pi] = a; return; }

    // ---- Aligner::align_gpu (src/aligner.cpp:40-109)
    int sw = (int)(uint16_t)r.score;  // Alignment::sw_score is a uint16_t
    int ref_start = r.ref_start, ref_end = r.ref_end + 1, query_start = r.query_start, query_end = r.query_end + 1;
    OpList fin{a.cigar, 0, RSA_EXT_CIGAR_INLINE, false};
    {   // extension to the read start
        int qstart = query_start, rstart = ref_start, score = sw, ed = edits;
        OpList front{ext_ops, 0, kFinishMaxOps, false};
        while (qstart > 0 && rstart > 0) {
            qstart--; rstart--;
            if (q[qstart] == t[rstart]) { score += sc.match; front.push(CG_EQ, 1); }
            else { score -= sc.mismatch; front.push(CG_X, 1); ed++; }
        }
        if (qstart == 0 && score + end_bonus > sw) {
            if (query_start > 0) {  // soft clip replaced by the reversed extension
                if (front.overflow) fin.overflow = true;
                for (int k = front.n - 1; k >= 0; --k) fin.push(front.ops[k] & 0xFu, front.ops[k] >> 4);
            }
            query_start = 0; ref_start = rstart; sw = score + end_bonus; edits = ed;
        } else if (query_start > 0) {
            fin.push(CG_SOFT, (uint32_t)query_start);  // ConvertAlignment
        }
    }
    for (int k = 0; k < core.n; ++k) fin.push(core.ops[k] & 0xFu, core.ops[k] >> 4);
    if (core.overflow) fin.overflow = true;
    {   // extension to the read end
        int qend = query_end, rend = ref_end, score = sw, ed = edits;
        OpList back{ext_ops, 0, kFinishMaxOps, false};
        while (qend < qlen && rend < tlen) {
            if (q[qend] == t[rend]) { score += sc.match; back.push(CG_EQ, 1); }
            else { score -= sc.mismatch; back.push(CG_X, 1); ed++; }
            qend++; rend++;
        }
        if (qend == qlen && score + end_bonus > sw) {
            if (query_end < qlen) {
                if (back.overflow) fin.overflow = true;
                for (int k = 0; k < back.n; ++k) fin.push(back.ops[k] & 0xFu, back.ops[k] >> 4);
            }
            query_end = qlen; ref_end = rend; sw = score + end_bonus; edits = ed;
        } else if (qlen - query_end > 0) {
            fin.push(CG_SOFT, (uint32_t)(qlen - query_end));
        }
    }
    if (fin.overflow) {  // more than RSA_EXT_CIGAR_INLINE ops: the caller takes the record path for this pair
        a.status = 3; a.n_cigar = 0;
#pragma unroll
        for (int k = 0; k < RSA_EXT_CIGAR_INLINE; ++k) a.cigar[k] = 0;
        out[pi] = a;
        return;
    }
    a.sw_score = sw; a.edit_distance = edits; a.ref_start = ref_start; a.ref_end = ref_end;
    a.query_start = query_start; a.query_end = query_end; a.n_cigar = (int16_t)fin.n;
    out[pi] = a;
}

}  // namespace rsa
