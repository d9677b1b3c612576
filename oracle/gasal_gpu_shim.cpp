// oracle/gasal_gpu_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The REFERENCE's own GPU path, unmodified, as a checker and comparator on the GPU box: GASAL2
// (GASAL2/src/{args_parser,host_batch,ctors,interfaces,res}.cpp + gasal_align.cu with its kernel headers) and the
// reference's batch driver solve_ssw_on_gpu (src/gasal2_ssw.cpp:19-256) are compiled from the sources where they lie
// under /root/reference (nothing copied; recipe in oracle/Makefile, output oracle/_ref/libgasal_gpu.so, sm_100a,
// MAX_QUERY_LEN=500 N_CODE=0x4E without N_PENALTY like the reference's build.sh:23) and driven the way src/pc.cpp:644-672
// drives them: slices of at most STREAM_BATCH_SIZE (512) pairs, one blocking call per slice, per worker `thread_id`.
//
// Used by tests/test_gpu_reference_gpu.py (the product against the reference's real CUDA kernels on the same B200)
// and tests/perf_reference_gpu.py (the reference GPU path's GCUPS on this box).
#include "gasal2_ssw.h"  // the reference's header (-I$(REF_ROOT)/src); GASAL2 headers pre-included from GASAL2/src

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" int gasal_gpu_slice_size(void) { return STREAM_BATCH_SIZE; }

// out5: n x 5 int32 (score, query_start, query_end, ref_start, ref_end); cigars: n strings of `cigar_stride` bytes,
// NUL-terminated (truncated if longer).  Returns 0.
extern "C" int gasal_gpu_batch(int thread_id, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf,
                               const int64_t* toff, int32_t* out5, char* cigars, int cigar_stride) {
    std::vector<gasal_tmp_res> res;
    std::vector<std::string> qs, ts;
    for (int64_t lo = 0; lo < n; lo += STREAM_BATCH_SIZE) {
        const int64_t hi = lo + STREAM_BATCH_SIZE < n ? lo + STREAM_BATCH_SIZE : n;
        qs.clear();
        ts.clear();
        for (int64_t i = lo; i < hi; ++i) {
            qs.emplace_back(qbuf + qoff[i], (size_t)(qoff[i + 1] - qoff[i]));
            ts.emplace_back(tbuf + toff[i], (size_t)(toff[i + 1] - toff[i]));
        }
        solve_ssw_on_gpu(thread_id, res, qs, ts);  // default scores 2/8/12/1 (src/gasal2_ssw.h:46-47)
        for (int64_t i = lo; i < hi; ++i) {
            const gasal_tmp_res& r = res[(size_t)(i - lo)];
            int32_t* o = out5 + 5 * i;
            o[0] = r.score; o[1] = r.query_start; o[2] = r.query_end; o[3] = r.ref_start; o[4] = r.ref_end;
            if (cigars) {
                char* c = cigars + (size_t)i * cigar_stride;
                const size_t m = r.cigar_str.size() < (size_t)cigar_stride - 1 ? r.cigar_str.size() : (size_t)cigar_stride - 1;
                memcpy(c, r.cigar_str.data(), m);
                c[m] = 0;
            }
        }
    }
    return 0;
}
