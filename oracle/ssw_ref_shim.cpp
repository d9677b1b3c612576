// oracle/ssw_ref_shim.cpp -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Builds the REFERENCE's CPU extension path -- Aligner::align (src/aligner.cpp:114-210: SSW striped
// Smith-Waterman byte->word + reverse pass + banded traceback from ext/ssw/ssw.c, then the end-bonus
// extension) and Aligner::align_gpu (src/aligner.cpp:13-112, the consumer of the GPU records) -- from the
// reference's own sources where they lie under /root/reference (nothing copied; see oracle/Makefile),
// and exposes batch drivers for bench.py's `--impl reference` arm / cpu_baseline and for tests.
//
// The reference's src/gasal2_ssw.h drags in the GASAL2/CUDA headers; only its result struct is needed
// here, so the header is pre-empted through its include guard and the struct is declared as the
// reference declares it (src/gasal2_ssw.h:31-38).
#define STROBEALIGN_GASAL2_SSW_H
#include <unistd.h>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <pthread.h>
#include <vector>

struct gasal_tmp_res {
    int score;
    int query_start;
    int query_end;
    int ref_start;
    int ref_end;
    std::string cigar_str;
};

#include "ssw/ssw_cpp.cpp"
#include "cigar.cpp"
#include "aligner.cpp"

namespace {
struct Out {
    int32_t *score, *qs, *qe, *rs, *re, *ed;
    char* pool;
    int64_t* coff;  // per pair fixed slot: coff[i] = i * slot
    int slot;
};

void store(const Out& o, int64_t i, const AlignmentInfo& a) {
    o.score[i] = a.sw_score; o.qs[i] = (int32_t)a.query_start; o.qe[i] = (int32_t)a.query_end;
    o.rs[i] = (int32_t)a.ref_start; o.re[i] = (int32_t)a.ref_end; o.ed[i] = (int32_t)a.edit_distance;
    if (o.pool) {
        std::string s = a.cigar.to_string();
        size_t w = std::min<size_t>(s.size(), (size_t)o.slot - 1);
        memcpy(o.pool + i * o.slot, s.data(), w);
        o.pool[i * o.slot + w] = 0;
    }
}
}  // namespace

// Aligner::align over pairs [0,n) with `threads` std::threads (the reference runs one Aligner per worker
// thread, src/main.cpp:455-517).  cigar_pool: n fixed slots of cigar_slot bytes (NUL-terminated text) or NULL.
extern "C" int ssw_ref_align_batch(int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf,
                                   const int64_t* toff, int match, int mismatch, int gap_open, int gap_extend,
                                   int end_bonus, int threads, int32_t* score, int32_t* qs, int32_t* qe,
                                   int32_t* rs, int32_t* re, int32_t* ed, char* cigar_pool, int cigar_slot) {
    if (threads < 1) threads = 1;
    Out o{score, qs, qe, rs, re, ed, cigar_pool, nullptr, cigar_slot};
    auto work = [&](int64_t lo, int64_t hi) {
        Aligner aligner(AlignmentParameters{match, mismatch, gap_open, gap_extend, end_bonus});
        for (int64_t i = lo; i < hi; ++i) {
            std::string q(qbuf + qoff[i], qbuf + qoff[i + 1]);
            std::string t(tbuf + toff[i], tbuf + toff[i + 1]);
            store(o, i, aligner.align(q, t));
        }
    };
    // plain pthreads: this object carries a static libstdc++ and is dlopen()ed into processes that already
    // hold another one, where std::thread's runtime hooks do not survive
    if (threads == 1) { work(0, n); return 0; }
    struct Job { decltype(work)* fn; int64_t lo, hi; };
    std::vector<Job> jobs(threads);
    std::vector<pthread_t> tids(threads);
    for (int k = 0; k < threads; ++k) {
        jobs[k] = Job{&work, n * k / threads, n * (k + 1) / threads};
        pthread_create(&tids[k], nullptr, [](void* p) -> void* { Job* j = (Job*)p; (*j->fn)(j->lo, j->hi); return nullptr; }, &jobs[k]);
    }
    for (int k = 0; k < threads; ++k) pthread_join(tids[k], nullptr);
    return 0;
}

// Aligner::align_gpu on externally supplied GPU-path records (score/start/end/CIGAR text per pair).
extern "C" int ssw_ref_align_gpu_batch(int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf,
                                       const int64_t* toff, int match, int mismatch, int gap_open,
                                       int gap_extend, int end_bonus, const int32_t* g_score,
                                       const int32_t* g_qs, const int32_t* g_qe, const int32_t* g_rs,
                                       const int32_t* g_re, const char* g_cigar_pool, const int64_t* g_coff,
                                       int32_t* score, int32_t* qs, int32_t* qe, int32_t* rs, int32_t* re,
                                       int32_t* ed, char* cigar_pool, int cigar_slot) {
    Out o{score, qs, qe, rs, re, ed, cigar_pool, nullptr, cigar_slot};
    Aligner aligner(AlignmentParameters{match, mismatch, gap_open, gap_extend, end_bonus});
    for (int64_t i = 0; i < n; ++i) {
        std::string q(qbuf + qoff[i], qbuf + qoff[i + 1]);
        std::string t(tbuf + toff[i], tbuf + toff[i + 1]);
        gasal_tmp_res g{g_score[i], g_qs[i], g_qe[i], g_rs[i], g_re[i],
                        std::string(g_cigar_pool + g_coff[i], g_cigar_pool + g_coff[i + 1])};
        store(o, i, aligner.align_gpu(q, t, g));
    }
    return 0;
}

// The Hamming shortcut of extend_seed_part (src/aln.cpp:391-404): hamming_distance (src/aligner.hpp:54-67), the 5 % test
// with the reference's own expression, hamming_align (src/aligner.cpp:254-302) -- the reference's functions, called as
// the reference calls them.  status: 0 shortcut taken, 1 gapped path.
extern "C" int ssw_ref_hamming_batch(int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf, const int64_t* toff,
                                     int match, int mismatch, int end_bonus, int32_t* hamming, int32_t* status,
                                     int32_t* score, int32_t* qs, int32_t* qe, int32_t* rs, int32_t* re, int32_t* ed,
                                     char* cigar_pool, int cigar_slot) {
    Out o{score, qs, qe, rs, re, ed, cigar_pool, nullptr, cigar_slot};
    for (int64_t i = 0; i < n; ++i) {
        const std::string query(qbuf + qoff[i], qbuf + qoff[i + 1]);
        const std::string ref_segm_ham(tbuf + toff[i], tbuf + toff[i + 1]);
        auto hamming_dist = hamming_distance(query, ref_segm_ham);
        hamming[i] = hamming_dist;
        AlignmentInfo info;
        status[i] = 1;
        if (hamming_dist >= 0 && (((float) hamming_dist / query.size()) < 0.05)) {
            info = hamming_align(query, ref_segm_ham, match, mismatch, end_bonus);
            status[i] = 0;
        }
        store(o, i, info);
    }
    return 0;
}
