// integration/gasal2_ssw.h -- drop-in replacement for the reference's src/gasal2_ssw.h.
//
// Same include guard, same result struct, same entry point and default arguments
// (reference src/gasal2_ssw.h:31-47), same macros that the callers in src/pc.cpp use
// (STREAM_BATCH_SIZE, src/pc.cpp:644-672), but no GASAL2/CUDA headers: the GPU sits behind the C ABI of
// include/rsa_ext.h (librsa_ext.so).  Replace src/gasal2_ssw.{h,cpp} with this pair and link
// -lrsa_ext instead of -lgasal; src/pc.cpp, src/aligner.cpp and ext/ssw need no edit.
#ifndef STROBEALIGN_GASAL2_SSW_H
#define STROBEALIGN_GASAL2_SSW_H
#include <unistd.h>
#include <vector>
#include <cmath>
#include <sstream>
#include <cassert>
#include <string>
#include <iostream>
#include <fstream>

#define NB_STREAMS 1
#define THREAD_NUM_MAX 256
// Callers slice their todo lists into STREAM_BATCH_SIZE pairs (src/pc.cpp:644-672); the engine itself takes
// any n >= 1, so this may be raised (e.g. to 1 << 20, one call per chunk) without touching the engine.
#ifndef STREAM_BATCH_SIZE
#define STREAM_BATCH_SIZE 512
#endif
#define MAX_QUERY_LEN 500
#define MAX_TARGET_LEN 2000
#define DEBUG
#define MAX(a, b) (a > b ? a : b)

#ifdef RSA_EXT_WINDOWS
#define RSA_EXT_ALNINFO  // the window build takes AlignmentInfo from the device as well
#endif
#ifdef RSA_EXT_ALNINFO
#include "rsa_ext.h"
#endif

struct gasal_tmp_res {
    int score;
    int query_start;
    int query_end;
    int ref_start;
    int ref_end;
    std::string cigar_str;
#ifdef RSA_EXT_ALNINFO
    // Optional build (INTEGRATION.md "take AlignmentInfo straight from the device"): what gasal_fail +
    // Aligner::align_gpu would compute from this record, filled by the engine's finish kernel.
    rsa_ext_alninfo_t aln;
    int aln_end_bonus = -1;  // end bonus (-L) the device used; < 0: no device record, use the host path
#endif
};

void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &todo_querys,
                      std::vector<std::string> &todo_refs, int match_score = 2, int mismatch_score = 8,
                      int gap_open_score = 12, int gap_extend_score = 1);

// The GPU worker `thread_id` is placed on (thread_id % visible devices, RSA_EXT_DEVICES caps the count): the other glue
// layers (integration/seed_glue.cpp, sam_glue.cpp) keep a worker's state on the same device.
int rsa_ext_veneer_device(int thread_id);

#ifdef RSA_EXT_WINDOWS
// SURVEY 8(f) rank 1, caller half: a reference window named by (contig, start, length) instead of a std::string built
// with substr (src/pc.cpp:214-242, 333-368).  The todo list of a chunk becomes std::vector<RsaWindow>; the whole list
// goes down in ONE call (no 512-pair slice copies, src/pc.cpp:644-672) and the engine reads the windows from the
// reference resident in HBM (rsa_ext_set_reference / rsa_ext_submit_ref_windows).  The string is only materialised
// for the rare record the host has to finish itself (gasal_fail -> Aligner::align, long CIGARs).
struct RsaWindow {
    const std::string *contig;  // references.sequences[ref_id]
    uint32_t ref_id;
    uint32_t start;
    uint32_t len;
    RsaWindow(const std::vector<std::string> &sequences, size_t ref_id_, size_t start_, size_t want)
        : contig(&sequences[ref_id_]), ref_id((uint32_t)ref_id_) {
        // std::string::substr semantics: the length is clipped to the end of the contig
        const size_t sz = contig->size();
        start = (uint32_t)(start_ < sz ? start_ : sz);
        len = (uint32_t)(want < sz - start ? want : sz - start);
    }
    size_t size() const { return len; }
    size_t length() const { return len; }
    std::string str() const { return contig->substr(start, len); }
    operator std::string() const { return str(); }
};

// One call per chunk: `sequences` is references.sequences (uploaded once per GPU, shared by all workers).
void solve_ssw_on_gpu_windows(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &todo_querys,
                              std::vector<RsaWindow> &todo_refs, const std::vector<std::string> &sequences,
                              int match_score = 2, int mismatch_score = 8, int gap_open_score = 12, int gap_extend_score = 1);

// SURVEY 8(f) rank 3, caller half: the Hamming shortcut of extend_seed_part (src/aln.cpp:391-404) for a whole chunk in
// ONE blocking call.  Pair i = query i (qcat[qoff[i] .. qoff[i+1])) against the |query i| bases of contig ref_id[i] that
// start at start[i]; the engine reads them from the genome resident in HBM (rsa_ext_hamming_ref_windows).  On return
// out[i].status == 0: hamming_align's AlignmentInfo (src/aligner.cpp:254-302); 1: the pair needs the gapped path;
// 3: more CIGAR runs than the record holds (host path).  Called from integration/hamming_glue.cpp.
void solve_hamming_on_gpu_windows(int thread_id, size_t n, const char *qcat, const int64_t *qoff, const uint32_t *ref_id,
                                  const uint32_t *start, const std::vector<std::string> &sequences, int match_score,
                                  int mismatch_score, int gap_open_score, int gap_extend_score, int end_bonus,
                                  int32_t *hamming, rsa_ext_alninfo_t *out);
#endif

#ifdef RSA_EXT_ALNINFO
// The two call-site helpers of the optional build.  They replace, at the four caller loops of src/pc.cpp
// (:735-744 and siblings), `gasal_fail(q, r, rec)` and `aligner.align_gpu(q, r, rec)`; integration/patch_caller.py
// makes exactly that substitution on a build-time copy.  Anything the device did not settle (CIGAR longer than
// RSA_EXT_CIGAR_INLINE ops, another end bonus than the aligner's) takes the reference's own host code.
void rsa_ext_veneer_end_bonus(int end_bonus);  // tells the veneer the aligner's -L; from then on it skips CIGAR text
                                               // for records the device settled

template <class RefT, class Rec>
bool rsa_ext_gasal_fail(std::string &query, RefT &ref_in, Rec &rec) {
    if (rec.aln_end_bonus >= 0) {
        if (rec.aln.status == 0) return false;                         // accepted on the device
        if (rec.aln.status == 1 || rec.aln.status == 2) return true;   // gasal_fail / window over MAX_TARGET_LEN
    }
    std::string ref = ref_in;            // (a copy only on this rare path; RsaWindow materialises its string here)
    return gasal_fail(query, ref, rec);  // src/pc.cpp:466-478
}

template <class AlignerT, class RefT, class Rec>
auto rsa_ext_align_gpu(const AlignerT &aligner, const std::string &query, const RefT &ref_in, Rec &rec)
    -> decltype(aligner.align_gpu(query, std::string(), rec)) {
    const int want = aligner.parameters.end_bonus;
    rsa_ext_veneer_end_bonus(want);
    if (rec.aln_end_bonus == want && rec.aln.status == 0) {
        decltype(aligner.align_gpu(query, std::string(), rec)) info;  // AlignmentInfo, src/aligner.hpp:20-30
        info.cigar = decltype(info.cigar)(const_cast<uint32_t *>(rec.aln.cigar), (size_t)rec.aln.n_cigar);
        info.edit_distance = (unsigned)rec.aln.edit_distance;
        info.ref_start = (unsigned)rec.aln.ref_start;
        info.ref_end = (unsigned)rec.aln.ref_end;
        info.query_start = (unsigned)rec.aln.query_start;
        info.query_end = (unsigned)rec.aln.query_end;
        info.sw_score = rec.aln.sw_score;
        return info;
    }
    // host path for this record (long CIGAR, or a record computed before the veneer learned the aligner's -L).  The
    // veneer only omits the CIGAR text of records the device settled with the aligner's own bonus, so the text is here.
    if (rec.cigar_str.empty() && rec.score > 0 && rec.aln_end_bonus >= 0 && rec.aln.status == 0) {
        std::cerr << "[RSA_EXT ERROR:] device record without CIGAR text reached the host path (end bonus "
                  << rec.aln_end_bonus << " vs " << want << ")" << std::endl;
        exit(EXIT_FAILURE);
    }
    const std::string ref = ref_in;
    return aligner.align_gpu(query, ref, rec);
}
#endif  // RSA_EXT_ALNINFO
#endif  // STROBEALIGN_GASAL2_SSW_H
