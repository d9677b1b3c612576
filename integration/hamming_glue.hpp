// integration/hamming_glue.hpp -- binding of the device-side Hamming shortcut (include/rsa_ext.h: rsa_ext_hamming_ref_windows)
// into the reference's host pipeline (the caller half).
//
// The reference decides the shortcut candidate by candidate on the worker thread, inside extend_seed_part
// (src/aln.cpp:374-431): substr of the projected window, hamming_distance, the 5 % test, hamming_align.  In the pipeline's
// two-phase form (align_SE_part / align_PE_part fill an AlignTmpRes, src/pc.cpp extends the todo list later) nothing that
// follows in phase one looks at the outcome, with one exception: a "good pair" whose two mates both took the shortcut
// feeds the insert-size estimator while it has fewer than 400 samples (src/aln.cpp:1450-1466).  So:
//
//   * extend_seed_part (patched by integration/patch_hamming.py) leaves an eligible candidate PENDING -- todo entry with
//     done_align = false and align_res.ref_id = kHammingPending, align_res.ref_start = the projected window start -- when
//     deferral is on: always for single-end reads, for paired-end reads once the estimator has its 400 samples (until then
//     the reference's host code runs unchanged, so the estimator sees exactly the same updates);
//   * before src/pc.cpp builds a chunk's todo list (its "step1" loops, :610, :905, :1219, :1613) hamming_pass_se / _pe send
//     all pending candidates of the chunk down in ONE call; a candidate that passes becomes the done entry the reference
//     would have stored (same Alignment fields), one that fails stays in the todo list and takes the gapped path.
#ifndef RSA_HAMMING_GLUE_HPP
#define RSA_HAMMING_GLUE_HPP
#include <string>
#include <vector>

#include "aligner.hpp"
#include "nam.hpp"
#include "refs.hpp"
#include "sam.hpp"

namespace rsa_glue {

constexpr int kHammingPending = -0x48414D;  // align_res[j].ref_id of a candidate whose shortcut is still to be decided
constexpr size_t kHammingMaxQuery = 500;    // MAX_QUERY_LEN of the boundary; longer reads keep the host path

// Scope guard around the body of align_SE_part / align_PE_part: is deferral on for the candidates of this read (pair)?
struct HammingDefer {
    explicit HammingDefer(bool on);
    ~HammingDefer();
    bool prev;
};
bool hamming_deferred();

// What extend_seed_part pushes for a deferred candidate (the same four vectors it fills on its other paths).
inline void push_pending(AlignTmpRes& align_tmp_res, const Nam& nam, int projected_ref_start) {
    align_tmp_res.todo_nams.push_back(nam);
    align_tmp_res.is_extend_seed.push_back(true);
    align_tmp_res.done_align.push_back(false);
    Alignment pending;
    pending.ref_id = kHammingPending;
    pending.ref_start = projected_ref_start;
    align_tmp_res.align_res.push_back(pending);
}

// seq2 == nullptr: single-end.  The query of entry j is (is_read1[j] ? seq1 : seq2)[i] or its reverse complement
// (nam.is_rc), exactly the string extend_seed_part compares on the host.
void hamming_pass(int thread_id, const std::vector<const std::string*>& seq1, const std::vector<const std::string*>* seq2,
                  std::vector<AlignTmpRes>& chunk, const References& references, const Aligner& aligner);

template <class Rec>
void hamming_pass_se(int thread_id, const std::vector<Rec>& records, std::vector<AlignTmpRes>& chunk, const References& references,
                     const Aligner& aligner) {
    std::vector<const std::string*> seqs;
    seqs.reserve(records.size());
    for (const Rec& r : records) seqs.push_back(&r.seq);
    hamming_pass(thread_id, seqs, nullptr, chunk, references, aligner);
}

template <class Rec>
void hamming_pass_pe(int thread_id, const std::vector<Rec>& records1, const std::vector<Rec>& records2, std::vector<AlignTmpRes>& chunk,
                     const References& references, const Aligner& aligner) {
    std::vector<const std::string*> s1, s2;
    s1.reserve(records1.size());
    s2.reserve(records2.size());
    for (const Rec& r : records1) s1.push_back(&r.seq);
    for (const Rec& r : records2) s2.push_back(&r.seq);
    hamming_pass(thread_id, s1, &s2, chunk, references, aligner);
}

}  // namespace rsa_glue
#endif
