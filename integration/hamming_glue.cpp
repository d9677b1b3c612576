// integration/hamming_glue.cpp -- see hamming_glue.hpp.  Compiled with -DRSA_EXT_WINDOWS (the genome is resident in HBM,
// windows travel as offsets).
#include "hamming_glue.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gasal2_ssw.h"
#include "revcomp.hpp"

namespace rsa_glue {

namespace {
thread_local bool t_defer = false;

// per worker thread: the chunk's candidates in todo order
struct Scratch {
    std::vector<char> qcat;
    std::vector<int64_t> qoff;
    std::vector<uint32_t> ref_id, start;
    std::vector<int32_t> dist;
    std::vector<rsa_ext_alninfo_t> info;
    struct Where { uint32_t read, entry; };
    std::vector<Where> where;
};
thread_local Scratch t_scratch;
}  // namespace

HammingDefer::HammingDefer(bool on) : prev(t_defer) {
    static const bool host_only = getenv("RSA_EXT_HOST_HAMMING") != nullptr;   // (once per process: this runs per read)
    t_defer = on && !host_only;
}
HammingDefer::~HammingDefer() { t_defer = prev; }
bool hamming_deferred() { return t_defer; }

void hamming_pass(int thread_id, const std::vector<const std::string*>& seq1, const std::vector<const std::string*>* seq2,
                  std::vector<AlignTmpRes>& chunk, const References& references, const Aligner& aligner) {
    Scratch& s = t_scratch;
    s.qcat.clear(); s.qoff.clear(); s.ref_id.clear(); s.start.clear(); s.where.clear();
    std::string rc[2];
    for (size_t i = 0; i < chunk.size(); ++i) {
        AlignTmpRes& r = chunk[i];
        bool have_rc[2] = {false, false};
        for (size_t j = 0; j < r.todo_nams.size(); ++j) {
            if (r.done_align[j] || !r.is_extend_seed[j] || r.align_res[j].ref_id != kHammingPending) continue;
            const int mate = r.is_read1[j] ? 0 : 1;
            const std::string& seq = mate == 0 ? *seq1[i] : *(*seq2)[i];
            const std::string* query = &seq;
            if (r.todo_nams[j].is_rc) {
                if (!have_rc[mate]) { rc[mate] = reverse_complement(seq); have_rc[mate] = true; }
                query = &rc[mate];
            }
            s.qoff.push_back((int64_t)s.qcat.size());
            s.qcat.insert(s.qcat.end(), query->begin(), query->end());
            s.ref_id.push_back((uint32_t)r.todo_nams[j].ref_id);
            s.start.push_back((uint32_t)r.align_res[j].ref_start);
            s.where.push_back({(uint32_t)i, (uint32_t)j});
        }
    }
    const size_t n = s.where.size();
    if (n == 0) return;
    s.qoff.push_back((int64_t)s.qcat.size());
    s.qcat.resize(s.qcat.size() + 16);
    s.dist.resize(n);
    s.info.resize(n);
    const AlignmentParameters& p = aligner.parameters;
    solve_hamming_on_gpu_windows(thread_id, n, s.qcat.data(), s.qoff.data(), s.ref_id.data(), s.start.data(), references.sequences,
                                 p.match, p.mismatch, p.gap_open, p.gap_extend, p.end_bonus, s.dist.data(), s.info.data());
    for (size_t k = 0; k < n; ++k) {
        AlignTmpRes& r = chunk[s.where[k].read];
        const size_t j = s.where[k].entry;
        Alignment& alignment = r.align_res[j];
        const Nam& nam = r.todo_nams[j];
        const int projected_ref_start = alignment.ref_start;
        const size_t qlen = (size_t)(s.qoff[k + 1] - s.qoff[k]);
        const rsa_ext_alninfo_t& d = s.info[k];
        if (d.status == 1) {  // the 5 % test failed: the candidate stays in the todo list (src/aln.cpp:408-412)
            alignment = Alignment();
            continue;
        }
        AlignmentInfo info;
        if (d.status == 0) {
            info.cigar = Cigar(const_cast<uint32_t*>(d.cigar), (size_t)d.n_cigar);
            info.edit_distance = (unsigned)d.edit_distance;
            info.ref_start = (unsigned)d.ref_start;
            info.ref_end = (unsigned)d.ref_end;
            info.query_start = (unsigned)d.query_start;
            info.query_end = (unsigned)d.query_end;
            info.sw_score = d.sw_score;
        } else if (d.status == 3) {  // more =/X runs than the record holds: the reference's own function
            const std::string query(s.qcat.data() + s.qoff[k], qlen);
            const std::string ref_segm_ham = references.sequences[nam.ref_id].substr((size_t)projected_ref_start, qlen);
            info = hamming_align(query, ref_segm_ham, p.match, p.mismatch, p.end_bonus);
        } else {
            fprintf(stderr, "[RSA_EXT ERROR:] Hamming shortcut: unexpected status %d\n", (int)d.status);
            exit(EXIT_FAILURE);
        }
        // src/aln.cpp:413-429
        const int softclipped = info.query_start + (qlen - info.query_end);
        alignment.cigar = std::move(info.cigar);
        alignment.edit_distance = info.edit_distance;
        alignment.global_ed = info.edit_distance + softclipped;
        alignment.score = info.sw_score;
        alignment.ref_start = projected_ref_start + info.ref_start;
        alignment.length = info.ref_span();
        alignment.is_rc = nam.is_rc;
        alignment.is_unaligned = false;
        alignment.ref_id = nam.ref_id;
        alignment.gapped = false;
        r.done_align[j] = true;
    }
}

}  // namespace rsa_glue
