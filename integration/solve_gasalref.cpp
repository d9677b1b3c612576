// integration/solve_gasalref.cpp -- TEST INFRASTRUCTURE: solve_ssw_on_gpu backed by the reference's own
// GASAL2 kernels compiled for the host (oracle/ref_shim.cpp -> libgasal_ref512.so).  Linked only into
// integration/_build/rabbitsalign_gasalref, the golden-SAM generator; never into the product.
#include "gasal2_ssw.h"
#include <cstdint>
#include <cstdlib>
#include <cstring>

extern "C" int gasal_ref_batch(int n, const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                               int match, int mismatch, int gap_open_m1, int gap_ext, int32_t *score, int32_t *qs,
                               int32_t *qe, int32_t *rs, int32_t *re, int32_t *n_ops_out, char *cigar_pool,
                               int64_t pool_cap, int64_t *cigar_off);

void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &query_seqs,
                      std::vector<std::string> &target_seqs, int match_score, int mismatch_score, int gap_open_score,
                      int gap_extend_score) {
    (void)thread_id;
    const int n = (int)query_seqs.size();
    gasal_results.resize(n);
    if (n == 0) return;
    std::string qb, tb;
    std::vector<int64_t> qo(n + 1), to(n + 1);
    for (int i = 0; i < n; ++i) {
        if (query_seqs[i].size() > MAX_QUERY_LEN) {
            std::cerr << "gasal2 : read size is too big, " << query_seqs[i].size() << " > " << MAX_QUERY_LEN << std::endl;
            exit(0);
        }
        qo[i] = (int64_t)qb.size(); to[i] = (int64_t)tb.size();
        qb += query_seqs[i]; tb += target_seqs[i];
    }
    qo[n] = (int64_t)qb.size(); to[n] = (int64_t)tb.size();
    std::vector<int32_t> sc(n), qs(n), qe(n), rs(n), re(n), no(n);
    std::vector<int64_t> co(n + 1);
    std::vector<char> pool(4 * (qb.size() + tb.size()) + 64 * (size_t)n + 64);
    if (gasal_ref_batch(n, qb.data(), qo.data(), tb.data(), to.data(), match_score, mismatch_score, gap_open_score - 1,
                        gap_extend_score, sc.data(), qs.data(), qe.data(), rs.data(), re.data(), no.data(), pool.data(),
                        (int64_t)pool.size(), co.data()) != 0) {
        std::cerr << "gasal_ref_batch failed" << std::endl;
        exit(EXIT_FAILURE);
    }
    for (int i = 0; i < n; ++i)
        gasal_results[i] = {sc[i], qs[i], qe[i], rs[i], re[i], std::string(pool.data() + co[i], pool.data() + co[i + 1])};
}
