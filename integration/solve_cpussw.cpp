// integration/solve_cpussw.cpp -- TEST/BENCH INFRASTRUCTURE: a solve_ssw_on_gpu that reports every pair as
// failed, so the caller's gate (reference src/pc.cpp:466-478, 735-744) sends every extension to the
// reference's CPU SSW path Aligner::align.  Linked only into integration/_build/rabbitsalign_cpussw.
#include "gasal2_ssw.h"

void solve_ssw_on_gpu(int, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &query_seqs,
                      std::vector<std::string> &, int, int, int, int) {
    gasal_results.assign(query_seqs.size(), gasal_tmp_res{0, -1, -1, -1, -1, ""});
}
