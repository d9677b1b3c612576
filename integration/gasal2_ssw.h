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

struct gasal_tmp_res {
    int score;
    int query_start;
    int query_end;
    int ref_start;
    int ref_end;
    std::string cigar_str;
};

void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &todo_querys,
                      std::vector<std::string> &todo_refs, int match_score = 2, int mismatch_score = 8,
                      int gap_open_score = 12, int gap_extend_score = 1);
#endif  // STROBEALIGN_GASAL2_SSW_H
