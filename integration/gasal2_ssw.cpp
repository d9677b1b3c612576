// integration/gasal2_ssw.cpp -- solve_ssw_on_gpu on top of the C ABI (include/rsa_ext.h).
//
// Behaviour kept from the reference's src/gasal2_ssw.cpp:19-256:
//   * one GPU context per worker `thread_id` (< THREAD_NUM_MAX), created on the first call with that call's
//     scores (:29-57,92-102) and never re-configured afterwards;
//   * the result vector is resized to the batch and overwritten (:33);
//   * a query longer than MAX_QUERY_LEN prints the same message and exit(0)s (:84-87); a CUDA failure prints
//     and exit(EXIT_FAILURE)s (GASAL2/src/gasal.h:15-22);
//   * blocking call; CIGAR text as the reference prints it (:184-243).
// Differences: any batch size, storage is released at process exit, and with several GPUs visible the
// workers are spread over them (thread_id % device count; RSA_EXT_DEVICES=n caps the count) -- the
// reference uses device 0 only (:34).  CUDA initialisation and the primary context are started from a static
// initializer on a helper thread, so they overlap the index load instead of stalling every worker's first batch
// (measured 1-3 s inside a busy 16-worker process, tools/pipe_trace.sh); RSA_EXT_NO_WARMUP=1 turns that off.
#include "gasal2_ssw.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <utility>

#include "rsa_ext.h"

namespace {

struct Worker {
    rsa_ext_t *h = nullptr;
    std::vector<const char *> qp, tp;
    std::vector<int32_t> ql, tl;
    std::vector<rsa_ext_result_t> res;
#ifdef RSA_EXT_ALNINFO
    std::vector<rsa_ext_alninfo_t> aln;
#endif
    std::vector<uint8_t> rle;
    ~Worker() { if (h) rsa_ext_destroy(h); }
};

Worker g_workers[THREAD_NUM_MAX];
std::mutex g_create_mutex;

[[noreturn]] void die(const char *what, rsa_ext_t *h) {
    fprintf(stderr, "[RSA_EXT ERROR:] %s: %s\n", what, rsa_ext_last_error(h));
    exit(EXIT_FAILURE);
}

int device_count_cap() {
    const char *e = getenv("RSA_EXT_DEVICES");
    return e ? atoi(e) : 0;
}

#ifdef RSA_EXT_ALNINFO
// End bonus the finish kernel uses.  Starts at strobealign's default; the first call site that runs tells us the
// aligner's value (rsa_ext_veneer_end_bonus), after which records the device settled carry no CIGAR text.
// Bonus and "confirmed" flag live in ONE atomic word (bit 30 = confirmed), so a worker can never pair a stale bonus
// with confirmed == true.
constexpr int kBonusConfirmed = 1 << 30;
std::atomic<int> g_end_bonus_state{10};
#endif

int usable_devices() {
    int ndev = rsa_ext_device_count();
    const int cap = device_count_cap();
    return (cap > 0 && cap < ndev) ? cap : ndev;
}

// Start CUDA early and keep a pool of ready handles: the context, the streams and every buffer a 512-pair slice
// needs are created by ONE helper thread while the host loads the index, not by 16 workers inside their first
// batches (allocations issued while other workers are enqueueing stalled everyone for 0.05-0.9 s each).
// The pool uses strobealign's default scores; a first call with other scores, or more workers than the pool holds,
// creates its handle the ordinary way.  A failure here is not reported -- the first real call reports it.
constexpr int kDefaultScores[4] = {2, 8, 12, 1};  // src/cmdline.hpp:46-50 / the prototype's default arguments
constexpr int kPoolReadLen = 250, kPoolWindowLen = 500;  // shapes the pooled handles are pre-sized for

struct Warmup {
    std::thread t;
    std::mutex m;
    std::condition_variable cv;
    bool done = true;
    std::vector<std::pair<int, rsa_ext_t *>> pool;  // (device, handle)

    Warmup() {
        if (getenv("RSA_EXT_NO_WARMUP")) return;
        done = false;
        // RSA_EXT_WARMUP=sync: create the CUDA context(s) right here, before main() starts its threads (context
        // creation inside a process that is already building an index on 16 threads was measured 2-4x slower)
        const char *mode = getenv("RSA_EXT_WARMUP");
        if (mode && !strcmp(mode, "sync")) {
            const int nd = usable_devices();
            for (int d = 0; d < nd; ++d) {
                rsa_ext_config_t cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.device = d;
                rsa_ext_t *h = nullptr;
                if (rsa_ext_create(&cfg, &h) == RSA_EXT_OK) rsa_ext_destroy(h);
            }
        }
        t = std::thread([this] {
            const int ndev = usable_devices();
            const char *pe = getenv("RSA_EXT_POOL");
            const int hw = (int)std::thread::hardware_concurrency();
            const int want = pe ? atoi(pe) : std::min(std::max(hw, 1), 32);
            for (int k = 0; k < want && ndev > 0; ++k) {
                rsa_ext_config_t cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.device = k % ndev;
                cfg.max_query_len = MAX_QUERY_LEN;
                cfg.max_target_len = MAX_TARGET_LEN;
                cfg.match = kDefaultScores[0]; cfg.mismatch = kDefaultScores[1];
                cfg.gap_open = kDefaultScores[2]; cfg.gap_extend = kDefaultScores[3];
                rsa_ext_t *h = nullptr;
                if (rsa_ext_create(&cfg, &h) != RSA_EXT_OK) break;
                rsa_ext_reserve(h, STREAM_BATCH_SIZE, kPoolReadLen, kPoolWindowLen);
                if (k < ndev) {  // one tiny batch per device loads the common kernels
                    const std::string q(150, 'A'), w(200, 'A');
                    const char *qp = q.data(), *tp = w.data();
                    const int32_t ql = (int32_t)q.size(), tl = (int32_t)w.size();
                    rsa_ext_result_t r;
                    if (rsa_ext_submit_ptrs(h, 1, &qp, &ql, &tp, &tl, &r) == RSA_EXT_OK) rsa_ext_wait(h);
                }
                std::lock_guard<std::mutex> lk(m);
                pool.emplace_back(cfg.device, h);
            }
            std::lock_guard<std::mutex> lk(m);
            done = true;
            cv.notify_all();
        });
    }
    // a ready handle on `device` with the default scores, or nullptr
    rsa_ext_t *take(int device, int match, int mismatch, int gap_open, int gap_extend) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return done; });
        if (match != kDefaultScores[0] || mismatch != kDefaultScores[1] || gap_open != kDefaultScores[2] ||
            gap_extend != kDefaultScores[3])
            return nullptr;
        for (size_t i = 0; i < pool.size(); ++i)
            if (pool[i].first == device) {
                rsa_ext_t *h = pool[i].second;
                pool.erase(pool.begin() + (long)i);
                return h;
            }
        return nullptr;
    }
    ~Warmup() {
        if (t.joinable()) t.join();
        for (auto &e : pool) rsa_ext_destroy(e.second);
    }
} g_warmup;

}  // namespace

void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &query_seqs,
                      std::vector<std::string> &target_seqs, int match_score, int mismatch_score, int gap_open_score,
                      int gap_extend_score) {
    // the reference only assert()s these (src/gasal2_ssw.cpp:27-28), which vanishes under NDEBUG and then indexes its
    // per-thread statics out of bounds; fail loudly instead
    if (thread_id < 0 || thread_id >= THREAD_NUM_MAX || query_seqs.size() != target_seqs.size()) {
        fprintf(stderr, "[RSA_EXT ERROR:] solve_ssw_on_gpu: thread_id %d outside [0, %d) or %zu queries vs %zu windows\n",
                thread_id, THREAD_NUM_MAX, query_seqs.size(), target_seqs.size());
        exit(EXIT_FAILURE);
    }
    Worker &w = g_workers[thread_id];
    const size_t n = query_seqs.size();
    gasal_results.resize(n);
    if (n == 0) return;

    if (!w.h) {
        std::lock_guard<std::mutex> lock(g_create_mutex);
        const int ndev = usable_devices();
        rsa_ext_config_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.device = ndev > 0 ? thread_id % ndev : 0;
        cfg.max_query_len = MAX_QUERY_LEN;
        cfg.max_target_len = MAX_TARGET_LEN;
        cfg.match = match_score;
        cfg.mismatch = mismatch_score;
        cfg.gap_open = gap_open_score;
        cfg.gap_extend = gap_extend_score;
        w.h = g_warmup.take(cfg.device, match_score, mismatch_score, gap_open_score, gap_extend_score);
        if (!w.h && rsa_ext_create(&cfg, &w.h) != RSA_EXT_OK) die("rsa_ext_create", nullptr);
    }

    w.qp.resize(n); w.tp.resize(n); w.ql.resize(n); w.tl.resize(n); w.res.resize(n);
    for (size_t i = 0; i < n; ++i) {
        w.qp[i] = query_seqs[i].data(); w.ql[i] = (int32_t)query_seqs[i].size();
        w.tp[i] = target_seqs[i].data(); w.tl[i] = (int32_t)target_seqs[i].size();
    }
#ifdef RSA_EXT_ALNINFO
    const int bonus_state = g_end_bonus_state.load(std::memory_order_acquire);
    const int end_bonus = bonus_state & (kBonusConfirmed - 1);
    const bool text_free = (bonus_state & kBonusConfirmed) != 0;  // end_bonus is the aligner's own value
    w.aln.resize(n);
    if (rsa_ext_request_alninfo(w.h, w.aln.data(), end_bonus) != RSA_EXT_OK) die("rsa_ext_request_alninfo", w.h);
#endif
    int rc = rsa_ext_submit_ptrs(w.h, (int64_t)n, w.qp.data(), w.ql.data(), w.tp.data(), w.tl.data(), w.res.data());
    if (rc == RSA_EXT_ERR_QUERY_LEN) {
        size_t mx = 0;
        for (size_t i = 0; i < n; ++i) mx = MAX(mx, query_seqs[i].length());
        std::cerr << "gasal2 : read size is too big, " << mx << " > " << MAX_QUERY_LEN << std::endl;
        exit(0);
    }
    if (rc != RSA_EXT_OK) die("rsa_ext_submit_ptrs", w.h);
    if (rsa_ext_wait(w.h) != RSA_EXT_OK) die("rsa_ext_wait", w.h);

    char text[4096];
    for (size_t i = 0; i < n; ++i) {
        const rsa_ext_result_t &r = w.res[i];
        const uint8_t *rle = r.rle;
        if (r.n_ops > RSA_EXT_RLE_INLINE) {
            w.rle.resize((size_t)r.n_ops);
            if (rsa_ext_rle_overflow(w.h, (int64_t)i, w.rle.data(), r.n_ops) != r.n_ops) die("rsa_ext_rle_overflow", w.h);
            rle = w.rle.data();
        }
        std::string cigar;
#ifdef RSA_EXT_ALNINFO
        const bool need_text = !text_free || w.aln[i].status == 3;
#else
        const bool need_text = true;
#endif
        if (r.n_ops > 0 && need_text) {
            int len = rsa_ext_rle_to_text(rle, r.n_ops, text, (int32_t)sizeof text);
            if (len < 0) {
                std::vector<char> big((size_t)r.n_ops * 8 + 16);
                len = rsa_ext_rle_to_text(rle, r.n_ops, big.data(), (int32_t)big.size());
                cigar.assign(big.data(), (size_t)(len > 0 ? len : 0));
            } else {
                cigar.assign(text, (size_t)len);
            }
        }
        gasal_results[i] = {r.score, r.query_start, r.query_end, r.ref_start, r.ref_end, std::move(cigar)};
#ifdef RSA_EXT_ALNINFO
        gasal_results[i].aln = w.aln[i];
        gasal_results[i].aln_end_bonus = end_bonus;
#endif
    }
}

#ifdef RSA_EXT_ALNINFO
void rsa_ext_veneer_end_bonus(int end_bonus) {
    const int want = (end_bonus & (kBonusConfirmed - 1)) | kBonusConfirmed;
    if (g_end_bonus_state.load(std::memory_order_relaxed) == want) return;
    g_end_bonus_state.store(want, std::memory_order_release);
}
#endif
