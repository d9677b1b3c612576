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
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <utility>

#include "rsa_ext.h"

namespace {

struct Worker {
    rsa_ext_t *h = nullptr;
    int device = 0;
    bool has_reference = false;                 // window build: the handle sees the genome resident on its GPU
    std::vector<char> qcat;                     // window build: the batch's queries back to back
    std::vector<int64_t> qoff, woff;
    std::vector<int32_t> wlen;
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

// RSA_EXT_STATS=1: where the time of the calls went, summed over the workers, printed at exit (diagnostics)
struct VeneerStats {
    std::atomic<long long> calls{0}, pairs{0}, ns_gather{0}, ns_submit{0}, ns_wait{0}, ns_unpack{0}, ns_max_call{0};
    std::atomic<long long> ham_calls{0}, ham_pairs{0}, ns_hamming{0}, ns_max_ham{0}, ns_attach{0}, ns_acquire{0};
    const bool on = getenv("RSA_EXT_STATS") != nullptr;
    ~VeneerStats() {
        if (!on || !calls.load()) return;
        fprintf(stderr, "[rsa_ext veneer] %lld calls, %lld pairs; summed over workers: gather %.1f ms, submit %.1f ms, wait %.1f ms, "
                        "unpack %.1f ms; longest call %.1f ms\n", calls.load(), pairs.load(), ns_gather.load() / 1e6, ns_submit.load() / 1e6,
                ns_wait.load() / 1e6, ns_unpack.load() / 1e6, ns_max_call.load() / 1e6);
        if (ham_calls.load())
            fprintf(stderr, "[rsa_ext veneer] Hamming shortcut: %lld calls, %lld pairs, %.1f ms summed over workers, longest call %.1f ms\n",
                    ham_calls.load(), ham_pairs.load(), ns_hamming.load() / 1e6, ns_max_ham.load() / 1e6);
        fprintf(stderr, "[rsa_ext veneer] first calls waiting for a pooled handle: %.1f ms summed over workers\n", ns_acquire.load() / 1e6);
        if (ns_attach.load())
            fprintf(stderr, "[rsa_ext veneer] waiting for / uploading the resident genome: %.1f ms summed over workers\n", ns_attach.load() / 1e6);
    }
} g_vstats;
inline long long now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
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
// (whole-chunk builds raise the macro: a chunk of 10 000 reads sends 5-20 k pairs; a pool sized for 8192 made every worker
// regrow its pinned staging and device buffers inside its first calls -- cudaHostAlloc / cudaMalloc under the driver lock,
// 0.1-0.2 s each with 16 workers, tools/r2_call30.sh)
constexpr int kPoolPairs = STREAM_BATCH_SIZE < 8192 ? STREAM_BATCH_SIZE : 24576;

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
                rsa_ext_reserve(h, kPoolPairs, kPoolReadLen, kPoolWindowLen);
                if (k < ndev) {  // one tiny batch per device loads the common kernels
                    const std::string q(150, 'A'), w(200, 'A');
                    const char *qp = q.data(), *tp = w.data();
                    const int32_t ql = (int32_t)q.size(), tl = (int32_t)w.size();
                    rsa_ext_result_t r;
                    if (rsa_ext_submit_ptrs(h, 1, &qp, &ql, &tp, &tl, &r) == RSA_EXT_OK) rsa_ext_wait(h);
#ifdef RSA_EXT_WINDOWS
                    const int64_t off2[2] = {0, ql};   // ... and the Hamming kernel (gpuham builds)
                    int32_t dist;
                    rsa_ext_alninfo_t info;
                    rsa_ext_hamming_align(h, 1, qp, off2, qp, off2, 10, &dist, &info);
#endif
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

namespace {

// The worker's handle: created on its first call with that call's scores (src/gasal2_ssw.cpp:29-57,92-102).
Worker &acquire_worker(int thread_id, size_t n_queries, size_t n_targets, int match_score, int mismatch_score,
                       int gap_open_score, int gap_extend_score) {
    // the reference only assert()s these (src/gasal2_ssw.cpp:27-28), which vanishes under NDEBUG and then indexes its
    // per-thread statics out of bounds; fail loudly instead
    if (thread_id < 0 || thread_id >= THREAD_NUM_MAX || n_queries != n_targets) {
        fprintf(stderr, "[RSA_EXT ERROR:] solve_ssw_on_gpu: thread_id %d outside [0, %d) or %zu queries vs %zu windows\n",
                thread_id, THREAD_NUM_MAX, n_queries, n_targets);
        exit(EXIT_FAILURE);
    }
    Worker &w = g_workers[thread_id];
    if (!w.h) {
        const long long t_acq = now_ns();
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
        w.device = cfg.device;
        w.h = g_warmup.take(cfg.device, match_score, mismatch_score, gap_open_score, gap_extend_score);
        if (!w.h && rsa_ext_create(&cfg, &w.h) != RSA_EXT_OK) die("rsa_ext_create", nullptr);
        if (g_vstats.on) g_vstats.ns_acquire += now_ns() - t_acq;
    }
    return w;
}

[[noreturn]] void die_query_too_long(const std::vector<std::string> &query_seqs) {
    size_t mx = 0;
    for (const std::string &q : query_seqs) mx = MAX(mx, q.length());
    std::cerr << "gasal2 : read size is too big, " << mx << " > " << MAX_QUERY_LEN << std::endl;  // gasal2_ssw.cpp:84-87
    exit(0);
}

// records -> gasal_tmp_res (CIGAR text as src/gasal2_ssw.cpp:184-243 prints it)
void unpack_results(Worker &w, size_t n, std::vector<gasal_tmp_res> &gasal_results, bool text_free, int end_bonus) {
    (void)text_free; (void)end_bonus;
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

}  // namespace

int rsa_ext_veneer_device(int thread_id) {
    const int ndev = usable_devices();
    return ndev > 0 && thread_id >= 0 ? thread_id % ndev : 0;
}

void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &query_seqs,
                      std::vector<std::string> &target_seqs, int match_score, int mismatch_score, int gap_open_score,
                      int gap_extend_score) {
    Worker &w = acquire_worker(thread_id, query_seqs.size(), target_seqs.size(), match_score, mismatch_score,
                               gap_open_score, gap_extend_score);
    const size_t n = query_seqs.size();
    gasal_results.resize(n);
    if (n == 0) return;

    const long long t0 = now_ns();
    w.qp.resize(n); w.tp.resize(n); w.ql.resize(n); w.tl.resize(n); w.res.resize(n);
    for (size_t i = 0; i < n; ++i) {
        w.qp[i] = query_seqs[i].data(); w.ql[i] = (int32_t)query_seqs[i].size();
        w.tp[i] = target_seqs[i].data(); w.tl[i] = (int32_t)target_seqs[i].size();
    }
    int end_bonus = 0;
    bool text_free = false;
#ifdef RSA_EXT_ALNINFO
    const int bonus_state = g_end_bonus_state.load(std::memory_order_acquire);
    end_bonus = bonus_state & (kBonusConfirmed - 1);
    text_free = (bonus_state & kBonusConfirmed) != 0;  // end_bonus is the aligner's own value
    w.aln.resize(n);
    if (rsa_ext_request_alninfo(w.h, w.aln.data(), end_bonus) != RSA_EXT_OK) die("rsa_ext_request_alninfo", w.h);
#endif
    const long long t1 = now_ns();
    int rc = rsa_ext_submit_ptrs(w.h, (int64_t)n, w.qp.data(), w.ql.data(), w.tp.data(), w.tl.data(), w.res.data());
    if (rc == RSA_EXT_ERR_QUERY_LEN) die_query_too_long(query_seqs);
    if (rc != RSA_EXT_OK) die("rsa_ext_submit_ptrs", w.h);
    const long long t2 = now_ns();
    rc = rsa_ext_wait(w.h);
    if (rc == RSA_EXT_ERR_QUERY_LEN) die_query_too_long(query_seqs);  // (large batches are validated chunk by chunk)
    if (rc != RSA_EXT_OK) die("rsa_ext_wait", w.h);
    const long long t3 = now_ns();
    unpack_results(w, n, gasal_results, text_free, end_bonus);
    if (g_vstats.on) {
        const long long t4 = now_ns();
        g_vstats.calls++; g_vstats.pairs += (long long)n;
        g_vstats.ns_gather += t1 - t0; g_vstats.ns_submit += t2 - t1; g_vstats.ns_wait += t3 - t2; g_vstats.ns_unpack += t4 - t3;
        long long mx = g_vstats.ns_max_call.load();
        while (t4 - t0 > mx && !g_vstats.ns_max_call.compare_exchange_weak(mx, t4 - t0)) {}
    }
}

#ifdef RSA_EXT_WINDOWS
namespace {
// The reference genome as ONE buffer (contigs back to back) + where each contig starts: built once per process from
// references.sequences, uploaded once per GPU by the first worker that lands there and shared by the others.
struct ResidentGenome {
    std::mutex m;
    const std::vector<std::string> *source = nullptr;
    std::string concat;
    std::vector<int64_t> contig_off;
    rsa_ext_t *owner[64] = {};  // per device: the handle that uploaded (donor for rsa_ext_share_reference)
} g_genome;

void attach_reference(Worker &w, const std::vector<std::string> &sequences) {
    std::lock_guard<std::mutex> lk(g_genome.m);
    if (!g_genome.source) {
        size_t total = 0;
        for (const std::string &s : sequences) total += s.size();
        g_genome.concat.reserve(total);
        g_genome.contig_off.reserve(sequences.size() + 1);
        for (const std::string &s : sequences) { g_genome.contig_off.push_back((int64_t)g_genome.concat.size()); g_genome.concat += s; }
        g_genome.contig_off.push_back((int64_t)g_genome.concat.size());
        g_genome.source = &sequences;
    } else if (g_genome.source != &sequences) {
        fprintf(stderr, "[RSA_EXT ERROR:] solve_ssw_on_gpu_windows: one reference per process\n");
        exit(EXIT_FAILURE);
    }
    const int d = w.device;
    if (d < 0 || d >= 64) die("device ordinal", nullptr);
    if (!g_genome.owner[d]) {
        if (rsa_ext_set_reference(w.h, g_genome.concat.data(), (int64_t)g_genome.concat.size()) != RSA_EXT_OK) die("rsa_ext_set_reference", w.h);
        g_genome.owner[d] = w.h;
    } else if (rsa_ext_share_reference(w.h, g_genome.owner[d]) != RSA_EXT_OK) {
        die("rsa_ext_share_reference", w.h);
    }
    w.has_reference = true;
}
}  // namespace

void solve_ssw_on_gpu_windows(int thread_id, std::vector<gasal_tmp_res> &gasal_results, std::vector<std::string> &query_seqs,
                              std::vector<RsaWindow> &windows, const std::vector<std::string> &sequences, int match_score,
                              int mismatch_score, int gap_open_score, int gap_extend_score) {
    Worker &w = acquire_worker(thread_id, query_seqs.size(), windows.size(), match_score, mismatch_score, gap_open_score,
                               gap_extend_score);
    const size_t n = query_seqs.size();
    gasal_results.resize(n);
    if (n == 0) return;
    const long long t0 = now_ns();
    if (!w.has_reference) attach_reference(w, sequences);
    const long long t_attach = now_ns();

    // queries back to back (the engine copies them to the GPU from here); windows as (offset, length) in the genome
    w.qoff.resize(n + 1); w.woff.resize(n); w.wlen.resize(n); w.res.resize(n);
    size_t qbytes = 0;
    for (size_t i = 0; i < n; ++i) { w.qoff[i] = (int64_t)qbytes; qbytes += query_seqs[i].size(); }
    w.qoff[n] = (int64_t)qbytes;
    w.qcat.resize(qbytes + 16);
    for (size_t i = 0; i < n; ++i) {
        memcpy(w.qcat.data() + w.qoff[i], query_seqs[i].data(), query_seqs[i].size());
        w.woff[i] = g_genome.contig_off[windows[i].ref_id] + (int64_t)windows[i].start;
        w.wlen[i] = (int32_t)windows[i].len;
    }
    const int bonus_state = g_end_bonus_state.load(std::memory_order_acquire);
    const int end_bonus = bonus_state & (kBonusConfirmed - 1);
    const bool text_free = (bonus_state & kBonusConfirmed) != 0;
    w.aln.resize(n);
    if (rsa_ext_request_alninfo(w.h, w.aln.data(), end_bonus) != RSA_EXT_OK) die("rsa_ext_request_alninfo", w.h);
    const long long t1 = now_ns();
    int rc = rsa_ext_submit_ref_windows(w.h, (int64_t)n, w.qcat.data(), w.qoff.data(), w.woff.data(), w.wlen.data(), w.res.data());
    if (rc == RSA_EXT_ERR_QUERY_LEN) die_query_too_long(query_seqs);
    if (rc != RSA_EXT_OK) die("rsa_ext_submit_ref_windows", w.h);
    const long long t2 = now_ns();
    rc = rsa_ext_wait(w.h);
    if (rc == RSA_EXT_ERR_QUERY_LEN) die_query_too_long(query_seqs);
    if (rc != RSA_EXT_OK) die("rsa_ext_wait", w.h);
    const long long t3 = now_ns();
    unpack_results(w, n, gasal_results, text_free, end_bonus);
    if (g_vstats.on) {
        const long long t4 = now_ns();
        g_vstats.calls++; g_vstats.pairs += (long long)n; g_vstats.ns_attach += t_attach - t0;
        g_vstats.ns_gather += t1 - t_attach; g_vstats.ns_submit += t2 - t1; g_vstats.ns_wait += t3 - t2; g_vstats.ns_unpack += t4 - t3;
        long long mx = g_vstats.ns_max_call.load();
        while (t4 - t0 > mx && !g_vstats.ns_max_call.compare_exchange_weak(mx, t4 - t0)) {}
    }
}
void solve_hamming_on_gpu_windows(int thread_id, size_t n, const char *qcat, const int64_t *qoff, const uint32_t *ref_id,
                                  const uint32_t *start, const std::vector<std::string> &sequences, int match_score,
                                  int mismatch_score, int gap_open_score, int gap_extend_score, int end_bonus,
                                  int32_t *hamming, rsa_ext_alninfo_t *out) {
    Worker &w = acquire_worker(thread_id, n, n, match_score, mismatch_score, gap_open_score, gap_extend_score);
    if (n == 0) return;
    const long long t0 = now_ns();
    if (!w.has_reference) attach_reference(w, sequences);
    const long long t_attach = now_ns();
    w.woff.resize(n);
    for (size_t i = 0; i < n; ++i) w.woff[i] = g_genome.contig_off[ref_id[i]] + (int64_t)start[i];
    if (rsa_ext_hamming_ref_windows(w.h, (int64_t)n, qcat, qoff, w.woff.data(), end_bonus, hamming, out) != RSA_EXT_OK)
        die("rsa_ext_hamming_ref_windows", w.h);
    if (g_vstats.on) {
        const long long t1 = now_ns();
        g_vstats.ham_calls++; g_vstats.ham_pairs += (long long)n; g_vstats.ns_hamming += t1 - t_attach; g_vstats.ns_attach += t_attach - t0;
        long long mx = g_vstats.ns_max_ham.load();
        while (t1 - t_attach > mx && !g_vstats.ns_max_ham.compare_exchange_weak(mx, t1 - t_attach)) {}
    }
}
#endif  // RSA_EXT_WINDOWS

#ifdef RSA_EXT_ALNINFO
void rsa_ext_veneer_end_bonus(int end_bonus) {
    const int want = (end_bonus & (kBonusConfirmed - 1)) | kBonusConfirmed;
    if (g_end_bonus_state.load(std::memory_order_relaxed) == want) return;
    g_end_bonus_state.store(want, std::memory_order_release);
}
#endif
