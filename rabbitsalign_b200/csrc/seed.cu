// seed.cu -- host side of the seeding C ABI (include/rsa_seed.h), linked into librsa_ext.so.
//
// Replaces, for a whole batch of reads, the per-read host calls randstrobes_query / find_nams / find_nams_rescue of the
// reference (src/aln.cpp:1937-1958); the index the reference built on the host (StrobemerIndex::randstrobes and
// ::randstrobe_start_indices, src/index.hpp:163-184) is uploaded once per GPU and shared by that GPU's workers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels_seed.cuh"

using namespace rsaseed;

namespace {
thread_local std::string g_seed_error;
std::mutex g_seed_cold;  // allocations / first launches take process-wide driver locks

// small tier: what an ordinary read needs (150-250 bp: ~30-50 syncmers, tens of hits, a handful of NAMs)
constexpr Caps kCapsSmall{128, 384, 16, 48, 128, 128};
// large tier: reads from repeats, long reads, many reference sequences
constexpr Caps kCapsLarge{512, 32768, 512, 8192, 16384, 512};
constexpr int kLargeWarpsPerSm = 16;  // the large tier runs one WARP per read (kernels_seed.cuh, CoWarp)
}  // namespace

struct rsa_seed_index {
    rsa_seed_config_t cfg{};
    IndexEntry* d_entries = nullptr;
    uint64_t* d_starts = nullptr;
    int64_t n = 0, n_starts = 0;
    int n_sms = 148;
};

struct rsa_seed {
    rsa_seed_index* ix = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;
    // device
    uint8_t* d_reads = nullptr; size_t reads_cap = 0;
    int64_t* d_roff = nullptr; size_t roff_cap = 0;
    rsa_seed_read_t* d_per = nullptr; size_t per_cap = 0;
    rsa_seed_nam_t* d_nams = nullptr; size_t nams_cap = 0;
    uint32_t* d_retry = nullptr; size_t retry_cap = 0;
    unsigned long long* d_counters = nullptr;
    uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;
    uint8_t* d_scratch_large = nullptr; size_t scratch_large_cap = 0;
    // pinned host
    rsa_seed_read_t* h_per = nullptr; size_t h_per_cap = 0;
    rsa_seed_nam_t* h_nams = nullptr; size_t h_nams_cap = 0;
    unsigned long long* h_counters = nullptr;
    int64_t staged_reads = 0;
    rsa_seed_stats_t stats{};
    std::string err;
};

namespace {

#define SEED_TRY(h, call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return RSA_SEED_ERR_CUDA;                                                      \
        }                                                                                  \
    } while (0)

template <class T>
int grow_dev(rsa_seed* h, T*& p, size_t& cap, size_t need_elems) {
    if (need_elems <= cap) return RSA_SEED_OK;
    std::lock_guard<std::mutex> lk(g_seed_cold);
    const size_t want = std::max(need_elems, cap + cap / 2);
    if (p) SEED_TRY(h, cudaFree(p));
    p = nullptr; cap = 0;
    SEED_TRY(h, cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return RSA_SEED_OK;
}

template <class T>
int grow_pin(rsa_seed* h, T*& p, size_t& cap, size_t need_elems) {
    if (need_elems <= cap) return RSA_SEED_OK;
    std::lock_guard<std::mutex> lk(g_seed_cold);
    const size_t want = std::max(need_elems, cap + cap / 2);
    if (p) SEED_TRY(h, cudaFreeHost(p));
    p = nullptr; cap = 0;
    SEED_TRY(h, cudaHostAlloc(&p, want * sizeof(T), cudaHostAllocDefault));
    cap = want;
    return RSA_SEED_OK;
}

Params make_params(const rsa_seed_index* ix) {
    Params P;
    const rsa_seed_config_t& c = ix->cfg;
    P.k = c.k; P.s = c.s; P.t_syncmer = c.t_syncmer; P.w_min = c.w_min; P.w_max = c.w_max; P.max_dist = c.max_dist;
    P.bits = c.bits; P.rescue_level = c.rescue_level; P.filter_cutoff = c.filter_cutoff; P.rescue_cutoff = c.rescue_cutoff;
    P.q = c.q; P.n_entries = ix->n;
    return P;
}

// Launch the two tiers over the reads resident in d_reads/d_roff; fills d_per / d_nams and the counters.
// Returns RSA_SEED_OK, or 1 when the NAM buffer was too small (h_counters[0] tells how many are needed).
int run_kernels(rsa_seed* h, int64_t n_reads) {
    rsa_seed_index* ix = h->ix;
    const Params P = make_params(ix);
    const Index I{ix->d_entries, ix->d_starts, (long long)ix->n};
    const int blocks = (int)std::min<int64_t>((n_reads + kSeedThreads - 1) / kSeedThreads, (int64_t)ix->n_sms * 4);
    // RSA_SEED_FORCE_LARGE=1 (tests): a small tier that holds nothing, so every read with seeds takes the warp-per-read tier
    static const bool force_large = getenv("RSA_SEED_FORCE_LARGE") != nullptr;
    const Caps caps_small = force_large ? Caps{1, 1, 1, 1, 1, 1} : kCapsSmall;
    const size_t stride = scratch_bytes(kCapsSmall);
    int rc;
    if ((rc = grow_dev(h, h->d_scratch, h->scratch_cap, stride * (size_t)blocks * kSeedThreads))) return rc;
    if ((rc = grow_dev(h, h->d_retry, h->retry_cap, (size_t)n_reads))) return rc;
    SEED_TRY(h, cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->st));
    SEED_TRY(h, cudaEventRecord(h->ev0, h->st));
    seed_kernel<CoThread><<<blocks, kSeedThreads, 0, h->st>>>(h->d_reads, h->d_roff, nullptr, (int)n_reads, I, P, caps_small, h->d_scratch,
                                                     stride, h->d_per, h->d_nams, (unsigned long long)h->nams_cap, h->d_counters,
                                                     h->d_retry, 0);
    h->stats.kernel_launches++;
    SEED_TRY(h, cudaGetLastError());
    SEED_TRY(h, cudaEventRecord(h->evm, h->st));
    SEED_TRY(h, cudaMemcpyAsync(h->h_counters, h->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
    SEED_TRY(h, cudaStreamSynchronize(h->st));
    const int64_t n_retry = (int64_t)h->h_counters[3];
    h->stats.reads_retried = n_retry;
    if (n_retry > 0) {
        // the reads the small tier could not hold: one warp per read, large scratch slices
        const size_t lstride = scratch_bytes(kCapsLarge);
        const int warps_per_block = kSeedThreads / 32;
        const int64_t want_warps = std::min<int64_t>(n_retry, (int64_t)ix->n_sms * kLargeWarpsPerSm);
        const int lblocks = (int)((want_warps + warps_per_block - 1) / warps_per_block);
        if ((rc = grow_dev(h, h->d_scratch_large, h->scratch_large_cap, lstride * (size_t)lblocks * warps_per_block))) return rc;
        seed_kernel<CoWarp><<<lblocks, kSeedThreads, 0, h->st>>>(h->d_reads, h->d_roff, h->d_retry, (int)n_retry, I, P, kCapsLarge,
                                                                  h->d_scratch_large, lstride, h->d_per, h->d_nams,
                                                                  (unsigned long long)h->nams_cap, h->d_counters, nullptr, 1);
        h->stats.kernel_launches++;
        SEED_TRY(h, cudaGetLastError());
    }
    SEED_TRY(h, cudaEventRecord(h->ev1, h->st));
    SEED_TRY(h, cudaMemcpyAsync(h->h_counters, h->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
    SEED_TRY(h, cudaStreamSynchronize(h->st));
    float ms = 0, ms_small = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    cudaEventElapsedTime(&ms_small, h->ev0, h->evm);
    h->stats.kernel_ms += ms;
    h->stats.kernel_ms_large += ms - ms_small;
    return h->h_counters[0] > (unsigned long long)h->nams_cap ? 1 : RSA_SEED_OK;
}

int upload_reads(rsa_seed* h, int64_t n_reads, const char* reads, const int64_t* roff) {
    const int64_t bytes = roff[n_reads] - roff[0];
    if (bytes < 0) { h->err = "offsets are not monotone"; return RSA_SEED_ERR_ARG; }
    int rc;
    if ((rc = grow_dev(h, h->d_reads, h->reads_cap, (size_t)bytes + 16))) return rc;
    if ((rc = grow_dev(h, h->d_roff, h->roff_cap, (size_t)n_reads + 1))) return rc;
    if ((rc = grow_dev(h, h->d_per, h->per_cap, (size_t)n_reads))) return rc;
    if ((rc = grow_dev(h, h->d_nams, h->nams_cap, (size_t)n_reads * 16 + 1024))) return rc;
    // offsets relative to the first read
    if (roff[0] == 0) {
        SEED_TRY(h, cudaMemcpyAsync(h->d_roff, roff, sizeof(int64_t) * (size_t)(n_reads + 1), cudaMemcpyHostToDevice, h->st));
    } else {
        std::vector<int64_t> rel((size_t)n_reads + 1);
        for (int64_t i = 0; i <= n_reads; ++i) rel[(size_t)i] = roff[i] - roff[0];
        SEED_TRY(h, cudaMemcpyAsync(h->d_roff, rel.data(), sizeof(int64_t) * (size_t)(n_reads + 1), cudaMemcpyHostToDevice, h->st));
        SEED_TRY(h, cudaStreamSynchronize(h->st));
    }
    if (bytes) SEED_TRY(h, cudaMemcpyAsync(h->d_reads, reads + roff[0], (size_t)bytes, cudaMemcpyHostToDevice, h->st));
    h->stats.h2d_bytes += bytes + (int64_t)sizeof(int64_t) * (n_reads + 1);
    return RSA_SEED_OK;
}

}  // namespace

extern "C" const char* rsa_seed_last_error(const rsa_seed_t* h) { return h ? h->err.c_str() : g_seed_error.c_str(); }

extern "C" int rsa_seed_index_upload(const rsa_seed_config_t* cfg, const void* randstrobes, int64_t n, const uint64_t* starts,
                                     int64_t n_starts, rsa_seed_index_t** out) {
    if (!out) return RSA_SEED_ERR_ARG;
    *out = nullptr;
    if (!cfg || !starts || n < 0 || (n > 0 && !randstrobes)) { g_seed_error = "bad argument"; return RSA_SEED_ERR_ARG; }
    if (cfg->k < 8 || cfg->k > 32 || cfg->s < 1 || cfg->s > cfg->k || cfg->k - cfg->s + 1 > 32 || cfg->bits < 8 || cfg->bits > 31 ||
        cfg->w_min < 0 || cfg->w_max < cfg->w_min || n_starts != ((int64_t)1 << cfg->bits) + 1) {
        g_seed_error = "inconsistent index parameters (k in [8,32], s <= k, bits in [8,31], (1 << bits) + 1 bucket starts)";
        return RSA_SEED_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_seed_cold);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_seed_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
        return RSA_SEED_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_seed_error = "bad device ordinal"; return RSA_SEED_ERR_ARG; }
    auto ix = std::make_unique<rsa_seed_index>();
    ix->cfg = *cfg;
    ix->n = n;
    ix->n_starts = n_starts;
    auto fail = [&](const char* what, cudaError_t ce) {
        g_seed_error = std::string(what) + ": " + cudaGetErrorString(ce);
        if (ix->d_entries) cudaFree(ix->d_entries);
        if (ix->d_starts) cudaFree(ix->d_starts);
        return RSA_SEED_ERR_CUDA;
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaMalloc(&ix->d_entries, sizeof(IndexEntry) * (size_t)std::max<int64_t>(n, 1))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ix->d_starts, sizeof(uint64_t) * (size_t)n_starts)) != cudaSuccess) return fail("cudaMalloc", e);
    if (n && (e = cudaMemcpy(ix->d_entries, randstrobes, sizeof(IndexEntry) * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(ix->d_starts, starts, sizeof(uint64_t) * (size_t)n_starts, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("cudaDeviceSynchronize", e);
    cudaDeviceGetAttribute(&ix->n_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    *out = ix.release();
    return RSA_SEED_OK;
}

extern "C" void rsa_seed_index_free(rsa_seed_index_t* ix) {
    if (!ix) return;
    cudaSetDevice(ix->cfg.device);
    if (ix->d_entries) cudaFree(ix->d_entries);
    if (ix->d_starts) cudaFree(ix->d_starts);
    delete ix;
}

extern "C" int rsa_seed_create(rsa_seed_index_t* ix, rsa_seed_t** out) {
    if (!out) return RSA_SEED_ERR_ARG;
    *out = nullptr;
    if (!ix) { g_seed_error = "no index"; return RSA_SEED_ERR_ARG; }
    std::unique_lock<std::mutex> lk(g_seed_cold);
    auto h = std::make_unique<rsa_seed>();
    h->ix = ix;
    cudaError_t e;
    auto fail = [&](const char* what, cudaError_t ce) {
        g_seed_error = std::string(what) + ": " + cudaGetErrorString(ce);
        return RSA_SEED_ERR_CUDA;
    };
    if ((e = cudaSetDevice(ix->cfg.device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreate(&h->evm)) != cudaSuccess) return fail("event", e);
    if ((e = cudaMalloc(&h->d_counters, 4 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&h->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocDefault)) != cudaSuccess) return fail("pinned", e);
    *out = h.release();
    return RSA_SEED_OK;
}

extern "C" void rsa_seed_destroy(rsa_seed_t* h) {
    if (!h) return;
    cudaSetDevice(h->ix->cfg.device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (void* p : {(void*)h->d_reads, (void*)h->d_roff, (void*)h->d_per, (void*)h->d_nams, (void*)h->d_retry, (void*)h->d_counters,
                    (void*)h->d_scratch, (void*)h->d_scratch_large})
        if (p) cudaFree(p);
    for (void* p : {(void*)h->h_per, (void*)h->h_nams, (void*)h->h_counters})
        if (p) cudaFreeHost(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evm) cudaEventDestroy(h->evm);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rsa_seed_find_nams(rsa_seed_t* h, int64_t n_reads, const char* reads, const int64_t* roff,
                                  const rsa_seed_read_t** per_read, const rsa_seed_nam_t** nams, int64_t* n_nams) {
    if (!h) return RSA_SEED_ERR_ARG;
    if (n_reads < 0 || !roff || (n_reads > 0 && !reads) || !per_read || !nams || !n_nams || n_reads > 0x7FFFFFFF) {
        h->err = "bad argument";
        return RSA_SEED_ERR_ARG;
    }
    *per_read = nullptr; *nams = nullptr; *n_nams = 0;
    h->stats = rsa_seed_stats_t{};
    h->stats.reads = n_reads;
    if (n_reads == 0) return RSA_SEED_OK;
    for (int64_t i = 0; i < n_reads; ++i)
        if (roff[i + 1] < roff[i] || roff[i + 1] - roff[i] > 65535) { h->err = "read lengths must be in [0, 65535]"; return RSA_SEED_ERR_ARG; }
    SEED_TRY(h, cudaSetDevice(h->ix->cfg.device));
    int rc;
    if ((rc = upload_reads(h, n_reads, reads, roff))) return rc;
    for (int attempt = 0;; ++attempt) {
        rc = run_kernels(h, n_reads);
        if (rc < 0) return rc;
        if (rc == 0) break;
        if (attempt >= 2) { h->err = "NAM buffer still too small after growing it"; return RSA_SEED_ERR_STATE; }
        if ((rc = grow_dev(h, h->d_nams, h->nams_cap, (size_t)h->h_counters[0] + 1024))) return rc;  // and run again
    }
    const int64_t total = (int64_t)h->h_counters[0];
    if ((rc = grow_pin(h, h->h_per, h->h_per_cap, (size_t)n_reads))) return rc;
    if ((rc = grow_pin(h, h->h_nams, h->h_nams_cap, (size_t)total + 1))) return rc;
    SEED_TRY(h, cudaMemcpyAsync(h->h_per, h->d_per, sizeof(rsa_seed_read_t) * (size_t)n_reads, cudaMemcpyDeviceToHost, h->st));
    if (total) SEED_TRY(h, cudaMemcpyAsync(h->h_nams, h->d_nams, sizeof(rsa_seed_nam_t) * (size_t)total, cudaMemcpyDeviceToHost, h->st));
    SEED_TRY(h, cudaStreamSynchronize(h->st));
    h->stats.d2h_bytes = (int64_t)(sizeof(rsa_seed_read_t) * (size_t)n_reads + sizeof(rsa_seed_nam_t) * (size_t)total);
    h->stats.nams = total;
    h->stats.reads_failed = (int64_t)h->h_counters[1];
    h->stats.reads_rescued = (int64_t)h->h_counters[2];
    *per_read = h->h_per;
    *nams = h->h_nams;
    *n_nams = total;
    return RSA_SEED_OK;
}

extern "C" int rsa_seed_get_stats(const rsa_seed_t* h, rsa_seed_stats_t* out) {
    if (!h || !out) return RSA_SEED_ERR_ARG;
    *out = h->stats;
    return RSA_SEED_OK;
}

extern "C" int rsa_seed_stage(rsa_seed_t* h, int64_t n_reads, const char* reads, const int64_t* roff) {
    if (!h || n_reads <= 0 || !reads || !roff || n_reads > 0x7FFFFFFF) { if (h) h->err = "bad argument"; return RSA_SEED_ERR_ARG; }
    SEED_TRY(h, cudaSetDevice(h->ix->cfg.device));
    h->stats = rsa_seed_stats_t{};
    int rc = upload_reads(h, n_reads, reads, roff);
    if (rc) return rc;
    SEED_TRY(h, cudaStreamSynchronize(h->st));
    h->staged_reads = n_reads;
    return RSA_SEED_OK;
}

extern "C" int rsa_seed_run_staged(rsa_seed_t* h) {
    if (!h) return RSA_SEED_ERR_ARG;
    if (h->staged_reads <= 0) { h->err = "nothing staged"; return RSA_SEED_ERR_STATE; }
    SEED_TRY(h, cudaSetDevice(h->ix->cfg.device));
    h->stats.reads = h->staged_reads;
    h->stats.kernel_ms = 0;
    h->stats.kernel_ms_large = 0;
    h->stats.kernel_launches = 0;
    for (int attempt = 0;; ++attempt) {
        int rc = run_kernels(h, h->staged_reads);
        if (rc < 0) return rc;
        if (rc == 0) break;
        if (attempt >= 2) { h->err = "NAM buffer still too small after growing it"; return RSA_SEED_ERR_STATE; }
        if ((rc = grow_dev(h, h->d_nams, h->nams_cap, (size_t)h->h_counters[0] + 1024))) return rc;
        h->stats.kernel_ms = 0;
    }
    h->stats.nams = (int64_t)h->h_counters[0];
    h->stats.reads_failed = (int64_t)h->h_counters[1];
    h->stats.reads_rescued = (int64_t)h->h_counters[2];
    return RSA_SEED_OK;
}

extern "C" void* rsa_seed_stream(rsa_seed_t* h) { return h ? (void*)h->st : nullptr; }
