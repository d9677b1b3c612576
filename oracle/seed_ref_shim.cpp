// oracle/seed_ref_shim.cpp -- TEST INFRASTRUCTURE (checker + CPU baseline), never on the product path.
//
// The reference's own seeding path, compiled where it lies under /root/reference by oracle/Makefile
// (src/randstrobes.cpp, src/nam.cpp, src/index.cpp, src/refs.cpp, src/indexparameters.cpp, src/io.cpp) into
// oracle/_ref/libseed_ref.so and driven the way src/aln.cpp:1927-1958 (align_PE_read_part) / :2380-2400
// (align_SE_read_part) drives it per read:
//
//     query_randstrobes = randstrobes_query(seq, index_parameters)                 src/randstrobes.cpp:207
//     [fraction, nams]  = find_nams(query_randstrobes, index)                      src/nam.cpp:771
//     if (rescue_level > 1 && (nams.empty() || fraction < 0.7))
//         nams = find_nams_rescue(query_randstrobes, index, rescue_cutoff)         src/nam.cpp:955
//
// NAMs come back in the order the reference's vector holds them BEFORE its by-score sort (nam_id = index).
// Also exported: the index arrays (so the GPU path searches the very same index), the iteration order of the
// reference's robin_hood map for a set of reference ids (the one piece of the NAM order that is a property of that
// container, not of the algorithm), and a multi-threaded timing loop for the CPU baseline.
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "refs.hpp"
#include "index.hpp"
#include "indexparameters.hpp"
#include "randstrobes.hpp"
#include "nam.hpp"
#include "robin_hood.h"

namespace {
struct SeedRef {
    References refs;
    IndexParameters params;
    StrobemerIndex index;
    SeedRef(std::vector<std::string> seqs, std::vector<std::string> names, int read_len)
        : refs(std::move(seqs), std::move(names)), params(IndexParameters::from_read_length(read_len)), index(refs, params) {}
};
}  // namespace

// one output record per NAM: the fields of struct Nam (src/nam.hpp:11-38)
struct SeedRefNam {
    int32_t query_start, query_end, query_prev_hit_startpos;
    int32_t ref_start, ref_end, ref_prev_hit_startpos;
    int32_t n_hits, ref_id;
    float score;
    int32_t is_rc;
};

extern "C" {

void* seedref_build(const char* concat, const int64_t* contig_off, int n_contigs, int read_len, int threads) {
    std::vector<std::string> seqs, names;
    for (int c = 0; c < n_contigs; ++c) {
        seqs.emplace_back(concat + contig_off[c], (size_t)(contig_off[c + 1] - contig_off[c]));
        names.push_back("contig" + std::to_string(c));
    }
    SeedRef* s = new SeedRef(std::move(seqs), std::move(names), read_len);
    s->index.populate(0.0002f, (size_t)threads);  // src/main.cpp:369, opt.f default (src/cmdline.hpp)
    return s;
}

void seedref_free(void* h) { delete static_cast<SeedRef*>(h); }

// ints: [0] bits, [1] filter_cutoff, [2] k, [3] s, [4] t_syncmer, [5] w_min, [6] w_max, [7] max_dist
int seedref_export(void* h, const void** randstrobes, int64_t* n, const uint64_t** starts, int64_t* n_starts, int32_t* ints,
                   uint64_t* q) {
    SeedRef* s = static_cast<SeedRef*>(h);
    static_assert(sizeof(RefRandstrobe) == 16, "RefRandstrobe layout");
    *randstrobes = s->index.randstrobes.data();
    *n = (int64_t)s->index.randstrobes.size();
    *starts = s->index.randstrobe_start_indices.data();
    *n_starts = (int64_t)s->index.randstrobe_start_indices.size();
    ints[0] = s->index.bits;
    ints[1] = (int32_t)s->index.filter_cutoff;
    ints[2] = s->params.syncmer.k;
    ints[3] = s->params.syncmer.s;
    ints[4] = s->params.syncmer.t_syncmer;
    ints[5] = (int32_t)s->params.randstrobe.w_min;
    ints[6] = (int32_t)s->params.randstrobe.w_max;
    ints[7] = s->params.randstrobe.max_dist;
    *q = s->params.randstrobe.q;
    return 0;
}

static void seed_one(const SeedRef* s, const char* seq, size_t len, int rescue_level, unsigned rescue_cutoff,
                     std::vector<Nam>& nams, float& fraction, bool& rescued) {
    auto query_randstrobes = randstrobes_query(std::string_view(seq, len), s->params);
    auto res = find_nams(query_randstrobes, s->index);
    fraction = res.first;
    nams = std::move(res.second);
    rescued = false;
    if (rescue_level > 1 && (nams.empty() || fraction < 0.7)) {
        nams = find_nams_rescue(query_randstrobes, s->index, rescue_cutoff);
        rescued = true;
    }
}

// Pass 1 (out == nullptr): counts only, returns the total.  Pass 2: fills out[] (reads in order, NAMs in vector order).
int64_t seedref_find_nams(void* h, const char* reads, const int64_t* roff, int64_t n_reads, int rescue_level, int rescue_cutoff,
                          int32_t* nam_count, float* fraction, uint8_t* rescued, SeedRefNam* out, int64_t out_cap) {
    SeedRef* s = static_cast<SeedRef*>(h);
    int64_t total = 0;
    std::vector<Nam> nams;
    for (int64_t i = 0; i < n_reads; ++i) {
        float f;
        bool r;
        seed_one(s, reads + roff[i], (size_t)(roff[i + 1] - roff[i]), rescue_level, (unsigned)rescue_cutoff, nams, f, r);
        nam_count[i] = (int32_t)nams.size();
        fraction[i] = f;
        rescued[i] = r ? 1 : 0;
        if (out) {
            for (const Nam& n : nams) {
                if (total >= out_cap) return -1;
                out[total++] = SeedRefNam{n.query_start, n.query_end, n.query_prev_hit_startpos, n.ref_start, n.ref_end,
                                          n.ref_prev_hit_startpos, n.n_hits, n.ref_id, n.score, n.is_rc ? 1 : 0};
            }
        } else {
            total += (int64_t)nams.size();
        }
    }
    return total;
}

// randstrobes_query of one read (src/randstrobes.cpp:207): out = n x {hash, start, end, is_reverse} as uint64; returns n
int64_t seedref_randstrobes(void* h, const char* seq, int64_t len, uint64_t* out, int64_t cap) {
    SeedRef* s = static_cast<SeedRef*>(h);
    auto qr = randstrobes_query(std::string_view(seq, (size_t)len), s->params);
    int64_t n = 0;
    for (const auto& r : qr) {
        if (n >= cap) return -1;
        out[4 * n] = r.hash; out[4 * n + 1] = r.start; out[4 * n + 2] = r.end; out[4 * n + 3] = r.is_reverse ? 1 : 0;
        ++n;
    }
    return n;
}

// CPU baseline: the same per-read work on `threads` threads (reads split into contiguous ranges); returns NAMs found.
int64_t seedref_time(void* h, const char* reads, const int64_t* roff, int64_t n_reads, int rescue_level, int rescue_cutoff,
                     int threads) {
    SeedRef* s = static_cast<SeedRef*>(h);
    std::vector<int64_t> found((size_t)threads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t] {
            const int64_t lo = n_reads * t / threads, hi = n_reads * (t + 1) / threads;
            std::vector<Nam> nams;
            for (int64_t i = lo; i < hi; ++i) {
                float f;
                bool r;
                seed_one(s, reads + roff[i], (size_t)(roff[i + 1] - roff[i]), rescue_level, (unsigned)rescue_cutoff, nams, f, r);
                found[(size_t)t] += (int64_t)nams.size();
            }
        });
    for (auto& x : th) x.join();
    int64_t total = 0;
    for (int64_t v : found) total += v;
    return total;
}

// Iteration order of the reference's hits_per_ref container (robin_hood::unordered_map<unsigned int, std::vector<Hit>>,
// reserve(100): src/nam.cpp:775-777) after inserting `keys` in this order: out[] = keys in iteration order.
void seedref_map_order(const uint32_t* keys, int n, uint32_t* out) {
    struct Hit4 { int a, b, c, d; };
    robin_hood::unordered_map<unsigned int, std::vector<Hit4>> m;
    m.reserve(100);
    for (int i = 0; i < n; ++i) m[keys[i]];
    int k = 0;
    for (auto& kv : m) out[k++] = kv.first;
}

}  // extern "C"
