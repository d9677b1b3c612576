// kernels_seed.cuh -- seeding on the device: syncmers -> randstrobes -> index lookup -> hits -> NAM merge (+ rescue).
//
// One thread per read walks the whole per-read pipeline of the reference (src/aln.cpp:1937-1958) with the reference's
// control flow and tie rules, so every intermediate list has the reference's order:
//
//   read_syncmers      SyncmerIterator::next                      src/randstrobes.cpp:57-127   (xxh64: src/hash.hpp:104-119)
//   randstrobe_at      RandstrobeIterator::get                    src/randstrobes.cpp:151-176
//   index_find         StrobemerIndex::find / is_filtered / get_count   src/index.hpp:60-97,107-109,127-161
//   emit_hits          add_to_hits_per_ref (+ _pre for rescue)    src/nam.cpp:68-107
//   merge_group        merge_hits_into_nams                       src/nam.cpp:368-510  (find_nams, sort = true)
//   merge_group_fast   merge_hits_into_nams_fast                  src/nam.cpp:117-366  (find_nams_rescue, pre_sort build)
//   seed_read          find_nams (src/nam.cpp:771-922) and find_nams_rescue (:955-1012)
//
// Why thread-per-read: the per-read work is a chain of small, data-dependent lists (30 syncmers, 50 randstrobes, tens of
// hits) whose order is part of the result; the parallelism is across the reads of a batch (10^4 in a pipeline chunk, 10^6
// in a bench batch), and the time goes into ~5 dependent random 16-32 byte probes of the index per randstrobe, i.e. the
// kernel is bound by HBM/L2 latency x sectors, which tens of thousands of resident threads cover.  Each thread owns a
// scratch slice in global memory (two tiers: a small one for the common case, a large one for the reads that overflow
// it); the lists stay in L1/L2 while a read is processed.
//
// The reference's hits_per_ref is a hash map keyed by reference id; its iteration order decides in which order the
// groups' NAMs are appended.  Here groups are kept in first-touch order and every NAM carries its group index; the
// binding applies the container's order (include/rsa_seed.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/rsa_seed.h"

// every per-read function is also compiled for the host: tests/seed_host_check.cu runs this very code on the CPU against
// the reference's seeding path where no GPU is available (test infrastructure; the product library only launches kernels)
#define RSA_SEED_HD __host__ __device__

namespace rsaseed {

struct IndexEntry {  // == RefRandstrobe (src/randstrobes.hpp:21-50)
    uint64_t hash;
    uint32_t position;
    uint32_t packed;  // ref_index << 8 | strobe2_offset
};

struct Params {
    int k, s, t_syncmer, w_min, w_max, max_dist, bits, rescue_level;
    uint32_t filter_cutoff, rescue_cutoff;
    uint64_t q;
    long long n_entries;
};

struct Hit {  // 16 bytes
    uint16_t qs, qe;
    int32_t rs, re;
    uint16_t grp, pad;
};

struct RescueHit {  // src/nam.cpp:936-942
    uint32_t position, count;
    uint16_t qs, qe;
    uint32_t pad;
};

struct Caps {
    int syn, hits, groups, open, nams, resc;
};

RSA_SEED_HD inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
RSA_SEED_HD inline size_t scratch_bytes(const Caps& c) {
    return align16((size_t)c.syn * 8) + align16((size_t)c.syn * 4) + 2 * align16((size_t)c.hits * sizeof(Hit)) +
           align16((size_t)c.groups * 2 * 4) + align16((size_t)c.open * sizeof(rsa_seed_nam_t)) +
           align16((size_t)c.nams * sizeof(rsa_seed_nam_t)) + 2 * align16((size_t)c.resc * sizeof(RescueHit));
}

struct Scratch {
    uint64_t* syn_hash;
    int32_t* syn_pos;
    Hit* hits;       // as emitted (one strand at a time)
    Hit* grouped;    // one group, contiguous
    uint32_t* group_ref;  // [2][groups]: reference ids in first-touch order per strand
    rsa_seed_nam_t* open;
    rsa_seed_nam_t* nams;
    RescueHit* resc;  // [2][resc]
    RSA_SEED_HD Scratch(uint8_t* base, const Caps& c) {
        uint8_t* p = base;
        syn_hash = reinterpret_cast<uint64_t*>(p); p += align16((size_t)c.syn * 8);
        syn_pos = reinterpret_cast<int32_t*>(p); p += align16((size_t)c.syn * 4);
        hits = reinterpret_cast<Hit*>(p); p += align16((size_t)c.hits * sizeof(Hit));
        grouped = reinterpret_cast<Hit*>(p); p += align16((size_t)c.hits * sizeof(Hit));
        group_ref = reinterpret_cast<uint32_t*>(p); p += align16((size_t)c.groups * 2 * 4);
        open = reinterpret_cast<rsa_seed_nam_t*>(p); p += align16((size_t)c.open * sizeof(rsa_seed_nam_t));
        nams = reinterpret_cast<rsa_seed_nam_t*>(p); p += align16((size_t)c.nams * sizeof(rsa_seed_nam_t));
        resc = reinterpret_cast<RescueHit*>(p);
    }
};

// ---- hashing (src/hash.hpp:104-119: xxh64 of one 64-bit value) --------------------------------------------------------
RSA_SEED_HD __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
RSA_SEED_HD __forceinline__ uint64_t xxh64_u64(uint64_t input) {
    constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
                       P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t result = P5 + 8;
    input *= P2;
    input = rotl64(input, 31);
    result ^= input * P1;
    result = rotl64(result, 27);
    result = result * P1 + P4;
    result ^= result >> 33;
    result = result * P2;
    result ^= result >> 29;
    result = result * P3;
    result ^= result >> 32;
    return result;
}

// a,A -> 0; c,C -> 1; g,G -> 2; t,T,u,U -> 3; everything else 4 (seq_nt4_table, src/randstrobes.cpp:13-30)
RSA_SEED_HD __forceinline__ int nt4(uint8_t ch) {
    const uint8_t u = ch & 0xDFu;  // upper case
    if (u == 'A') return 0;
    if (u == 'C') return 1;
    if (u == 'G') return 2;
    if (u == 'T' || u == 'U') return 3;
    return ch < 4 ? (int)ch : 4;  // the table also maps the raw bytes 0..3 to themselves
}

// ---- syncmers (SyncmerIterator::next, src/randstrobes.cpp:57-127) -----------------------------------------------------
// Returns the number of syncmers, or -1 when they do not fit `cap`.
template <class Co>
RSA_SEED_HD inline int read_syncmers(const uint8_t* seq, int len, const Params& P, uint64_t* out_hash, int32_t* out_pos, int cap) {
    const int k = P.k, s = P.s, t = P.t_syncmer;
    const uint64_t kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1), smask = s >= 32 ? ~0ull : ((1ull << (2 * s)) - 1);
    const int kshift = (k - 1) * 2, sshift = (s - 1) * 2;
    const int win = k - s + 1;  // s-mers per k-mer
    uint64_t qs[32];            // the deque of s-mer hashes (a ring; win <= 25 because k <= 32 and s >= 8)
    int head = 0, size = 0;
    uint64_t min_val = ~0ull;
    long long min_pos = -1;
    int l = 0, n = 0;
    uint64_t xk0 = 0, xk1 = 0, xs0 = 0, xs1 = 0;
    for (int i = 0; i < len; ++i) {
        const int c = nt4(seq[i]);
        if (c < 4) {
            xk0 = (xk0 << 2 | (uint64_t)c) & kmask;
            xk1 = xk1 >> 2 | (uint64_t)(3 - c) << kshift;
            xs0 = (xs0 << 2 | (uint64_t)c) & smask;
            xs1 = xs1 >> 2 | (uint64_t)(3 - c) << sshift;
            if (++l < s) continue;
            const uint64_t hash_s = xxh64_u64(xs0 < xs1 ? xs0 : xs1);
            qs[(head + size) & 31] = hash_s;
            ++size;
            if (size < win) continue;
            if (size == win) {
                // the last s-mer of the first k-mer: leftmost minimum (strict <, forward scan)
                for (int j = 0; j < size; ++j) {
                    const uint64_t v = qs[(head + j) & 31];
                    if (v < min_val) { min_val = v; min_pos = (long long)i - k + j + 1; }
                }
            } else {
                head = (head + 1) & 31;
                --size;
                if (min_pos == (long long)i - k) {
                    // the minimum left the window: brute force, rightmost minimum (reverse scan, strict <)
                    min_val = ~0ull;
                    min_pos = (long long)i - s + 1;
                    for (int j = size - 1; j >= 0; --j) {
                        const uint64_t v = qs[(head + j) & 31];
                        if (v < min_val) { min_val = v; min_pos = (long long)i - k + j + 1; }
                    }
                } else if (hash_s < min_val) {
                    min_val = hash_s;
                    min_pos = (long long)i - s + 1;
                }
            }
            if (min_pos == (long long)i - k + t) {
                if (n >= cap) return -1;
                if (Co::leader()) {
                    out_hash[n] = xxh64_u64(xk0 < xk1 ? xk0 : xk1);
                    out_pos[n] = i - k + 1;
                }
                ++n;
            }
        } else {
            min_val = ~0ull;
            min_pos = -1;
            l = 0;
            xs0 = xs1 = xk0 = xk1 = 0;
            head = 0;
            size = 0;
        }
    }
    return n;
}

// ---- randstrobes (RandstrobeIterator::get, src/randstrobes.cpp:151-176; reverse strand: :236-253) ------------------
struct Syncmers {
    const uint64_t* hash;
    const int32_t* pos;
    int n, len, k;
    bool rev;  // the reversed list with positions len - pos - k
    RSA_SEED_HD __forceinline__ uint64_t h(int i) const { return hash[rev ? n - 1 - i : i]; }
    RSA_SEED_HD __forceinline__ int p(int i) const { return rev ? len - pos[n - 1 - i] - k : pos[i]; }
};

RSA_SEED_HD __forceinline__ void randstrobe_at(const Syncmers& S, const Params& P, int idx, uint64_t& hash, int& start, int& end) {
    const int w_end = min(idx + P.w_max, S.n - 1);
    const uint64_t h1 = S.h(idx);
    const int p1 = S.p(idx);
    const int max_position = p1 + P.max_dist;
    uint64_t min_val = ~0ull, h2 = h1;
    int p2 = p1;
    for (int i = idx + P.w_min; i <= w_end && S.p(i) <= max_position; ++i) {
        const uint64_t hi = S.h(i);
#ifdef __CUDA_ARCH__
        const uint64_t res = (uint64_t)__popcll((h1 ^ hi) & P.q);
#else
        const uint64_t res = (uint64_t)__builtin_popcountll((h1 ^ hi) & P.q);
#endif
        if (res < min_val) { min_val = res; h2 = hi; p2 = S.p(i); }
    }
    hash = h1 + h2;
    start = p1;
    end = p2 + P.k;
}

// ---- index (src/index.hpp) --------------------------------------------------------------------------------------------
struct Index {
    const IndexEntry* e;
    const uint64_t* starts;
    long long n;
    RSA_SEED_HD __forceinline__ uint64_t get_hash(long long pos) const { return (pos >= 0 && pos < n) ? e[pos].hash : ~0ull; }
};

// StrobemerIndex::find (src/index.hpp:60-85; both of its search branches return the first entry of the bucket whose hash
// equals the key): position or -1
RSA_SEED_HD inline long long index_find(const Index& ix, const Params& P, uint64_t key) {
    const uint64_t top = key >> (64 - P.bits);
    long long lo = (long long)ix.starts[top], hi = (long long)ix.starts[top + 1];
    if (lo == hi) return -1;
    const long long end = hi;
    while (lo < hi) {  // lower_bound on the hash
        const long long mid = lo + ((hi - lo) >> 1);
        if (ix.e[mid].hash < key) lo = mid + 1; else hi = mid;
    }
    return (lo < end && ix.e[lo].hash == key) ? lo : -1;
}

RSA_SEED_HD __forceinline__ bool index_is_filtered(const Index& ix, const Params& P, long long pos) {
    return ix.get_hash(pos) == ix.get_hash(pos + (long long)P.filter_cutoff);  // src/index.hpp:107-109
}

// StrobemerIndex::get_count (src/index.hpp:127-161): length of the run of equal hashes from `pos` inside its bucket
RSA_SEED_HD inline uint32_t index_get_count(const Index& ix, const Params& P, long long pos) {
    const uint64_t key = ix.e[pos].hash;
    const long long end = (long long)ix.starts[(key >> (64 - P.bits)) + 1];
    long long p = pos + 1;
    if (end - pos < 8) {
        while (p < end && ix.e[p].hash == key) ++p;
        return (uint32_t)(p - pos);
    }
    long long lo = pos, hi = end;  // upper_bound
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (ix.e[mid].hash <= key) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)(lo - pos);
}

// ---- cooperation policy ------------------------------------------------------------------------------------------------
// The per-read code below is written once and instantiated twice:
//   CoThread  one thread per read (small scratch tier; also what tests/seed_host_check.cu compiles for the host);
//   CoWarp    one warp per read (large tier: reads from repeats, whose hit lists run into the thousands).  All 32 lanes
//             execute the same sequential program on the same values (registers replicated, control flow warp-uniform);
//             scratch lists are written by lane 0 only and become visible to the others at the next sync().  The loops
//             that make such reads slow -- "which open NAM takes this hit" over hundreds of open NAMs, gathering a group
//             out of thousands of hits -- are split over the lanes, with ballots keeping the reference's first-match order.
struct CoThread {
    static constexpr int kLanes = 1;
    RSA_SEED_HD static int lane() { return 0; }
    RSA_SEED_HD static bool leader() { return true; }
    RSA_SEED_HD static void sync() {}
    RSA_SEED_HD static unsigned ballot(bool p) { return p ? 1u : 0u; }
    RSA_SEED_HD static int bcast(int v, int) { return v; }
    RSA_SEED_HD static int first_conflict(int) { return 1; }  // a single lane cannot collide with a lower one
};
#ifdef __CUDACC__
struct CoWarp {
    static constexpr int kLanes = 32;
    __device__ static int lane() { return (int)(threadIdx.x & 31u); }
    __device__ static bool leader() { return (threadIdx.x & 31u) == 0; }
    __device__ static void sync() { __syncwarp(); }
    __device__ static unsigned ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ static int bcast(int v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
    // lowest lane whose candidate (>= 0) is also the candidate of a lower lane; 32 when there is none
    __device__ static int first_conflict(int cand) {
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, cand);
        const bool conflict = cand >= 0 && (peers & ((1u << (threadIdx.x & 31u)) - 1u)) != 0;
        const unsigned cm = __ballot_sync(0xFFFFFFFFu, conflict);
        return cm ? __ffs((int)cm) - 1 : 32;
    }
};
#endif
RSA_SEED_HD __forceinline__ int first_bit(unsigned m) {  // index of the lowest set bit (m != 0)
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ffs((int)m) - 1;
#endif
}
RSA_SEED_HD __forceinline__ int pop_count(unsigned m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

// ---- per-read state ---------------------------------------------------------------------------------------------------
template <class Co>
struct ReadCtx {
    const Index& ix;
    const Params& P;
    const Caps& caps;
    Scratch& sc;
    int n_hits;        // hits of the current strand
    int n_groups[2];
    int n_nams;
    bool overflow;
    RSA_SEED_HD ReadCtx(const Index& ix_, const Params& P_, const Caps& caps_, Scratch& sc_)
        : ix(ix_), P(P_), caps(caps_), sc(sc_), n_hits(0), n_nams(0), overflow(false) { n_groups[0] = n_groups[1] = 0; }

    // hits_per_ref[ref_id] (operator[]): the group of a reference id, created on first touch
    RSA_SEED_HD int group_of(int strand, uint32_t ref_id) {
        uint32_t* g = sc.group_ref + strand * caps.groups;
        for (int i = 0; i < n_groups[strand]; ++i)
            if (g[i] == ref_id) return i;
        if (n_groups[strand] >= caps.groups) { overflow = true; return 0; }
        if (Co::leader()) g[n_groups[strand]] = ref_id;
        Co::sync();
        return n_groups[strand]++;
    }

    // add_to_hits_per_ref (src/nam.cpp:68-86); pre == true: add_to_hits_per_ref_pre (:88-107), which only creates the keys
    RSA_SEED_HD void emit_hits(int strand, int qs, int qe, long long position, bool pre) {
        int min_diff = 0x7FFFFFFF;
        const uint64_t hash = ix.get_hash(position);
        for (; ix.get_hash(position) == hash; ++position) {
            const IndexEntry en = ix.e[position];
            const int ref_start = (int)en.position;
            const int ref_end = ref_start + (int)(en.packed & 0xFFu) + P.k;
            const int d = (qe - qs) - (ref_end - ref_start);
            const int diff = d < 0 ? -d : d;
            if (diff <= min_diff) {
                const int g = group_of(strand, en.packed >> 8);
                if (overflow) return;
                if (!pre) {
                    if (n_hits >= caps.hits) { overflow = true; return; }
                    if (Co::leader()) {
                        Hit h;
                        h.qs = (uint16_t)qs; h.qe = (uint16_t)qe; h.rs = ref_start; h.re = ref_end; h.grp = (uint16_t)g; h.pad = 0;
                        sc.hits[n_hits] = h;
                    }
                    ++n_hits;
                }
                min_diff = diff;
            }
        }
    }

    RSA_SEED_HD void push_nam(rsa_seed_nam_t n, int strand, int grp) {  // nams.push_back with the score (src/nam.cpp:476-486)
        const int qspan = n.query_end - n.query_start, rspan = n.ref_end - n.ref_start;
        const int n_max = max(qspan, rspan), n_min = min(qspan, rspan);
        n.score = (2 * n_min - n_max) > 0 ? (float)(n.n_hits * (2 * n_min - n_max)) : 1.0f;
        n.flags = (uint32_t)strand | ((uint32_t)grp << 8);
        if (n_nams >= caps.nams) { overflow = true; return; }
        if (Co::leader()) sc.nams[n_nams] = n;
        ++n_nams;
    }

    // the hits of group g of the current strand, in emission order, contiguous in sc.grouped; returns their number
    RSA_SEED_HD int gather_group(int g) {
        int m = 0;
        for (int base = 0; base < n_hits; base += Co::kLanes) {
            const int i = base + Co::lane();
            const bool mine = i < n_hits && sc.hits[i].grp == g;
            const unsigned mask = Co::ballot(mine);
            if (mine) sc.grouped[m + pop_count(mask & ((1u << Co::lane()) - 1u))] = sc.hits[i];
            m += pop_count(mask);
        }
        Co::sync();
        return m;
    }

    // std::sort(hits) on (query_start, ref_start) (src/nam.cpp:18-21,396).  The keys of one group are unique, so any sort
    // gives the reference's order; the hits arrive sorted (increasing strobe positions, index entries sorted by position),
    // which makes this insertion sort a single pass.
    RSA_SEED_HD void sort_group(int m) {
        if (Co::leader()) {
            Hit* a = sc.grouped;
            for (int i = 1; i < m; ++i) {
                const Hit x = a[i];
                int j = i - 1;
                while (j >= 0 && (a[j].qs > x.qs || (a[j].qs == x.qs && a[j].rs > x.rs))) { a[j + 1] = a[j]; --j; }
                a[j + 1] = x;
            }
        }
        Co::sync();
    }

    RSA_SEED_HD static rsa_seed_nam_t new_nam(const Hit& h, uint32_t ref_id) {
        rsa_seed_nam_t n;
        n.query_start = h.qs; n.query_end = h.qe; n.ref_start = h.rs; n.ref_end = h.re;
        n.ref_id = (int32_t)ref_id;
        n.query_prev_hit_startpos = h.qs; n.ref_prev_hit_startpos = h.rs;
        n.n_hits = 1; n.score = 0; n.flags = 0;
        return n;
    }

    // flush the open NAMs the current hit has passed (query_end < c) and drop them, keeping the others' order
    RSA_SEED_HD void flush_passed(int& n_open, int c, int strand, int grp) {
        int keep = 0;
        for (int o = 0; o < n_open; ++o) {
            if (sc.open[o].query_end < c) push_nam(sc.open[o], strand, grp);
            else ++keep;
        }
        Co::sync();  // every lane has read the list before lane 0 compacts it
        if (Co::leader()) {
            int w = 0;
            for (int o = 0; o < n_open; ++o)
                if (!(sc.open[o].query_end < c)) sc.open[w++] = sc.open[o];
        }
        n_open = keep;
        Co::sync();
    }

    // merge_hits_into_nams for one (strand, reference id) group of m sorted hits (src/nam.cpp:400-500)
    RSA_SEED_HD void merge_group(int m, int strand, int grp, uint32_t ref_id) {
        int n_open = 0;
        unsigned int prev_q_start = 0;
        for (int i = 0; i < m && !overflow; ++i) {
            const Hit h = sc.grouped[i];
            // the first open NAM (in list order) this hit extends (kind 1) or lies inside (kind 2)
            int found = -1, kind = 0;
            for (int base = 0; base < n_open; base += Co::kLanes) {
                const int oi = base + Co::lane();
                int k = 0;
                if (oi < n_open) {
                    const rsa_seed_nam_t o = sc.open[oi];
                    if (o.query_prev_hit_startpos < (int)h.qs && (int)h.qs <= o.query_end && o.ref_prev_hit_startpos < h.rs && h.rs <= o.ref_end) {
                        if ((int)h.qe > o.query_end && h.re > o.ref_end) k = 1;
                        else if ((int)h.qe <= o.query_end && h.re <= o.ref_end) k = 2;
                    }
                }
                const unsigned mask = Co::ballot(k != 0);
                if (mask) {
                    const int src = first_bit(mask);
                    found = base + src;
                    kind = Co::bcast(k, src);
                    break;
                }
            }
            if (found >= 0) {
                if (Co::leader()) {
                    rsa_seed_nam_t& o = sc.open[found];
                    if (kind == 1) { o.query_end = h.qe; o.ref_end = h.re; }
                    o.query_prev_hit_startpos = h.qs; o.ref_prev_hit_startpos = h.rs;
                    o.n_hits++;
                }
            } else {
                if (n_open >= caps.open) { overflow = true; return; }
                if (Co::leader()) sc.open[n_open] = new_nam(h, ref_id);
                ++n_open;
            }
            Co::sync();
            if ((unsigned int)h.qs > prev_q_start + (unsigned int)P.k) {
                flush_passed(n_open, (int)h.qs, strand, grp);
                prev_q_start = h.qs;
            }
        }
        for (int o = 0; o < n_open; ++o) push_nam(sc.open[o], strand, grp);
        Co::sync();
    }

    // merge_hits_into_nams_fast (sort == false) for one group (src/nam.cpp:170-366): hits are taken in runs of equal
    // query_start; each open NAM, in list order, looks (binary search on ref_start) at the run's hits inside
    // (ref_prev_hit_startpos, ref_end] and takes the first one not taken yet that extends it or lies inside it.
    // With several lanes, kLanes open NAMs pick their hit at once; two lanes picking the same hit means the later one would
    // have seen it taken, so only the lanes before the first such collision commit and the rest look again.
    RSA_SEED_HD void merge_group_fast(int m, int strand, int grp, uint32_t ref_id) {
        Hit* hits = sc.grouped;
        int n_open = 0;
        unsigned int prev_q_start = 0;
        for (int i = 0; i < m && !overflow;) {
            const int i_start = i;
            int i_end = i + 1;
            while (i_end < m && hits[i_end].qs == hits[i].qs) ++i_end;
            i = i_end;
            const int i_size = i_end - i_start;
            const int query_start = hits[i_start].qs;
            int cnt_done = 0;
            Co::sync();
            if (Co::leader()) {
                // std::sort(hits.begin() + i_start, hits.begin() + i_end): (query_start equal) by ref_start, unique
                for (int a = i_start + 1; a < i_end; ++a) {
                    const Hit x = hits[a];
                    int j = a - 1;
                    while (j >= i_start && hits[j].rs > x.rs) { hits[j + 1] = hits[j]; --j; }
                    hits[j + 1] = x;
                }
                for (int a = i_start; a < i_end; ++a) hits[a].pad = 0;  // is_added[]
            }
            Co::sync();
            for (int oi0 = 0; oi0 < n_open && cnt_done < i_size;) {
                const int oi = oi0 + Co::lane();
                int cand = -1, kind = 0;
                if (oi < n_open) {
                    const rsa_seed_nam_t o = sc.open[oi];
                    int lo = i_start, hi = i_end;  // lower_bound(ref_start >= value), twice
                    const int v1 = o.ref_prev_hit_startpos + 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (hits[mid].rs < v1) lo = mid + 1; else hi = mid; }
                    const int lower = lo;
                    lo = i_start; hi = i_end;
                    const int v2 = o.ref_end + 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (hits[mid].rs < v2) lo = mid + 1; else hi = mid; }
                    const int upper = lo;
                    if (query_start <= o.query_end) {
                        for (int j = lower; j < upper; ++j) {
                            const Hit h = hits[j];
                            if (h.pad) continue;
                            if (o.ref_prev_hit_startpos < h.rs && h.rs <= o.ref_end) {
                                if ((int)h.qe > o.query_end && h.re > o.ref_end) { cand = j; kind = 1; break; }
                                else if ((int)h.qe <= o.query_end && h.re <= o.ref_end) { cand = j; kind = 2; break; }
                            }
                        }
                    }
                }
                const int stop = Co::first_conflict(cand);  // lanes [0, stop) commit
                const bool commit = cand >= 0 && Co::lane() < stop;
                if (commit) {
                    const Hit h = hits[cand];
                    rsa_seed_nam_t& o = sc.open[oi];
                    if (kind == 1) { o.query_end = h.qe; o.ref_end = h.re; }
                    o.query_prev_hit_startpos = h.qs; o.ref_prev_hit_startpos = h.rs;
                    o.n_hits++;
                    hits[cand].pad = 1;
                }
                cnt_done += pop_count(Co::ballot(commit));
                oi0 += stop < Co::kLanes ? stop : Co::kLanes;
                Co::sync();
            }
            for (int a = i_start; a < i_end; ++a) {
                if (!hits[a].pad) {
                    if (n_open >= caps.open) { overflow = true; return; }
                    if (Co::leader()) sc.open[n_open] = new_nam(hits[a], ref_id);
                    ++n_open;
                }
            }
            Co::sync();
            if ((unsigned int)query_start > prev_q_start + (unsigned int)P.k) {
                flush_passed(n_open, query_start, strand, grp);
                prev_q_start = (unsigned int)query_start;
            }
        }
        for (int o = 0; o < n_open; ++o) push_nam(sc.open[o], strand, grp);
        Co::sync();
    }

    // all groups of the current strand -> NAMs (merge_hits_into_nams_forward_and_reverse[_fast], one strand)
    RSA_SEED_HD void merge_strand(int strand, bool fast) {
        Co::sync();  // the strand's hits are complete
        for (int g = 0; g < n_groups[strand] && !overflow; ++g) {
            const int m = gather_group(g);
            const uint32_t ref_id = sc.group_ref[strand * caps.groups + g];
            if (fast) merge_group_fast(m, strand, g, ref_id);
            else { sort_group(m); merge_group(m, strand, g, ref_id); }
        }
    }
};

// One read: find_nams, then find_nams_rescue when the reference would run it.  Returns the number of NAMs left in
// sc.nams (strand 0 groups in first-touch order, then strand 1), or -1 on scratch overflow.
template <class Co>
RSA_SEED_HD inline int seed_read(const uint8_t* seq, int len, const Index& ix, const Params& P, const Caps& caps, Scratch& sc,
                                 float& fraction, bool& rescued) {
    fraction = 1.0f;
    rescued = false;
    ReadCtx<Co> rc(ix, P, caps, sc);
    int n_syn = 0;
    if (len >= P.w_max) {  // randstrobes_query: `if (seq.length() < parameters.randstrobe.w_max) return` (src/randstrobes.cpp:209)
        n_syn = read_syncmers<Co>(seq, len, P, sc.syn_hash, sc.syn_pos, caps.syn);
        if (n_syn < 0) return -1;
    }
    Co::sync();
    const int n_rs = n_syn > P.w_min ? n_syn - P.w_min : 0;  // randstrobes per strand (has_next: idx + w_min < size)
    // ---- find_nams (src/nam.cpp:771-922)
    int total_hits = 0, good_hits = 0;
    for (int strand = 0; strand < 2; ++strand) {
        Syncmers S{sc.syn_hash, sc.syn_pos, n_syn, len, P.k, strand == 1};
        rc.n_hits = 0;
        for (int idx = 0; idx < n_rs; ++idx) {
            uint64_t hash; int qs, qe;
            randstrobe_at(S, P, idx, hash, qs, qe);
            const long long pos = index_find(ix, P, hash);
            if (pos < 0) continue;
            ++total_hits;
            if (index_is_filtered(ix, P, pos)) continue;
            ++good_hits;
            rc.emit_hits(strand, qs, qe, pos, false);
            if (rc.overflow) return -1;
        }
        rc.merge_strand(strand, false);
        if (rc.overflow) return -1;
    }
    fraction = total_hits > 0 ? (float)good_hits / (float)total_hits : 1.0f;
    // `nonrepetitive_fraction < 0.7` in the reference compares the float with the DOUBLE 0.7 (src/aln.cpp:1944): a fraction
    // of exactly 0.7f is smaller than it and rescues
    if (!(P.rescue_level > 1 && (rc.n_nams == 0 || (double)fraction < 0.7))) return rc.n_nams;

    // ---- find_nams_rescue (src/nam.cpp:955-1012, the pre_sort build)
    rescued = true;
    rc.n_nams = 0;
    rc.n_groups[0] = rc.n_groups[1] = 0;
    if (n_rs > caps.resc) return -1;
    int n_resc[2] = {0, 0};
    for (int strand = 0; strand < 2; ++strand) {
        Syncmers S{sc.syn_hash, sc.syn_pos, n_syn, len, P.k, strand == 1};
        RescueHit* rh = sc.resc + strand * caps.resc;
        for (int idx = 0; idx < n_rs; ++idx) {
            uint64_t hash; int qs, qe;
            randstrobe_at(S, P, idx, hash, qs, qe);
            const long long pos = index_find(ix, P, hash);
            if (pos < 0) continue;
            RescueHit r;
            r.position = (uint32_t)pos; r.count = index_get_count(ix, P, pos); r.qs = (uint16_t)qs; r.qe = (uint16_t)qe; r.pad = 0;
            if (Co::leader()) {
                // std::sort(..., cmp1): by (count, query_start, query_end), keys unique: insert in place
                int j = n_resc[strand] - 1;
                while (j >= 0 && (rh[j].count > r.count || (rh[j].count == r.count && (rh[j].qs > r.qs || (rh[j].qs == r.qs && rh[j].qe > r.qe))))) {
                    rh[j + 1] = rh[j];
                    --j;
                }
                rh[j + 1] = r;
            }
            n_resc[strand]++;
        }
    }
    Co::sync();
    for (int strand = 0; strand < 2; ++strand) {
        RescueHit* rh = sc.resc + strand * caps.resc;
        // the hits to use: in count order until the cutoff; the keys of hits_per_ref are created in THIS order (_pre)
        int n_use = 0;
        for (int a = 0; a < n_resc[strand]; ++a) {
            if ((rh[a].count > P.rescue_cutoff && n_use >= 5) || rh[a].count > 1000) break;
            rc.emit_hits(strand, rh[a].qs, rh[a].qe, (long long)rh[a].position, true);
            if (rc.overflow) return -1;
            ++n_use;
        }
        Co::sync();
        if (Co::leader()) {
            // std::sort(rhs, cmp2): by query_start (unique per strand)
            for (int a = 1; a < n_use; ++a) {
                const RescueHit x = rh[a];
                int j = a - 1;
                while (j >= 0 && rh[j].qs > x.qs) { rh[j + 1] = rh[j]; --j; }
                rh[j + 1] = x;
            }
        }
        Co::sync();
        rc.n_hits = 0;
        for (int a = 0; a < n_use; ++a) {
            rc.emit_hits(strand, rh[a].qs, rh[a].qe, (long long)rh[a].position, false);
            if (rc.overflow) return -1;
        }
        rc.merge_strand(strand, true);
        if (rc.overflow) return -1;
    }
    return rc.n_nams;
}

constexpr int kSeedThreads = 128;

// Grid-stride over the reads (or over `list`, the reads the small tier could not hold).  Co = CoThread: every thread owns
// one scratch slice and one read at a time; Co = CoWarp: every warp does.
template <class Co>
__global__ void __launch_bounds__(kSeedThreads)
seed_kernel(const uint8_t* __restrict__ reads, const int64_t* __restrict__ roff, const uint32_t* __restrict__ list, int n,
            Index ix, Params P, Caps caps, uint8_t* __restrict__ scratch, size_t scratch_stride,
            rsa_seed_read_t* __restrict__ per_read, rsa_seed_nam_t* __restrict__ nam_out, unsigned long long nam_cap,
            unsigned long long* __restrict__ counters /* [0] NAM cursor, [1] overflowed reads, [2] rescued, [3] retry list length */,
            uint32_t* __restrict__ retry_list, int final_tier) {
    const int tid = (int)((blockIdx.x * blockDim.x + threadIdx.x) / Co::kLanes);   // worker = thread or warp
    const int stride = (int)(gridDim.x * blockDim.x / Co::kLanes);
    Scratch sc(scratch + (size_t)tid * scratch_stride, caps);
    for (int it = tid; it < n; it += stride) {
        const uint32_t r = list ? list[it] : (uint32_t)it;
        const int64_t off = roff[r];
        const int len = (int)(roff[r + 1] - off);
        float fraction;
        bool rescued;
        Co::sync();  // the previous read's lists are dead
        const int cnt = seed_read<Co>(reads + off, len, ix, P, caps, sc, fraction, rescued);
        Co::sync();
        rsa_seed_read_t pr;
        pr.nam_off = 0; pr.n_nams = 0; pr.nonrepetitive_fraction = fraction; pr.flags = rescued ? RSA_SEED_READ_RESCUED : 0u;
        if (cnt < 0) {
            if (Co::leader()) {
                pr.flags = RSA_SEED_READ_FAILED;
                if (final_tier) atomicAdd(&counters[1], 1ull);
                else { const unsigned long long slot = atomicAdd(&counters[3], 1ull); retry_list[slot] = r; }
                per_read[r] = pr;
            }
            continue;
        }
        unsigned long long at = 0;
        if (Co::leader()) {
            if (rescued) atomicAdd(&counters[2], 1ull);
            at = atomicAdd(&counters[0], (unsigned long long)cnt);
            pr.nam_off = (uint32_t)at;
            pr.n_nams = cnt;
            per_read[r] = pr;
        }
        const unsigned lo = (unsigned)Co::bcast((int)(uint32_t)(at & 0xFFFFFFFFull), 0), hi = (unsigned)Co::bcast((int)(uint32_t)(at >> 32), 0);
        at = ((unsigned long long)hi << 32) | lo;
        if (at + (unsigned long long)cnt <= nam_cap)
            for (int i = Co::lane(); i < cnt; i += Co::kLanes) nam_out[at + i] = sc.nams[i];
    }
}

}  // namespace rsaseed
