// integration/seed_glue.cpp -- see seed_glue.hpp.  Compiled with the reference's headers (-I$REF_ROOT/src -I$REF_ROOT/ext).
#include "seed_glue.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "robin_hood.h"
#include "rsa_ext.h"
#include "rsa_seed.h"

namespace rsa_glue {
namespace {

constexpr int kMaxDevices = 64;

struct DeviceIndex {
    std::mutex m;
    rsa_seed_index_t* ix[kMaxDevices] = {};
    const StrobemerIndex* source = nullptr;
} g_index;

// per worker thread: the handle, the chunk's results (pinned buffers of the handle) and the read cursor
struct WorkerState {
    rsa_seed_t* h = nullptr;
    const rsa_seed_read_t* per = nullptr;
    const rsa_seed_nam_t* nams = nullptr;
    size_t n_reads = 0, cursor = 0;
    std::vector<const std::string*> seqs;
    std::vector<char> buf;
    std::vector<int64_t> off;
    // state between the three stand-in calls of one read
    bool host_path = false;            // the GPU could not seed this read (RSA_SEED_READ_FAILED): reference code runs
    bool have_rescue = false;
    std::vector<Nam> rescue_nams;
    ~WorkerState() { if (h) rsa_seed_destroy(h); }
};
thread_local WorkerState t_state;

[[noreturn]] void die(const char* what, const rsa_seed_t* h) {
    fprintf(stderr, "[RSA_SEED ERROR:] %s: %s\n", what, rsa_seed_last_error(h));
    exit(EXIT_FAILURE);
}

int usable_devices() {
    int ndev = rsa_ext_device_count();
    const char* e = getenv("RSA_EXT_DEVICES");
    const int cap = e ? atoi(e) : 0;
    return (cap > 0 && cap < ndev) ? cap : ndev;
}

rsa_seed_index_t* device_index(int device, const IndexParameters& ip, const StrobemerIndex& index, const MappingParameters& mp) {
    std::lock_guard<std::mutex> lk(g_index.m);
    if (g_index.source && g_index.source != &index) {
        fprintf(stderr, "[RSA_SEED ERROR:] one index per process\n");
        exit(EXIT_FAILURE);
    }
    g_index.source = &index;
    if (!g_index.ix[device]) {
        rsa_seed_config_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.device = device;
        cfg.k = ip.syncmer.k; cfg.s = ip.syncmer.s; cfg.t_syncmer = ip.syncmer.t_syncmer;
        cfg.w_min = (int32_t)ip.randstrobe.w_min; cfg.w_max = (int32_t)ip.randstrobe.w_max; cfg.max_dist = ip.randstrobe.max_dist;
        cfg.q = ip.randstrobe.q;
        cfg.bits = index.get_bits();
        cfg.filter_cutoff = index.filter_cutoff;
        cfg.rescue_level = mp.rescue_level;
        cfg.rescue_cutoff = (uint32_t)mp.rescue_cutoff;
        static_assert(sizeof(RefRandstrobe) == 16, "RefRandstrobe layout");
        if (rsa_seed_index_upload(&cfg, index.randstrobes.data(), (int64_t)index.randstrobes.size(),
                                  index.randstrobe_start_indices.data(), (int64_t)index.randstrobe_start_indices.size(),
                                  &g_index.ix[device]) != RSA_SEED_OK)
            die("rsa_seed_index_upload", nullptr);
    }
    return g_index.ix[device];
}

// NAMs of read r in the reference's order, nam_id assigned
void build_nams(const WorkerState& w, size_t r, std::vector<Nam>& out) {
    const rsa_seed_read_t& pr = w.per[r];
    const rsa_seed_nam_t* src = w.nams + pr.nam_off;
    const int n = pr.n_nams;
    out.clear();
    out.reserve((size_t)n);
    auto push = [&](const rsa_seed_nam_t& g) {
        Nam nam;
        nam.nam_id = (int)out.size();
        nam.query_start = g.query_start; nam.query_end = g.query_end; nam.query_prev_hit_startpos = g.query_prev_hit_startpos;
        nam.ref_start = g.ref_start; nam.ref_end = g.ref_end; nam.ref_prev_hit_startpos = g.ref_prev_hit_startpos;
        nam.n_hits = g.n_hits; nam.ref_id = g.ref_id; nam.score = g.score; nam.is_rc = (g.flags & 1u) != 0;
        out.push_back(nam);
    };
    int i = 0;
    for (int strand = 0; strand < 2; ++strand) {
        const int lo = i;
        uint32_t max_group = 0;
        while (i < n && (int)(src[i].flags & 1u) == strand) { max_group = std::max(max_group, src[i].flags >> 8); ++i; }
        if (i == lo) continue;
        if (max_group == 0) {  // one reference sequence on this strand: nothing to re-order
            for (int k = lo; k < i; ++k) push(src[k]);
            continue;
        }
        // several: the reference's container decides (src/nam.cpp:775-777: reserve(100), keys inserted on first touch)
        struct HitLike { int a, b, c, d; };
        robin_hood::unordered_map<unsigned int, std::vector<HitLike>> order;
        order.reserve(100);
        std::vector<int> first_of_group(max_group + 1, -1);
        for (int k = lo; k < i; ++k) {
            const uint32_t g = src[k].flags >> 8;
            if (first_of_group[g] < 0) first_of_group[g] = k;
        }
        for (uint32_t g = 0; g <= max_group; ++g)
            if (first_of_group[g] >= 0) order[(unsigned int)src[first_of_group[g]].ref_id];
        for (auto& kv : order)
            for (int k = lo; k < i; ++k)
                if ((unsigned int)src[k].ref_id == kv.first) push(src[k]);
    }
}

}  // namespace

void seed_chunk(int thread_id, const std::vector<const std::string*>& seqs, const IndexParameters& ip, const StrobemerIndex& index,
                const MappingParameters& mp) {
    WorkerState& w = t_state;
    if (!w.h) {
        const int ndev = usable_devices();
        const int device = ndev > 0 ? thread_id % ndev : 0;
        if (device >= kMaxDevices) die("device ordinal", nullptr);
        rsa_seed_index_t* ix = device_index(device, ip, index, mp);
        if (rsa_seed_create(ix, &w.h) != RSA_SEED_OK) die("rsa_seed_create", nullptr);
    }
    w.seqs = seqs;
    w.n_reads = seqs.size();
    w.cursor = 0;
    w.per = nullptr;
    w.nams = nullptr;
    if (seqs.empty()) return;
    size_t bytes = 0;
    w.off.resize(seqs.size() + 1);
    for (size_t i = 0; i < seqs.size(); ++i) { w.off[i] = (int64_t)bytes; bytes += seqs[i]->size(); }
    w.off[seqs.size()] = (int64_t)bytes;
    w.buf.resize(bytes + 16);
    for (size_t i = 0; i < seqs.size(); ++i) memcpy(w.buf.data() + w.off[i], seqs[i]->data(), seqs[i]->size());
    int64_t n_nams = 0;
    if (rsa_seed_find_nams(w.h, (int64_t)seqs.size(), w.buf.data(), w.off.data(), &w.per, &w.nams, &n_nams) != RSA_SEED_OK)
        die("rsa_seed_find_nams", w.h);
}

QueryRandstrobeVector randstrobes_query(const std::string_view seq, const IndexParameters& parameters) {
    WorkerState& w = t_state;
    if (w.cursor >= w.n_reads || !w.per) {
        fprintf(stderr, "[RSA_SEED ERROR:] read %zu of a chunk of %zu was not seeded (seed_chunk missing before the loop?)\n", w.cursor, w.n_reads);
        exit(EXIT_FAILURE);
    }
    w.host_path = (w.per[w.cursor].flags & RSA_SEED_READ_FAILED) != 0;
    w.have_rescue = false;
    // the GPU already consumed the randstrobes; the host needs them only for the rare read it has to seed itself
    return w.host_path ? ::randstrobes_query(seq, parameters) : QueryRandstrobeVector();
}

std::pair<float, std::vector<Nam>> find_nams(const QueryRandstrobeVector& query_randstrobes, const StrobemerIndex& index) {
    WorkerState& w = t_state;
    const size_t r = w.cursor++;
    if (w.host_path) return ::find_nams(query_randstrobes, index);
    const rsa_seed_read_t& pr = w.per[r];
    std::vector<Nam> nams;
    build_nams(w, r, nams);
    if (pr.flags & RSA_SEED_READ_RESCUED) {
        // the device ran find_nams_rescue because the first pass came back empty or repetitive; hand the caller an empty
        // first-pass list so that its own `nams.empty() || fraction < 0.7` test takes the rescue branch (and counts it)
        w.rescue_nams = std::move(nams);
        w.have_rescue = true;
        return {pr.nonrepetitive_fraction, std::vector<Nam>()};
    }
    return {pr.nonrepetitive_fraction, std::move(nams)};
}

std::vector<Nam> find_nams_rescue(const QueryRandstrobeVector& query_randstrobes, const StrobemerIndex& index, unsigned int rescue_cutoff) {
    WorkerState& w = t_state;
    if (w.host_path) return ::find_nams_rescue(query_randstrobes, index, rescue_cutoff);
    if (!w.have_rescue) {
        fprintf(stderr, "[RSA_SEED ERROR:] the caller asked for rescue NAMs the device did not compute\n");
        exit(EXIT_FAILURE);
    }
    w.have_rescue = false;
    return std::move(w.rescue_nams);
}

}  // namespace rsa_glue
