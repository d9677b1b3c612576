// tests/seed_host_check.cu -- TEST INFRASTRUCTURE.  The per-read seeding code of the product
// (rabbitsalign_b200/csrc/kernels_seed.cuh: the functions the CUDA kernel calls) compiled for the HOST, so that its logic
// can be checked against the reference's seeding path (oracle/_ref/libseed_ref.so) in the `-m "not gpu"` suite, where no
// GPU exists.  Not part of librsa_ext.so; the product library has no such path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../rabbitsalign_b200/csrc/kernels_seed.cuh"

using namespace rsaseed;

extern "C" int64_t seed_host_check(const rsa_seed_config_t* cfg, const void* entries, int64_t n_entries, const uint64_t* starts,
                                   int64_t n_reads, const char* reads, const int64_t* roff, int large_tier,
                                   rsa_seed_read_t* per_read, rsa_seed_nam_t* nams, int64_t nam_cap) {
    Params P;
    P.k = cfg->k; P.s = cfg->s; P.t_syncmer = cfg->t_syncmer; P.w_min = cfg->w_min; P.w_max = cfg->w_max; P.max_dist = cfg->max_dist;
    P.bits = cfg->bits; P.rescue_level = cfg->rescue_level; P.filter_cutoff = cfg->filter_cutoff; P.rescue_cutoff = cfg->rescue_cutoff;
    P.q = cfg->q; P.n_entries = n_entries;
    const Index I{static_cast<const IndexEntry*>(entries), starts, (long long)n_entries};
    const Caps small{128, 384, 16, 48, 128, 128}, large{512, 32768, 512, 8192, 16384, 512};
    const Caps caps = large_tier ? large : small;
    std::vector<uint8_t> buf(scratch_bytes(caps) + 64);
    Scratch sc(buf.data(), caps);
    int64_t total = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        float fraction;
        bool rescued;
        const int cnt = seed_read<CoThread>(reinterpret_cast<const uint8_t*>(reads) + roff[r], (int)(roff[r + 1] - roff[r]), I, P, caps, sc,
                                  fraction, rescued);
        rsa_seed_read_t pr;
        pr.nam_off = (uint32_t)total; pr.n_nams = 0; pr.nonrepetitive_fraction = fraction;
        pr.flags = rescued ? RSA_SEED_READ_RESCUED : 0u;
        if (cnt < 0) pr.flags = RSA_SEED_READ_FAILED;
        else {
            if (total + cnt > nam_cap) return -1;
            pr.n_nams = cnt;
            memcpy(nams + total, sc.nams, sizeof(rsa_seed_nam_t) * (size_t)cnt);
            total += cnt;
        }
        per_read[r] = pr;
    }
    return total;
}

// the query randstrobes of one read, in the reference's order (forward strand, then reverse): n x {hash, start, end, is_reverse}
extern "C" int64_t seed_host_randstrobes(const rsa_seed_config_t* cfg, const char* seq, int64_t len, uint64_t* out, int64_t cap) {
    Params P;
    P.k = cfg->k; P.s = cfg->s; P.t_syncmer = cfg->t_syncmer; P.w_min = cfg->w_min; P.w_max = cfg->w_max; P.max_dist = cfg->max_dist;
    P.bits = cfg->bits; P.q = cfg->q;
    std::vector<uint64_t> sh(4096);
    std::vector<int32_t> sp(4096);
    if (len < P.w_max) return 0;
    const int n_syn = read_syncmers<CoThread>(reinterpret_cast<const uint8_t*>(seq), (int)len, P, sh.data(), sp.data(), 4096);
    if (n_syn < 0) return -2;
    const int n_rs = n_syn > P.w_min ? n_syn - P.w_min : 0;
    int64_t n = 0;
    for (int strand = 0; strand < 2; ++strand) {
        Syncmers S{sh.data(), sp.data(), n_syn, (int)len, P.k, strand == 1};
        for (int idx = 0; idx < n_rs; ++idx) {
            uint64_t hash; int qs, qe;
            randstrobe_at(S, P, idx, hash, qs, qe);
            if (n >= cap) return -1;
            out[4 * n] = hash; out[4 * n + 1] = (uint64_t)qs; out[4 * n + 2] = (uint64_t)qe; out[4 * n + 3] = (uint64_t)strand;
            ++n;
        }
    }
    return n;
}
