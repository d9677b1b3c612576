// kernels_plan.cuh -- chunk planning on the device.
//
// What the host planner (plan_chunk in engine.cu) computes with a two-pass counting sort on ONE host thread -- per-pair
// records, routing (packed / exact / not aligned), the (|q|, |t|) ordering of the packed kernel's candidates, their
// pairing into groups, every direction-tile offset -- done by six small kernels from the batch's raw offset arrays.
// The host ships 16 bytes per pair (its offsets) instead of a 45-byte-per-pair blob, and only counts pairs per query
// length (one streaming pass) so that it knows the launch geometry.  The reference has no counterpart: GASAL2 gives
// every pair the worst-case tile (GASAL2/src/ctors.cpp:112-115) and launches in arrival order.
//
//   plan_classify   one thread per pair: PairMeta / info / routing; packed candidates are counted into a
//                   (|q|, ceil4(|t|)/4) histogram (atomicAdd returns the pair's rank inside its bin), exact-routed
//                   pairs take their list slot and direction tile from atomic cursors;
//   plan_bin_sums + plan_bin_scan   exclusive scan of the histogram (two coalesced passes over 1 MB);
//   plan_scatter    sorted[bin_start + rank] = pair;
//   plan_groups     consecutive equal-|q| pairs of the sorted order become one group (A = low halves, B = high halves);
//   plan_offsets    exclusive scan of the groups' tile sizes -> FastGroup::dir_off, diroff[], the redo region's base.
//
// The order inside one histogram bin is the arrival order of the atomics, i.e. not reproducible -- but pairs of one bin
// have the same |q| and the same number of 4-row blocks, so group shapes, tile offsets and every result are.
#pragma once
#include "common.cuh"
#include "fast_layout.cuh"
#include "kernels_exact.cuh"
#include "kernels_fast.cuh"

namespace rsa {

constexpr int kPlanQ = kFastMaxQlen + 1;                  // |q| 0..512
constexpr int kPlanT4 = (kFastMaxTlen + 3) / 4 + 1;       // ceil4(|t|)/4 in 0..512
constexpr int kPlanBins = kPlanQ * kPlanT4;
constexpr int kPlanScanBlock = 1024;                      // bins per block of the two scan passes
constexpr int kPlanScanBlocks = (kPlanBins + kPlanScanBlock - 1) / kPlanScanBlock;
constexpr uint32_t kPlanNone = 0xFFFFFFFFu;

// Per query length: what the host's counting pass knows (uploaded with the chunk, 12 B x 513)
struct PlanQlen {
    uint32_t count;       // packed-kernel candidates of this |q| in the chunk
    uint32_t pos_base;    // first position of this |q| in the sorted order
    uint32_t group_base;  // first group slot of this |q|
};

// Device-side counters/results of one chunk's planning (zeroed before plan_classify)
struct PlanHeader {
    unsigned long long exact_bytes;  // direction bytes handed to statically exact-routed pairs (atomic cursor)
    unsigned long long redo_base;    // first byte of the redo head-room: align256(exact tiles + packed tiles)
    unsigned int exact_count[3];     // list cursors of the three exact classes
    unsigned int pad;
};

struct PlanArgs {
    // inputs (device copies of the caller's arrays, chunk slice)
    const int64_t* qoff;     // n + 1
    const int64_t* toff;     // n + 1, or nullptr in window form
    const int64_t* win_off;  // n   (window form)
    const int32_t* win_len;  // n
    const PlanQlen* qtab;    // kPlanQ entries
    int n;
    int max_tlen;
    int match;
    int exact_only;
    unsigned int list_base[3];
    // outputs
    PairMeta* meta;
    uint32_t* info;
    uint64_t* diroff;
    uint32_t* list;
    FastGroup* groups;
    // temporaries
    uint32_t* key;      // n: histogram bin of a packed candidate, kPlanNone otherwise
    uint32_t* rank;     // n
    uint32_t* sorted;   // n
    uint32_t* gbytes;   // one per group slot: tile bytes (0 for padding slots)
    uint32_t* hist;     // kPlanBins
    uint32_t* bsum;     // kPlanScanBlocks
    PlanHeader* hdr;
};

__device__ __forceinline__ bool plan_fast_shape_ok(int qlen, int tlen, int match) {
    return qlen >= kFastMinQlen && qlen <= kFastMaxQlen && tlen >= 1 && tlen <= kFastMaxTlen && fast_key_ok(qlen, match);
}

constexpr int kPlanThreads = 256;

__global__ void __launch_bounds__(kPlanThreads) plan_classify(PlanArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int64_t q0 = a.qoff[0];
    const int64_t qo = a.qoff[i];
    const int64_t ql = a.qoff[i + 1] - qo;
    int64_t to, tl;
    if (a.toff) { to = a.toff[i] - a.toff[0]; tl = a.toff[i + 1] - a.toff[i]; }
    else { to = a.win_off[i]; tl = a.win_len[i]; }
    PairMeta m;
    m.qoff = (uint32_t)(qo - q0);
    m.toff = (uint32_t)to;
    m.qlen = (uint16_t)ql;
    m.tlen = (uint16_t)(tl < 65535 ? tl : 65535);
    a.meta[i] = m;
    a.diroff[i] = 0;
    uint32_t info = 0, key = kPlanNone;
    if (ql == 0 || tl == 0) info = 3u << 16;
    else if (tl > a.max_tlen) info = 1u << 16;
    else if (!a.exact_only && plan_fast_shape_ok((int)ql, (int)tl, a.match)) {
        key = (uint32_t)ql * kPlanT4 + (uint32_t)((tl + 3) >> 2);
        a.rank[i] = atomicAdd(&a.hist[key], 1u);
    } else {
        const int rb = exact_row_bytes((int)ql);
        const int cls = rb == 64 ? 0 : (rb == 128 ? 1 : 2);
        const unsigned int slot = atomicAdd(&a.hdr->exact_count[cls], 1u);
        a.list[a.list_base[cls] + slot] = (uint32_t)i;
        const unsigned long long bytes = ((unsigned long long)tl * rb + 15ull) & ~15ull;
        a.diroff[i] = atomicAdd(&a.hdr->exact_bytes, bytes);
    }
    a.info[i] = info;
    a.key[i] = key;
}

// block b sums bins [b * 1024, (b + 1) * 1024)
__global__ void __launch_bounds__(256) plan_bin_sums(PlanArgs a) {
    __shared__ uint32_t wsum[8];
    const int base = blockIdx.x * kPlanScanBlock + threadIdx.x * 4;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (base + k < kPlanBins) ? a.hist[base + k] : 0u;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        a.bsum[blockIdx.x] = t;
    }
}

// hist[bin] <- number of candidates in all smaller bins (exclusive scan, in place)
__global__ void __launch_bounds__(256) plan_bin_scan(PlanArgs a) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t block_base;
    // this block's base: sum of the block sums before it
    uint32_t pre = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 256) pre += a.bsum[b];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) pre += __shfl_xor_sync(0xFFFFFFFFu, pre, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = pre;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        block_base = t;
    }
    __syncthreads();
    const int base = blockIdx.x * kPlanScanBlock + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < kPlanBins) ? a.hist[base + k] : 0u; s += v[k]; }
    // exclusive scan of the per-thread sums across the block
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if ((int)(threadIdx.x & 31) >= off) incl += o;
    }
    __syncthreads();  // wsum is reused
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
    uint32_t run = block_base + wbase + incl - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < kPlanBins) a.hist[base + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(kPlanThreads) plan_scatter(PlanArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t key = a.key[i];
    if (key == kPlanNone) return;
    a.sorted[a.hist[key] + a.rank[i]] = (uint32_t)i;
}

// one thread per sorted position; the even positions of each |q| run build the groups
__global__ void __launch_bounds__(kPlanThreads) plan_groups(PlanArgs a, int n_fast) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_fast) return;
    const uint32_t i = a.sorted[p];
    const PairMeta mi = a.meta[i];
    const PlanQlen qt = a.qtab[mi.qlen];
    const uint32_t k = (uint32_t)p - qt.pos_base;
    if (k & 1u) return;
    uint32_t j = i;
    uint32_t rows = mi.tlen;
    if (k + 1 < qt.count) {
        j = a.sorted[p + 1];
        rows = max(rows, (uint32_t)a.meta[j].tlen);
    }
    const uint32_t g = qt.group_base + (k >> 1);
    FastGroup fg;
    fg.a = i; fg.b = j; fg.dir_off = 0; fg.qlen = mi.qlen; fg.rows = (uint16_t)rows;
    a.groups[g] = fg;
    const FastGeom geo = fast_geom(mi.qlen);
    a.gbytes[g] = (uint32_t)fast_dir_bytes(geo, (int)rows);
}

// exclusive scan of the group tile sizes (one block), then the offsets every consumer reads
__global__ void __launch_bounds__(1024) plan_offsets(PlanArgs a, int n_groups) {
    __shared__ unsigned long long wsum[32];
    const int per = (n_groups + 1023) / 1024;
    const int g0 = threadIdx.x * per;
    unsigned long long s = 0;
    for (int g = g0; g < min(n_groups, g0 + per); ++g) s += a.gbytes[g];
    unsigned long long incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if ((int)(threadIdx.x & 31) >= off) incl += o;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) { if (w < (int)(threadIdx.x >> 5)) wbase += wsum[w]; total += wsum[w]; }
    const unsigned long long fast_base = (a.hdr->exact_bytes + 255ull) & ~255ull;  // exact tiles first, then the packed tiles
    unsigned long long run = fast_base + wbase + incl - s;
    for (int g = g0; g < min(n_groups, g0 + per); ++g) {
        const uint32_t bytes = a.gbytes[g];
        if (bytes) {
            FastGroup fg = a.groups[g];
            fg.dir_off = run;
            a.groups[g] = fg;
            a.diroff[fg.a] = run;
            if (fg.b != fg.a) { a.diroff[fg.b] = run; a.info[fg.b] |= 1u; }
        }
        run += bytes;
    }
    if (threadIdx.x == 0) a.hdr->redo_base = (fast_base + total + 255ull) & ~255ull;
}

}  // namespace rsa
