// engine.cu -- host side of librsa_ext.so: the C ABI of include/rsa_ext.h.
//
// Replaces, for RabbitSAlign's extension step, the GASAL2 host runtime the reference drives from
// src/gasal2_ssw.cpp:19-256: storage/stream construction (GASAL2/src/ctors.cpp:26-128), pinned batch
// filling (GASAL2/src/host_batch.cpp:79-153), the launch path gasal_aln_async
// (GASAL2/src/gasal_align.cu:29-307) and completion polling (:310-326).
//
// Shape of the engine (B200-first, not GASAL2's):
//   * a batch of any size is cut into chunks whose direction-bit scratch fits the handle's budget;
//   * three chunk slots rotate over five streams (H2D, two DP streams used alternately, a high-priority traceback
//     stream, D2H) so the copies of chunk k+1 overlap the kernels of chunk k and the traceback of chunk k runs beside
//     the DP kernel of chunk k+1; per chunk there is ONE metadata blob copy, two sequence copies in and one
//     64-byte-record copy out (the reference issues ~13 small copies per 512 pairs);
//   * the host plans each chunk (length classes, equal-length pairing for the packed kernel, scratch offsets) while
//     the GPU works on the previous ones -- for large batches on a per-handle helper thread that runs ahead of the
//     caller's thread (PlanAhead), for small ones inline;
//   * kernels: packed s16x2 DPX wavefront kernel (kernels_fast.cuh) for the bulk, exact int32 wavefront
//     kernel (kernels_exact.cuh) for whatever the packed kernel declines, one traceback kernel
//     (kernels_tb.cuh).  No CPU fallback exists here.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "fast_layout.cuh"
#include "kernels_exact.cuh"
#include "kernels_fast.cuh"
#include "kernels_tb.cuh"
#include "kernels_finish.cuh"
#include "kernels_hamming.cuh"
#include "kernels_plan.cuh"

using namespace rsa;

namespace {

thread_local std::string g_create_error;
// Cold-path gate.  Context creation, stream/event creation, first-use module loading and cudaMalloc/cudaHostAlloc all
// take process-wide driver locks; when the reference's 16 workers make their first call at the same moment the
// waiters spin on those locks and the holder slows down ~10x (measured: 2.06 s instead of 0.2 s for the primary
// context).  Handles therefore take this mutex (sleeping waiters) for rsa_ext_create and for their first submit.
std::mutex g_cold_mutex;
const std::chrono::steady_clock::time_point g_t0 = std::chrono::steady_clock::now();  // library load (trace origin)
double since_load_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count(); }

constexpr int kSlots = 3;  // chunks in flight: one computing, one with its copies in flight, one being planned/retired
constexpr int64_t kMaxChunkPairsDefault = 1 << 17;
// tuning knob (experiments): RSA_EXT_MAX_CHUNK_PAIRS overrides the chunk size cap
inline int64_t max_chunk_pairs() {
    static const int64_t v = [] {
        const char* e = getenv("RSA_EXT_MAX_CHUNK_PAIRS");
        const long long x = e ? atoll(e) : 0;
        return (int64_t)(x >= 256 ? x : kMaxChunkPairsDefault);
    }();
    return v;
}
#define kMaxChunkPairs (max_chunk_pairs())
constexpr int64_t kMaxChunkSeqBytes = (int64_t)1 << 30;
constexpr int64_t kDefaultScratch = (int64_t)12 << 30;  // all slots together (allocated lazily, per slot, as needed)
constexpr int kMaxTargetLenCap = 8192;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DevBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
};
struct PinBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
};

// Host-side description of one planned chunk: where each section lives inside the metadata blob and how
// many tasks each kernel gets.
struct ChunkPlan {
    int64_t lo = 0, hi = 0;  // pair range of the batch
    int64_t n = 0;
    size_t off_meta = 0, off_info = 0, off_diroff = 0, off_list = 0, off_groups = 0, off_redo = 0, blob_bytes = 0;
    int n_exact[3] = {0, 0, 0};     // tasks per exact class (C = 4, 8, 16), consecutive in `list`
    int max_tlen_exact[3] = {0, 0, 0};
    int n_fast_classes = 0;
    struct FastClass { int L; int C; int group_begin; int n_groups; int max_tlen; bool flex = false; };
    std::vector<FastClass> fast;
    int64_t n_fast_pairs = 0;
    int64_t n_failed = 0;
    uint64_t scratch_bytes = 0;
    uint64_t arena_bytes = 0;
    int64_t q_bytes = 0, t_bytes = 0;
    int64_t cells = 0;
    // device-planned chunks (kernels_plan.cuh): where the planner's inputs and temporaries live in the device blob
    bool dev = false;
    size_t off_in_q = 0, off_in_t = 0, off_in_wlen = 0, off_qtab = 0, off_hdr = 0, off_gbytes = 0, off_bsum = 0, off_hist = 0,
           off_key = 0, off_rank = 0, off_sorted = 0, zero_bytes = 0;
    int n_group_slots = 0;
    unsigned int list_base[3] = {0, 0, 0};
};

struct Slot {
    PinBuf h_blob;
    unsigned long long* h_arena_used = nullptr;  // pinned; [0] arena bytes used, [1] pairs the redo pass could not place
    DevBuf d_slab;                                            // one allocation; the views below are carved from it
    DevBuf d_blob, d_q, d_t, d_ends, d_res, d_arena, d_aln;   // views into d_slab (ensure_slot)
    DevBuf d_scratch;                                         // direction tiles, own allocation
    unsigned long long* d_arena_used = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_mid = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    cudaEvent_t ev_in = nullptr, ev_plan = nullptr;  // device planner: offsets copied / planning kernels done
    ChunkPlan plan;
    bool busy = false;
};

// bytes a chunk needs from a slot (ensure_slot)
struct SlotNeed { size_t blob, q, t, ends, res, arena, aln, scratch; };

// Plan-ahead ring (large batches): a helper thread plans chunk after chunk into these entries while the caller's
// thread only enqueues and retires; an entry is reused once its chunk has retired.
struct PlannedChunk {
    ChunkPlan plan;
    PinBuf blob;
    int rc = 0;
};
constexpr int kPlanRing = kSlots + 2;
constexpr int kSlotCounters = 4;  // per-slot device counters: arena bytes, pairs the redo pass could not place, pairs redone,
                                  // pairs NO kernel produced (a lost launch: reported as an error)
constexpr int64_t kPlanAheadMinPairs = 32768;  // below this the caller's thread plans inline (no thread hop)
struct PlanAhead {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    bool stop = false, active = false, cancel = false;
    int64_t lo = 0;
    long produced = 0, consumed = 0, released = 0;
    double plan_ms = 0;
    PlannedChunk ring[kPlanRing];
};

// Reference sequence resident in HBM (rsa_ext_set_reference); handles of one device may share one copy
// (rsa_ext_share_reference), e.g. the 16 workers of the pipeline.
struct RefBuf {
    int device = 0;
    uint8_t* d = nullptr;
    const char* host = nullptr;  // kept for the rare exact-only re-submission (status 4); the caller keeps it alive
    int64_t len = 0;
    // packed form for the DP kernel's vectorised staging (pack_reference_kernel): 64-base units, 16 B of 2-bit codes and
    // 8 B of "not ACGT" flags each
    uint4* d_pack = nullptr;
    uint2* d_flag = nullptr;
    ~RefBuf() {
        cudaSetDevice(device);
        if (d) cudaFree(d);
        if (d_pack) cudaFree(d_pack);
        if (d_flag) cudaFree(d_flag);
    }
};

struct ResidentChunk {
    ChunkPlan plan;
    uint8_t* d_blob = nullptr;
    size_t q_base = 0, t_base = 0;  // byte offsets into the resident sequence buffers
};

}  // namespace

// A chunk of variable-length reads has one launch per column class; classes are independent, so each gets its own stream
// (small classes are latency-bound single waves: back to back on few streams they cost ~0.3 ms each)
constexpr int kClsStreams = 7;

struct rsa_ext {
    rsa_ext_config_t cfg{};
    Scoring sc{};
    FastConsts fk{};
    bool fast_ok = false;
    int n_sms = 148;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_comp2 = nullptr, s_tb = nullptr, s_d2h = nullptr, s_plan = nullptr;
    cudaStream_t s_cls[kClsStreams] = {};  // side streams for the column classes of one chunk
    cudaEvent_t ev_cls_fork = nullptr, ev_cls_join[kClsStreams] = {};
    cudaEvent_t ev_fork = nullptr;  // orders the second DP stream behind what the caller put on the first
    int dp_toggle = 0;              // consecutive chunks alternate between the two DP streams so that the next
                                    // chunk's blocks fill the SMs while the previous kernel's last wave drains
    Slot slots[kSlots];
    size_t scratch_per_slot = 0;

    // pending batch
    bool pending = false;
    int deferred_rc = 0;  // error met inside rsa_ext_poll, reported by the following rsa_ext_wait
    bool resident_inflight = false;
    SlotNeed r_need{};
    bool warmed = false;  // first submit done (buffers allocated, kernels loaded): no cold-path gate any more
    int64_t n = 0;
    const char* qbuf = nullptr;
    const char* tbuf = nullptr;
    const int64_t* qoff = nullptr;
    const int64_t* toff = nullptr;
    rsa_ext_result_t* results = nullptr;
    const int64_t* win_off = nullptr;      // pending batch in window form (targets inside the resident reference)
    const int32_t* win_len = nullptr;
    std::shared_ptr<RefBuf> ref;           // resident reference (rsa_ext_set_reference / rsa_ext_share_reference)
    rsa_ext_alninfo_t* alninfo = nullptr;  // optional second output of the pending batch
    rsa_ext_alninfo_t* alninfo_next = nullptr;
    int end_bonus = 10;
    int64_t next_pair = 0;
    int head = 0, tail = 0, inflight = 0;
    int chunks_enqueued = 0;
    PlanAhead* pa = nullptr;  // created by the first large submit
    bool plan_ahead = false;  // the pending batch is planned by the helper thread
    bool dev_plan = false;    // the pending batch is planned on the device (kernels_plan.cuh)
    std::unordered_map<int64_t, std::vector<uint8_t>> overflow;
    std::vector<int64_t> retry;  // pairs whose redo found no scratch (status 4): re-run exact-only at wait()

    // submit_ptrs staging
    PinBuf own_q, own_t;
    std::vector<int64_t> own_qoff, own_toff;

    DevBuf ham_q, ham_t, ham_off, ham_out;  // rsa_ext_hamming_* staging
    PinBuf ham_pin_in, ham_pin_out;         // pinned bounce buffers for callers that pass pageable memory
    cudaEvent_t ham_ev[2] = {nullptr, nullptr};

    // resident set
    std::vector<ResidentChunk> res_chunks;
    DevBuf r_q, r_t, r_res, r_blobs;
    int64_t r_n = 0;
    std::vector<cudaEvent_t> r_events;  // 4 per chunk: DP begin/end (compute stream), traceback begin/end (tb stream)
    bool r_events_valid = false;

    rsa_ext_stats_t stats{};
    std::string err;
    std::vector<uint32_t> tmp_list[3];
    std::vector<uint32_t> tmp_sort, tmp_order, tmp_key, tmp_key2;
    std::vector<uint32_t> tmp_count;
};

namespace {

#define CU_TRY(h, call)                                                                               \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return RSA_EXT_ERR_CUDA;                                                                  \
        }                                                                                             \
    } while (0)

int ensure_dev(rsa_ext* h, DevBuf& b, size_t need) {
    if (need <= b.cap) return RSA_EXT_OK;
    size_t cap = std::max(need, b.cap + b.cap / 2);
    cap = align_up(cap, 1 << 20);
    if (b.p) CU_TRY(h, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CU_TRY(h, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return RSA_EXT_OK;
}

// planned tiles + head-room for pairs the redo pass re-tiles (symbols outside ACGTN)
inline size_t scratch_alloc_bytes(uint64_t planned) { return (size_t)align_up((size_t)planned, 256) + std::max<size_t>((size_t)16 << 20, (size_t)planned / 16); }

// Size a slot's device buffers for one chunk.  Everything except the direction scratch is carved from ONE
// allocation: cudaMalloc/cudaFree take a process-wide driver lock and stall every other worker's enqueue
// (measured ~8 ms per round of eight cudaMallocs with 16 workers), so a slot allocates at most twice per growth.

int ensure_slot(rsa_ext* h, Slot& s, const SlotNeed& n) {
    DevBuf* view[7] = {&s.d_blob, &s.d_q, &s.d_t, &s.d_ends, &s.d_res, &s.d_arena, &s.d_aln};
    const size_t need[7] = {n.blob, n.q, n.t, n.ends, n.res, n.arena, n.aln};
    size_t off[7], total = 0;
    for (int i = 0; i < 7; ++i) { off[i] = total; total += align_up(need[i], 256); }
    if (total > s.d_slab.cap) {
        size_t cap = align_up(std::max({total, 2 * s.d_slab.cap, (size_t)4 << 20}), 1 << 20);
        if (s.d_slab.p) CU_TRY(h, cudaFree(s.d_slab.p));
        s.d_slab = DevBuf{};
        CU_TRY(h, cudaMalloc(&s.d_slab.p, cap));
        s.d_slab.cap = cap;
    }
    for (int i = 0; i < 7; ++i) { view[i]->p = s.d_slab.p + off[i]; view[i]->cap = need[i]; }
    if (n.scratch > s.d_scratch.cap) {
        size_t cap = align_up(std::max({n.scratch, 2 * s.d_scratch.cap, (size_t)16 << 20}), 1 << 20);
        cap = std::min(cap, std::max(n.scratch, scratch_alloc_bytes(h->scratch_per_slot)));
        if (s.d_scratch.p) CU_TRY(h, cudaFree(s.d_scratch.p));
        s.d_scratch = DevBuf{};
        CU_TRY(h, cudaMalloc(&s.d_scratch.p, cap));
        s.d_scratch.cap = cap;
    }
    return RSA_EXT_OK;
}

// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Pageable memory reports cudaMemoryTypeUnregistered.
bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int ensure_pin(rsa_ext* h, PinBuf& b, size_t need) {
    if (need <= b.cap) return RSA_EXT_OK;
    size_t cap = std::max(need, b.cap + b.cap / 2);
    cap = align_up(cap, 1 << 16);
    if (b.p) CU_TRY(h, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    CU_TRY(h, cudaHostAlloc(&b.p, cap, cudaHostAllocDefault));
    b.cap = cap;
    return RSA_EXT_OK;
}

// ---- planning -------------------------------------------------------------------------------------
//
// Decide, for the pairs [lo, ...) of the pending batch, how many fit one chunk and lay out the metadata
// blob in `blob` (host memory, grown as needed through `grow`).  Pure host code: unit-tested without a GPU
// through rsa_ext_plan_debug().
struct PlanInput {
    int64_t n;
    const int64_t* qoff;
    const int64_t* toff;
    const char* qbuf;
    const char* tbuf;
    int max_qlen, max_tlen;
    size_t scratch_cap;
    bool exact_only;
    int64_t max_pairs = kMaxChunkPairsDefault;  // cap for this chunk (the first chunks of a batch ramp up)
    int match = 2;
    // window form (rsa_ext_submit_ref_windows): target i = [win_off[i], win_off[i] + win_len[i]) of the resident
    // reference; `toff`/`tbuf` are unused then and no window bytes travel
    const int64_t* win_off = nullptr;
    const int32_t* win_len = nullptr;
    int64_t ref_len = 0;
    int64_t t_start(int64_t i) const { return win_off ? win_off[i] : toff[i]; }
    int64_t t_len(int64_t i) const { return win_off ? (int64_t)win_len[i] : toff[i + 1] - toff[i]; }
};

// A pair may ride the packed kernel when its shape is inside what that kernel was instantiated for.
inline bool fast_shape_ok(int qlen, int tlen, int match = 2) {
    return qlen >= kFastMinQlen && qlen <= kFastMaxQlen && tlen >= 1 && tlen <= kFastMaxTlen &&
           fast_key_ok(qlen, match);  // the maximum-tracking key holds (score << 5 or 6 | column) in a positive s16
}

// per-query-length geometry, computed once
struct LenTables {
    uint32_t fast_row_bytes[kFastMaxQlen + 1];  // direction bytes per target row of a packed group
    uint16_t fast_C[kFastMaxQlen + 1];
    uint8_t fast_L[kFastMaxQlen + 1];
    uint32_t exact_row_bytes_[513];
    LenTables() {
        for (int q = 0; q <= kFastMaxQlen; ++q) {
            const FastGeom g = fast_geom(q < 1 ? 1 : q);
            fast_row_bytes[q] = (uint32_t)(g.L * g.W * 4);
            fast_L[q] = (uint8_t)g.L;
            fast_C[q] = (uint16_t)g.C;
        }
        for (int q = 0; q <= 512; ++q) exact_row_bytes_[q] = (uint32_t)exact_row_bytes(q);
    }
};
const LenTables& len_tables() { static const LenTables t; return t; }

// Section offsets of the metadata blob of an n-pair chunk; returns its size.
size_t blob_layout(ChunkPlan& plan, int64_t n) {
    size_t off = 0;
    plan.off_meta = off;   off = align_up(off + sizeof(PairMeta) * n, 16);
    plan.off_info = off;   off = align_up(off + sizeof(uint32_t) * n, 16);
    plan.off_diroff = off; off = align_up(off + sizeof(uint64_t) * n, 16);
    plan.off_list = off;   off = align_up(off + sizeof(uint32_t) * n, 16);
    plan.off_groups = off; off = align_up(off + sizeof(FastGroup) * (size_t)(n + 8 * 128), 16);
    plan.off_redo = off;   off = align_up(off + sizeof(RedoHeader) + sizeof(uint32_t) * (size_t)(n + 4), 16);
    plan.blob_bytes = off;
    return off;
}

// Variable-length batches: one launch per column class means many small launches, each a partly filled, latency-bound
// wave (a 131 k-pair chunk of indel-rich 250-bp reads has ten classes).  Neighbouring classes of one lane count whose widths
// fall into the same four-column bucket can share a "flex" launch of the bucket's width (FastDp<.., FLEX>), at the price of
// the idle column slots of the narrower groups.  Merge a bucket when those idle columns cost less than the partial waves
// the separate launches would leave (about half a wave each).  Classes are contiguous runs of the group array, so a merged
// class is just the union of the runs.
void merge_small_classes(std::vector<ChunkPlan::FastClass>& fast) {
    if (fast.size() < 3 || getenv("RSA_EXT_NO_FLEX")) return;
    std::vector<ChunkPlan::FastClass> out;
    size_t i = 0;
    while (i < fast.size()) {
        const int L = fast[i].L, b = fast_flex_width(L, fast[i].C);
        size_t j = i;
        double idle = 0, groups = 0;
        while (j < fast.size() && fast[j].L == L && fast_flex_width(L, fast[j].C) == b && b - fast[j].C <= kFlexSlack) {
            idle += (double)fast[j].n_groups * (b - fast[j].C) / b;
            groups += fast[j].n_groups;
            ++j;
        }
        const int members = (int)(j - i);
        const int gpb = fast_warps_per_block(L) * fast_groups_per_warp(L);
        const int blocks_per_sm = L == 16 ? (b <= 27 ? 6 : 4) : (b <= 20 ? 4 : (b <= 27 ? 3 : (L == 4 ? RSA_FAST_WIDE_BLOCKS : RSA_FAST_WIDE_BLOCKS8)));
        const double wave = 148.0 * blocks_per_sm * gpb;   // groups in flight in one full wave
        if (members >= 2 && idle < 0.5 * wave * (members - 1)) {
            ChunkPlan::FastClass m = fast[i];
            m.C = b;
            m.flex = true;
            m.n_groups = fast[j - 1].group_begin + fast[j - 1].n_groups - fast[i].group_begin;
            for (size_t k = i; k < j; ++k) m.max_tlen = std::max(m.max_tlen, fast[k].max_tlen);
            out.push_back(m);
        } else {
            for (size_t k = i; k < j; ++k) out.push_back(fast[k]);
        }
        i = j > i ? j : i + 1;
    }
    fast.swap(out);
}

int plan_chunk(rsa_ext* h, const PlanInput& in, int64_t lo, ChunkPlan& plan, std::vector<uint8_t>* vec_blob,
               PinBuf* pin_blob) {
    const LenTables& LT = len_tables();
    // 1) how many pairs: stop at the pair/sequence-byte caps or when the scratch budget is reached.
    //    Scratch is budgeted with the larger of the two layouts' needs so any routing fits.
    int64_t hi = lo;
    uint64_t scratch = 0;
    const int64_t q0 = in.qoff[lo], t0 = in.win_off ? 0 : in.toff[lo];
    const int64_t hi_cap = std::min<int64_t>(in.n, lo + std::min<int64_t>(kMaxChunkPairs, in.max_pairs));
    while (hi < hi_cap) {
        const int64_t ql = in.qoff[hi + 1] - in.qoff[hi], tl = in.t_len(hi);
        if (ql < 0 || tl < 0) { h->err = "offsets are not monotone"; return RSA_EXT_ERR_ARG; }
        if (ql > in.max_qlen) {
            h->err = "read size is too big, " + std::to_string(ql) + " > " + std::to_string(in.max_qlen);
            return RSA_EXT_ERR_QUERY_LEN;
        }
        uint64_t need = 0;
        if (ql > 0 && tl > 0 && tl <= in.max_tlen) {
            need = (uint64_t)tl * LT.exact_row_bytes_[ql] + 16;
            if (fast_shape_ok((int)ql, (int)tl, in.match))  // a lone pair owns a whole group
                need = std::max<uint64_t>(need, (uint64_t)fast_tile_rows((int)tl) * LT.fast_row_bytes[ql]);
        }
        if (hi > lo && (scratch + need > in.scratch_cap || in.qoff[hi + 1] - q0 > kMaxChunkSeqBytes ||
                        (!in.win_off && in.toff[hi + 1] - t0 > kMaxChunkSeqBytes)))
            break;
        scratch += need;
        ++hi;
    }
    const int64_t n = hi - lo;
    plan = ChunkPlan();
    plan.lo = lo; plan.hi = hi; plan.n = n;
    plan.q_bytes = in.qoff[hi] - q0;
    plan.t_bytes = in.win_off ? 0 : in.toff[hi] - t0;

    // 2) blob layout
    const size_t off = blob_layout(plan, n);
    uint8_t* blob;
    if (vec_blob) { vec_blob->resize(off); blob = vec_blob->data(); }
    else {
        int rc = ensure_pin(h, *pin_blob, off);
        if (rc) return rc;
        blob = pin_blob->p;
    }
    PairMeta* meta = reinterpret_cast<PairMeta*>(blob + plan.off_meta);
    uint32_t* info = reinterpret_cast<uint32_t*>(blob + plan.off_info);
    uint64_t* diroff = reinterpret_cast<uint64_t*>(blob + plan.off_diroff);
    uint32_t* list = reinterpret_cast<uint32_t*>(blob + plan.off_list);
    FastGroup* groups = reinterpret_cast<FastGroup*>(blob + plan.off_groups);
    memset(blob + plan.off_redo, 0, sizeof(RedoHeader));

    // 3) per-pair records; packed-kernel candidates collected as (key = qlen<<16 | tlen, index)
    for (int k = 0; k < 3; ++k) h->tmp_list[k].clear();
    std::vector<uint32_t>& cand_idx = h->tmp_sort;
    std::vector<uint32_t>& cand_key = h->tmp_key;
    cand_idx.resize((size_t)n);
    cand_key.resize((size_t)n);
    size_t m = 0;
    uint64_t arena = 0;
    int64_t cells = 0;
    uint32_t qmin = 0xFFFFFFFFu, qmax = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t qo = in.qoff[lo + i], to = in.t_start(lo + i);
        const int64_t ql = in.qoff[lo + i + 1] - qo, tl = in.t_len(lo + i);
        meta[i].qoff = (uint32_t)(qo - q0);
        meta[i].toff = (uint32_t)(to - t0);
        meta[i].qlen = (uint16_t)ql;
        meta[i].tlen = (uint16_t)std::min<int64_t>(tl, 65535);
        info[i] = 0;
        diroff[i] = 0;
        cells += ql * tl;
        if (ql == 0 || tl == 0) { info[i] = 3u << 16; plan.n_failed++; continue; }
        if (tl > in.max_tlen) { info[i] = 1u << 16; plan.n_failed++; continue; }
        arena += (uint64_t)(ql + tl + 1);
        if (!in.exact_only && fast_shape_ok((int)ql, (int)tl, in.match)) {
            cand_idx[m] = (uint32_t)i;
            cand_key[m] = ((uint32_t)ql << 16) | (uint32_t)tl;
            qmin = std::min(qmin, (uint32_t)ql);
            qmax = std::max(qmax, (uint32_t)ql);
            ++m;
        } else {
            const int cls = LT.exact_row_bytes_[ql] == 64 ? 0 : (LT.exact_row_bytes_[ql] == 128 ? 1 : 2);
            h->tmp_list[cls].push_back((uint32_t)i);
        }
    }
    plan.cells = cells;
    plan.arena_bytes = arena + 64;

    // 4) packed-kernel groups: stable counting sort of the candidates by target length, then (only if the
    //    chunk mixes query lengths) by query length; consecutive equal-qlen pairs share a group (A = low
    //    halves, B = high halves), an odd one out runs with B = A.  Groups of one column class are padded
    //    to whole warps (4 groups) with empty slots.
    uint64_t sc_off = 0;
    int n_groups = 0;
    plan.fast.clear();
    if (m > 0) {
        std::vector<uint32_t>& idx2 = h->tmp_order;
        std::vector<uint32_t>& key2 = h->tmp_key2;
        std::vector<uint32_t>& cnt = h->tmp_count;
        idx2.resize(m);
        key2.resize(m);
        cnt.assign(kFastMaxTlen + 2, 0);
        for (size_t k = 0; k < m; ++k) cnt[(cand_key[k] & 0xFFFFu) + 1]++;
        for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
        for (size_t k = 0; k < m; ++k) {
            const uint32_t p = cnt[cand_key[k] & 0xFFFFu]++;
            idx2[p] = cand_idx[k];
            key2[p] = cand_key[k];
        }
        const uint32_t* sidx = idx2.data();
        const uint32_t* skey = key2.data();
        if (qmin != qmax) {
            cnt.assign(kFastMaxQlen + 2, 0);
            for (size_t k = 0; k < m; ++k) cnt[(key2[k] >> 16) + 1]++;
            for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
            for (size_t k = 0; k < m; ++k) {
                const uint32_t p = cnt[key2[k] >> 16]++;
                cand_idx[p] = idx2[k];
                cand_key[p] = key2[k];
            }
            sidx = cand_idx.data();
            skey = cand_key.data();
        }
        const FastGroup empty{0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
        size_t k = 0;
        int cur_C = -1, cur_L = -1;
        constexpr int kGroupPad = 8;  // classes start on a warp boundary for every group width (8, 4 or 2 groups per warp)
        while (k < m) {
            const uint32_t a = sidx[k], ka = skey[k];
            const uint32_t ql = ka >> 16;
            uint32_t b = a, kb = ka;
            if (k + 1 < m && (skey[k + 1] >> 16) == ql) { b = sidx[k + 1]; kb = skey[k + 1]; k += 2; }
            else k += 1;
            const int C = LT.fast_C[ql], L = LT.fast_L[ql];
            if (C != cur_C || L != cur_L) {
                while (n_groups % kGroupPad) groups[n_groups++] = empty;
                if (!plan.fast.empty()) plan.fast.back().n_groups = n_groups - plan.fast.back().group_begin;
                plan.fast.push_back({L, C, n_groups, 0, 0});
                cur_C = C;
                cur_L = L;
            }
            const uint32_t rows = std::max(ka & 0xFFFFu, kb & 0xFFFFu);
            FastGroup fg;
            fg.a = a; fg.b = b;
            fg.dir_off = sc_off;
            fg.qlen = (uint16_t)ql;
            fg.rows = (uint16_t)rows;
            groups[n_groups++] = fg;
            diroff[a] = sc_off;
            if (b != a) { diroff[b] = sc_off; info[b] |= 1u; }
            sc_off += (uint64_t)fast_tile_rows((int)rows) * LT.fast_row_bytes[ql];
            plan.fast.back().max_tlen = std::max<int>(plan.fast.back().max_tlen, (int)rows);
            plan.n_fast_pairs += (b != a) ? 2 : 1;
        }
        while (n_groups % kGroupPad) groups[n_groups++] = empty;
        plan.fast.back().n_groups = n_groups - plan.fast.back().group_begin;
    }
    merge_small_classes(plan.fast);
    plan.n_fast_classes = (int)plan.fast.size();

    // 5) exact-kernel lists (statically routed pairs); pairs the packed kernel flags at run time get their
    //    tiles from the head-room behind `scratch_bytes` (exact_redo_kernel).
    uint32_t pos = 0;
    for (int c = 0; c < 3; ++c) {
        plan.n_exact[c] = (int)h->tmp_list[c].size();
        for (uint32_t i : h->tmp_list[c]) {
            list[pos++] = i;
            diroff[i] = sc_off;
            sc_off += align_up((size_t)meta[i].tlen * exact_row_bytes(meta[i].qlen), 16);
            plan.max_tlen_exact[c] = std::max<int>(plan.max_tlen_exact[c], meta[i].tlen);
        }
    }
    plan.scratch_bytes = sc_off;
    return RSA_EXT_OK;
}


// ---- device planner, host side -----------------------------------------------------------------------
//
// One streaming pass over the chunk's offsets: where the chunk ends (pair / byte caps, direction-scratch budget), how
// many pairs each kernel family gets and -- per query length -- how many packed candidates there are, which fixes the
// launch geometry (column classes, group slots) without sorting anything.  The sort, the pairing and every tile offset
// are computed on the device (kernels_plan.cuh).
//
// Scratch budget without knowing the pairing: per query length, with the candidates' 4-row-rounded window lengths
// sorted t_1 <= ... <= t_m and paired (t_1,t_2), (t_3,t_4), ..., the groups' row counts sum to
//   sum_k max(t_2k-1, t_2k) <= (sum_i t_i + t_m - t_1) / 2   (+ t_m for a lone last pair),
// so bytes(q) <= row_bytes(q) * ((sum + max - min + 1) / 2 + max).  Maintained incrementally while the chunk grows.
constexpr int64_t kDevPlanMinPairs = 16384;  // smaller batches are planned on the host (the planner's nine launches cost
                                             // more than 13 ns/pair there); RSA_EXT_DEV_PLAN_MIN overrides (tests)
inline int64_t dev_plan_min_pairs() {
    static const int64_t v = [] {
        const char* e = getenv("RSA_EXT_DEV_PLAN_MIN");
        return e ? (int64_t)atoll(e) : kDevPlanMinPairs;
    }();
    return v;
}

size_t blob_layout_dev(ChunkPlan& plan, int64_t n, int n_group_slots, bool windows) {
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    plan.off_in_q = take(sizeof(int64_t) * (size_t)(n + 1));
    plan.off_in_t = take(sizeof(int64_t) * (size_t)(n + 1));
    plan.off_in_wlen = take(windows ? sizeof(int32_t) * (size_t)n : 0);
    plan.off_qtab = take(sizeof(PlanQlen) * kPlanQ);
    plan.off_meta = take(sizeof(PairMeta) * (size_t)n);
    plan.off_info = take(sizeof(uint32_t) * (size_t)n);
    plan.off_diroff = take(sizeof(uint64_t) * (size_t)n);
    plan.off_list = take(sizeof(uint32_t) * (size_t)n);
    plan.off_groups = take(sizeof(FastGroup) * (size_t)n_group_slots);
    plan.off_redo = take(sizeof(RedoHeader) + sizeof(uint32_t) * (size_t)(n + 4));
    plan.off_key = take(sizeof(uint32_t) * (size_t)n);
    plan.off_rank = take(sizeof(uint32_t) * (size_t)n);
    plan.off_sorted = take(sizeof(uint32_t) * (size_t)n);
    // zeroed before planning, one memset: header, group tile sizes, scan block sums, histogram
    plan.off_hdr = take(sizeof(PlanHeader));
    plan.off_gbytes = take(sizeof(uint32_t) * (size_t)n_group_slots);
    plan.off_bsum = take(sizeof(uint32_t) * kPlanScanBlocks);
    plan.off_hist = take(sizeof(uint32_t) * kPlanBins);
    plan.zero_bytes = off - plan.off_hdr;
    plan.n_group_slots = n_group_slots;
    plan.blob_bytes = off;
    return off;
}

// `qtab` (kPlanQ entries) is filled for the upload.
int scan_chunk(rsa_ext* h, const PlanInput& in, int64_t lo, ChunkPlan& plan, PlanQlen* qtab) {
    const LenTables& LT = len_tables();
    plan = ChunkPlan();
    plan.dev = true;
    // per-|q| accumulators of the packed candidates; the bound is kept in HALF-row units:
    //   N(q) = sum4 + 3 * max4 - min4 + 1  >=  2 * (rows of all groups of this |q|)      (see above)
    uint32_t cnt[kPlanQ], mn4[kPlanQ], mx4[kPlanQ];
    memset(cnt, 0, sizeof cnt);
    memset(mx4, 0, sizeof mx4);
    const int64_t q0 = in.qoff[lo], t0 = in.win_off ? 0 : in.toff[lo];
    const int64_t hi_cap = std::min<int64_t>(in.n, lo + std::min<int64_t>(kMaxChunkPairs, in.max_pairs));
    const uint64_t cap2 = in.scratch_cap > 2048 ? 2 * ((uint64_t)in.scratch_cap - 2048) : 0;  // budget in half-row bytes
    const int64_t fast_tmax = std::min<int64_t>(in.max_tlen, kFastMaxTlen);
    const int64_t fast_qmax = in.exact_only ? -1 : kFastMaxQlen;
    bool key_ok[kPlanQ];  // the maximum-tracking key of this |q| fits its s16 half with this handle's match score
    for (int q = 0; q < kPlanQ; ++q) key_ok[q] = fast_key_ok(q, in.match);
    uint64_t scratch2 = 0, arena = 0;
    int64_t cells = 0, hi = lo, n_fast = 0, n_failed = 0;
    int64_t qprev = in.qoff[lo];
    int64_t tprev = in.win_off ? 0 : in.toff[lo];
    int64_t cur_q = -1;
    uint32_t c_cnt = 0, c_mn = 0, c_mx = 0, c_rowb = 0;
    for (; hi < hi_cap; ++hi) {
        const int64_t qnext = in.qoff[hi + 1];
        const int64_t ql = qnext - qprev;
        int64_t tl, tnext = 0;
        if (in.win_off) {
            tl = in.win_len[hi];
            const int64_t wo = in.win_off[hi];
            if (wo < 0 || tl < 0 || wo + tl > in.ref_len) {
                h->err = "window " + std::to_string(hi) + " lies outside the resident reference";
                return RSA_EXT_ERR_ARG;
            }
        } else {
            tnext = in.toff[hi + 1];
            tl = tnext - tprev;
        }
        if (ql < 0 || tl < 0) { h->err = "offsets are not monotone"; return RSA_EXT_ERR_ARG; }
        if (ql > in.max_qlen) {
            h->err = "gasal2 : read size is too big, " + std::to_string(ql) + " > " + std::to_string(in.max_qlen);
            return RSA_EXT_ERR_QUERY_LEN;
        }
        if (qnext - q0 > kMaxChunkSeqBytes || tnext - t0 > kMaxChunkSeqBytes) { if (hi > lo) break; }
        if (ql >= kFastMinQlen && ql <= fast_qmax && tl >= 1 && tl <= fast_tmax && key_ok[ql]) {
            // packed candidate: what it adds to the bound.  The accumulators of the current |q| live in registers
            // (batches are runs of equal read lengths); they go back to the tables when |q| changes.
            if (ql != cur_q) {
                if (cur_q >= 0) { cnt[cur_q] = c_cnt; mn4[cur_q] = c_mn; mx4[cur_q] = c_mx; }
                cur_q = ql;
                c_cnt = cnt[ql]; c_mn = mn4[ql]; c_mx = mx4[ql];
                c_rowb = LT.fast_row_bytes[ql];
            }
            const uint32_t t4 = (uint32_t)((tl + 3) & ~3);
            uint64_t add2;
            uint32_t nmn = t4, nmx = t4;
            if (c_cnt == 0) add2 = (uint64_t)c_rowb * (3ull * t4 + 1);
            else {
                nmn = std::min(c_mn, t4);
                nmx = std::max(c_mx, t4);
                add2 = (uint64_t)c_rowb * (t4 + 3ull * (nmx - c_mx) + (c_mn - nmn));
            }
            // tiles are whole 8-row units: up to 4 rows more than ceil4 per GROUP, i.e. per two pairs (a lone pair owns a
            // group): 4 per pair in these doubled units, 8 for the first pair of a length
            add2 += (uint64_t)c_rowb * (c_cnt == 0 ? 8u : 4u);
            if (scratch2 + add2 > cap2 && hi > lo) break;
            scratch2 += add2;
            ++c_cnt; c_mn = nmn; c_mx = nmx;
            ++n_fast;
            arena += (uint64_t)(ql + tl + 1);
        } else if (ql > 0 && tl > 0 && tl <= in.max_tlen) {
            const uint32_t rb = LT.exact_row_bytes_[ql];
            const uint64_t add2 = 2 * (uint64_t)align_up((size_t)tl * rb, 16);
            if (scratch2 + add2 > cap2 && hi > lo) break;
            scratch2 += add2;
            const int cls = rb == 64 ? 0 : (rb == 128 ? 1 : 2);
            plan.n_exact[cls]++;
            plan.max_tlen_exact[cls] = std::max<int>(plan.max_tlen_exact[cls], (int)tl);
            arena += (uint64_t)(ql + tl + 1);
        } else {
            ++n_failed;
        }
        cells += ql * tl;
        qprev = qnext;
        tprev = tnext;
    }
    if (cur_q >= 0) { cnt[cur_q] = c_cnt; mn4[cur_q] = c_mn; mx4[cur_q] = c_mx; }
    const int64_t n = hi - lo;
    plan.lo = lo; plan.hi = hi; plan.n = n;
    plan.n_fast_pairs = n_fast;
    plan.n_failed = n_failed;
    plan.q_bytes = in.qoff[hi] - q0;
    plan.t_bytes = in.win_off ? 0 : in.toff[hi] - t0;
    plan.cells = cells;
    plan.arena_bytes = arena + 64;
    plan.scratch_bytes = (scratch2 + 1) / 2 + 2048;  // upper bound (+ the 256-byte alignments of the packed region / redo base)
    // launch geometry from the per-length counts: classes in increasing |q| (= increasing C within 8 lanes, then 16 lanes),
    // each padded to whole warps of groups
    constexpr int kGroupPad = 8;
    uint32_t pos = 0;
    int n_groups = 0, cur_C = -1, cur_L = -1;
    for (int q = 0; q < kPlanQ; ++q) {
        qtab[q].count = cnt[q];
        qtab[q].pos_base = pos;
        qtab[q].group_base = 0;
        if (cnt[q] == 0) continue;
        const int C = LT.fast_C[q], L = LT.fast_L[q];
        if (C != cur_C || L != cur_L) {
            while (n_groups % kGroupPad) ++n_groups;
            if (!plan.fast.empty()) plan.fast.back().n_groups = n_groups - plan.fast.back().group_begin;
            plan.fast.push_back({L, C, n_groups, 0, 0});
            cur_C = C; cur_L = L;
        }
        qtab[q].group_base = (uint32_t)n_groups;
        n_groups += (int)((cnt[q] + 1) / 2);
        pos += cnt[q];
        plan.fast.back().max_tlen = std::max(plan.fast.back().max_tlen, (int)mx4[q]);  // (rounded up to 4 rows)
    }
    while (n_groups % kGroupPad) ++n_groups;
    if (!plan.fast.empty()) plan.fast.back().n_groups = n_groups - plan.fast.back().group_begin;
    merge_small_classes(plan.fast);
    plan.n_fast_classes = (int)plan.fast.size();
    plan.list_base[0] = 0;
    plan.list_base[1] = (unsigned)plan.n_exact[0];
    plan.list_base[2] = (unsigned)(plan.n_exact[0] + plan.n_exact[1]);
    blob_layout_dev(plan, n, std::max(n_groups, kGroupPad), in.win_off != nullptr);
    return RSA_EXT_OK;
}

// ---- launching -------------------------------------------------------------------------------------

struct ChunkDev {
    const uint8_t* blob;
    const uint8_t* q;
    const uint8_t* t;
    DpEnd* ends;
    rsa_ext_result_t* res;
    uint8_t* scratch;
    uint64_t scratch_cap;
    uint8_t* arena;
    unsigned long long* arena_used;
    uint64_t arena_cap;
    const uint4* tpack = nullptr;  // window form: the resident reference's packed planes (else null: ASCII targets)
    const uint2* tflag = nullptr;
};

template <int C>
void launch_exact(cudaStream_t st, const ChunkDev& d, const ChunkPlan& p, int list_begin, int n_list, int max_tlen,
                  const Scoring& sc) {
    const int tlen_pad = (int)align_up((size_t)max_tlen, 16);
    const int blocks = (n_list + kExactWarpsPerBlock - 1) / kExactWarpsPerBlock;
    exact_dp_kernel<C><<<blocks, 32 * kExactWarpsPerBlock, (size_t)tlen_pad * kExactWarpsPerBlock, st>>>(
        d.q, d.t, reinterpret_cast<const PairMeta*>(d.blob + p.off_meta),
        reinterpret_cast<const uint32_t*>(d.blob + p.off_list) + list_begin, n_list,
        reinterpret_cast<const uint64_t*>(d.blob + p.off_diroff), d.scratch, d.ends, sc, tlen_pad);
}

// Enqueue every kernel of one chunk: the DP kernels on `st`, the traceback on `st_tb` behind `ev_mid`, so the
// traceback of chunk k (a latency-bound pointer chase) overlaps the DP kernels of chunk k+1 (issue-bound).
// ev[0..3], when given, bracket the DP phase (on st) and the traceback (on st_tb).
int enqueue_compute(rsa_ext* h, cudaStream_t st, cudaStream_t st_tb, cudaEvent_t ev_mid, const ChunkDev& d,
                    const ChunkPlan& p, cudaEvent_t* ev) {
    CU_TRY(h, cudaMemsetAsync(d.ends, 0, sizeof(DpEnd) * p.n, st));
    CU_TRY(h, cudaMemsetAsync(const_cast<uint8_t*>(d.blob) + p.off_redo, 0, sizeof(RedoHeader), st));
    CU_TRY(h, cudaMemsetAsync(d.arena_used, 0, kSlotCounters * sizeof(unsigned long long), st));
    if (ev) CU_TRY(h, cudaEventRecord(ev[0], st));
    const PairMeta* meta = reinterpret_cast<const PairMeta*>(d.blob + p.off_meta);
    const uint32_t* info = reinterpret_cast<const uint32_t*>(d.blob + p.off_info);
    uint64_t* diroff = reinterpret_cast<uint64_t*>(const_cast<uint8_t*>(d.blob) + p.off_diroff);
    RedoHeader* redo = reinterpret_cast<RedoHeader*>(const_cast<uint8_t*>(d.blob) + p.off_redo);
    uint32_t* redo_list = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(d.blob) + p.off_redo + sizeof(RedoHeader));
    TbArgs tba{d.q, d.t, meta, info, diroff, d.scratch, d.res, h->sc, d.arena, d.arena_used, d.arena_cap};
    // packed kernel, one launch per column class.  A chunk of variable-length reads (indel-rich 250-bp data: five or more
    // classes) would run them back to back, each with its own partially filled last wave; classes are independent, so
    // they are spread over the chunk's stream and two side streams and joined before the redo pass.
    const bool spread = p.fast.size() > 1 && (h->cfg.flags & RSA_EXT_FLAG_SERIALIZE) == 0;
    if (spread) {
        CU_TRY(h, cudaEventRecord(h->ev_cls_fork, st));
        for (int k = 0; k < kClsStreams; ++k) CU_TRY(h, cudaStreamWaitEvent(h->s_cls[k], h->ev_cls_fork, 0));
    }
    // largest classes first: their waves fill the SMs, the small (single-wave, latency-bound) classes run beside them
    std::vector<int> order(p.fast.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    if (spread) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return p.fast[a].n_groups * p.fast[a].L > p.fast[b].n_groups * p.fast[b].L; });
    int cls_i = 0;
    bool used_side[kClsStreams] = {};
    for (int oi : order) {
        const auto& fc = p.fast[oi];
        cudaStream_t cst = st;
        const int slot = cls_i % (kClsStreams + 1);
        if (spread && slot != 0) { cst = h->s_cls[slot - 1]; used_side[slot - 1] = true; }
        ++cls_i;
        int rc = launch_fast_class(cst, fc.L, fc.C, fc.flex, d.q, d.t, meta,
                                   reinterpret_cast<const FastGroup*>(d.blob + p.off_groups) + fc.group_begin,
                                   fc.n_groups, d.scratch, d.ends, redo, redo_list, h->fk, fc.max_tlen, d.tpack, d.tflag);
        if (rc != 0) { h->err = "no packed-kernel instance for L=" + std::to_string(fc.L) + " C=" + std::to_string(fc.C); return RSA_EXT_ERR_STATE; }
        h->stats.kernel_launches++;
    }
    for (int k = 0; k < kClsStreams; ++k)
        if (used_side[k]) {
            CU_TRY(h, cudaEventRecord(h->ev_cls_join[k], h->s_cls[k]));
            CU_TRY(h, cudaStreamWaitEvent(st, h->ev_cls_join[k], 0));
        }
    // exact kernel: statically routed pairs
    int begin = 0;
    for (int c = 0; c < 3; ++c) {
        const int nl = p.n_exact[c];
        if (nl > 0) {
            if (c == 0) launch_exact<4>(st, d, p, begin, nl, p.max_tlen_exact[c], h->sc);
            else if (c == 1) launch_exact<8>(st, d, p, begin, nl, p.max_tlen_exact[c], h->sc);
            else launch_exact<16>(st, d, p, begin, nl, p.max_tlen_exact[c], h->sc);
            h->stats.kernel_launches++;
        }
        begin += nl;
    }
    // pairs the packed kernel declined: one warp-per-pair exact pass driven by the DPF_NEED_EXACT flags
    if (!p.fast.empty()) {
        int max_tlen = 0;
        for (const auto& fc : p.fast) max_tlen = std::max(max_tlen, fc.max_tlen);
        const unsigned long long redo_base = align_up((size_t)p.scratch_bytes, 256);
        const unsigned long long* redo_base_dev =
            p.dev ? &reinterpret_cast<const PlanHeader*>(d.blob + p.off_hdr)->redo_base : nullptr;
        launch_exact_redo(st, d.q, d.t, meta, diroff, d.scratch, d.ends, redo, redo_list, h->sc, max_tlen, h->n_sms,
                          redo_base, d.scratch_cap, redo_base_dev, d.arena_used);
        h->stats.kernel_launches++;
    }
    if (ev) CU_TRY(h, cudaEventRecord(ev[1], st));
    CU_TRY(h, cudaEventRecord(ev_mid, st));
    CU_TRY(h, cudaStreamWaitEvent(st_tb, ev_mid, 0));
    if (ev) CU_TRY(h, cudaEventRecord(ev[2], st_tb));
    if (!p.fast.empty()) {
        // the packed kernel's pairs, in group order (pairs a and b of a group walk the same words)
        int total_groups = 0;
        for (const auto& fc : p.fast) total_groups = std::max(total_groups, fc.group_begin + fc.n_groups);
        tb_groups_kernel<<<(unsigned)((2 * total_groups + kTbThreads - 1) / kTbThreads), kTbThreads, 0, st_tb>>>(
            tba, reinterpret_cast<const FastGroupRef*>(d.blob + p.off_groups), total_groups, d.ends);
        h->stats.kernel_launches++;
    }
    // everything else: exact-kernel pairs, redone pairs, failed records
    tb_kernel<<<(unsigned)((p.n + kTbThreads - 1) / kTbThreads), kTbThreads, 0, st_tb>>>(tba, (int)p.n, d.ends);
    h->stats.kernel_launches++;
    if (ev) CU_TRY(h, cudaEventRecord(ev[3], st_tb));
    CU_TRY(h, cudaGetLastError());
    return RSA_EXT_OK;
}

PlanInput pending_plan_input(const rsa_ext* h) {
    PlanInput in{h->n, h->qoff, h->toff, h->qbuf, h->tbuf, h->cfg.max_query_len, h->cfg.max_target_len,
                 h->scratch_per_slot, (h->cfg.flags & RSA_EXT_FLAG_EXACT_ONLY) != 0 || !h->fast_ok};
    in.match = h->sc.match;
    in.win_off = h->win_off;
    in.win_len = h->win_len;
    in.ref_len = h->ref ? h->ref->len : 0;
    return in;
}

// Chunk-size ramp of a batch: the first chunks are small so that the GPU starts while the host still plans.
// Inline planning shares the caller's thread with enqueue/retire (16k, 32k, 64k); the plan-ahead thread runs
// continuously, so its ramp only has to keep (plan + H2D) of chunk k+1 below the GPU time of chunk k.
// Balanced tail: when the greedy cut of a chunk leaves less than half a chunk behind it, the batch ends with a small
// chunk whose launches are single, partly filled waves (latency-bound: ~0.3 ms per column class for almost no work).
// Cutting the last two chunks evenly instead costs one more planning pass over this chunk's pairs.  Returns the pair cap
// for a second pass, or 0 to keep the cut.
inline int64_t balanced_tail_cap(int64_t n, int64_t lo, int64_t hi) {
    const int64_t tail = n - hi, got = hi - lo;
    if (tail <= 0 || got < 8192 || 2 * tail >= got) return 0;
    return (n - lo + 1) / 2;
}

int64_t ramp_pairs(bool plan_ahead, long chunk_index) {
    static const std::vector<int64_t> knob = [] {  // RSA_EXT_RAMP="16384,65536": experiment knob
        std::vector<int64_t> v;
        if (const char* e = getenv("RSA_EXT_RAMP")) {
            for (const char* p = e; *p;) { v.push_back(atoll(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
        }
        return v;
    }();
    if (!knob.empty()) return chunk_index < (long)knob.size() ? std::max<int64_t>(256, knob[chunk_index]) : kMaxChunkPairs;
    if (plan_ahead) {
        static const int64_t r[] = {16384, 24576, 40960, 65536};
        return chunk_index < 4 ? r[chunk_index] : kMaxChunkPairs;
    }
    // inline planning (device planner or small batches): measured on the bench batch (1 M pairs, e2e / resident):
    // 16k,32k,64k 0.946; 32k,64k 0.957; 64k 0.945; none 0.916 (profiles/r2_e2e_ramp.md)
    return chunk_index < 2 ? ((int64_t)32768 << chunk_index) : kMaxChunkPairs;
}

void plan_ahead_main(rsa_ext* h) {
    PlanAhead& pa = *h->pa;
    cudaSetDevice(h->cfg.device);  // ensure_pin allocates pinned memory from this thread
    std::unique_lock<std::mutex> lk(pa.m);
    for (;;) {
        pa.cv.wait(lk, [&] { return pa.stop || pa.active; });
        if (pa.stop) return;
        while (pa.active) {
            if (pa.stop) return;
            if (pa.cancel || pa.lo >= h->n) { pa.active = false; pa.cv.notify_all(); break; }
            if (pa.produced - pa.released >= kPlanRing) { pa.cv.wait(lk); continue; }
            PlannedChunk& e = pa.ring[pa.produced % kPlanRing];
            PlanInput in = pending_plan_input(h);
            in.max_pairs = ramp_pairs(true, pa.produced);
            const int64_t lo = pa.lo;
            lk.unlock();
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = plan_chunk(h, in, lo, e.plan, nullptr, &e.blob);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            lk.lock();
            e.rc = rc;
            pa.plan_ms += ms;
            pa.produced++;
            if (rc) pa.active = false; else pa.lo = e.plan.hi;
            pa.cv.notify_all();
        }
    }
}

// End the helper's work on the pending batch (normal completion or error) and wait until it is idle.
void plan_ahead_finish(rsa_ext* h) {
    if (!h->plan_ahead) return;
    PlanAhead& pa = *h->pa;
    std::unique_lock<std::mutex> lk(pa.m);
    pa.cancel = true;
    pa.cv.notify_all();
    pa.cv.wait(lk, [&] { return !pa.active; });
    h->stats.host_plan_ms += pa.plan_ms;
    h->plan_ahead = false;
}

// The planning kernels of a device-planned chunk, on their own high-priority stream behind the copy of the offsets:
// their few small blocks slot in as blocks of the previous chunk's DP kernel retire, while the sequence bytes are still
// on their way.
int enqueue_plan_kernels(rsa_ext* h, Slot& s) {
    const ChunkPlan& p = s.plan;
    uint8_t* b = s.d_blob.p;
    cudaStream_t st = h->s_plan;
    CU_TRY(h, cudaStreamWaitEvent(st, s.ev_in, 0));
    CU_TRY(h, cudaMemsetAsync(b + p.off_hdr, 0, p.zero_bytes, st));
    CU_TRY(h, cudaMemsetAsync(b + p.off_groups, 0xFF, sizeof(FastGroup) * (size_t)p.n_group_slots, st));
    PlanArgs a;
    a.qoff = reinterpret_cast<const int64_t*>(b + p.off_in_q);
    a.toff = h->win_off ? nullptr : reinterpret_cast<const int64_t*>(b + p.off_in_t);
    a.win_off = h->win_off ? reinterpret_cast<const int64_t*>(b + p.off_in_t) : nullptr;
    a.win_len = h->win_off ? reinterpret_cast<const int32_t*>(b + p.off_in_wlen) : nullptr;
    a.qtab = reinterpret_cast<const PlanQlen*>(b + p.off_qtab);
    a.n = (int)p.n;
    a.max_tlen = h->cfg.max_target_len;
    a.match = h->sc.match;
    a.exact_only = ((h->cfg.flags & RSA_EXT_FLAG_EXACT_ONLY) != 0 || !h->fast_ok) ? 1 : 0;
    for (int c = 0; c < 3; ++c) a.list_base[c] = p.list_base[c];
    a.meta = reinterpret_cast<PairMeta*>(b + p.off_meta);
    a.info = reinterpret_cast<uint32_t*>(b + p.off_info);
    a.diroff = reinterpret_cast<uint64_t*>(b + p.off_diroff);
    a.list = reinterpret_cast<uint32_t*>(b + p.off_list);
    a.groups = reinterpret_cast<FastGroup*>(b + p.off_groups);
    a.key = reinterpret_cast<uint32_t*>(b + p.off_key);
    a.rank = reinterpret_cast<uint32_t*>(b + p.off_rank);
    a.sorted = reinterpret_cast<uint32_t*>(b + p.off_sorted);
    a.gbytes = reinterpret_cast<uint32_t*>(b + p.off_gbytes);
    a.hist = reinterpret_cast<uint32_t*>(b + p.off_hist);
    a.bsum = reinterpret_cast<uint32_t*>(b + p.off_bsum);
    a.hdr = reinterpret_cast<PlanHeader*>(b + p.off_hdr);
    const unsigned pair_blocks = (unsigned)((p.n + kPlanThreads - 1) / kPlanThreads);
    plan_classify<<<pair_blocks, kPlanThreads, 0, st>>>(a);
    h->stats.kernel_launches++;
    if (p.n_fast_pairs > 0) {
        plan_bin_sums<<<kPlanScanBlocks, 256, 0, st>>>(a);
        plan_bin_scan<<<kPlanScanBlocks, 256, 0, st>>>(a);
        plan_scatter<<<pair_blocks, kPlanThreads, 0, st>>>(a);
        plan_groups<<<(unsigned)((p.n_fast_pairs + kPlanThreads - 1) / kPlanThreads), kPlanThreads, 0, st>>>(a, (int)p.n_fast_pairs);
        h->stats.kernel_launches += 4;
    }
    plan_offsets<<<1, 1024, 0, st>>>(a, p.n_group_slots);
    h->stats.kernel_launches++;
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(s.ev_plan, st));
    return RSA_EXT_OK;
}

int enqueue_chunk(rsa_ext* h, Slot& s) {
    const auto t_plan0 = std::chrono::steady_clock::now();
    const uint8_t* h_blob = nullptr;
    int rc;
    if (h->plan_ahead) {
        PlanAhead& pa = *h->pa;
        std::unique_lock<std::mutex> lk(pa.m);
        pa.cv.wait(lk, [&] { return pa.consumed < pa.produced; });
        PlannedChunk& e = pa.ring[pa.consumed % kPlanRing];
        pa.consumed++;
        if (e.rc) return e.rc;
        s.plan = e.plan;
        h_blob = e.blob.p;  // stays untouched until this chunk retires (retire_chunk releases the entry)
    } else if (h->dev_plan) {
        PlanInput in = pending_plan_input(h);
        in.max_pairs = ramp_pairs(false, h->chunks_enqueued);
        if ((rc = ensure_pin(h, s.h_blob, sizeof(PlanQlen) * kPlanQ))) return rc;
        rc = scan_chunk(h, in, h->next_pair, s.plan, reinterpret_cast<PlanQlen*>(s.h_blob.p));
        if (!rc) {
            if (const int64_t cap = balanced_tail_cap(in.n, h->next_pair, s.plan.hi)) {
                in.max_pairs = cap;
                rc = scan_chunk(h, in, h->next_pair, s.plan, reinterpret_cast<PlanQlen*>(s.h_blob.p));
            }
        }
        h->stats.host_plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
        if (rc) return rc;
        h_blob = s.h_blob.p;
    } else {
        PlanInput in = pending_plan_input(h);
        in.max_pairs = ramp_pairs(false, h->chunks_enqueued);
        rc = plan_chunk(h, in, h->next_pair, s.plan, nullptr, &s.h_blob);
        if (!rc) {
            if (const int64_t cap = balanced_tail_cap(in.n, h->next_pair, s.plan.hi)) {
                in.max_pairs = cap;
                rc = plan_chunk(h, in, h->next_pair, s.plan, nullptr, &s.h_blob);
            }
        }
        h->stats.host_plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
        if (rc) return rc;
        h_blob = s.h_blob.p;
    }
    h->chunks_enqueued++;
    const ChunkPlan& p = s.plan;
    const bool trace = getenv("RSA_EXT_TRACE") != nullptr;
    auto t_last = t_plan0;
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rsa_ext %9.1f chunk %p n=%lld] %-22s %8.3f ms\n", since_load_ms(), (void*)h, (long long)p.n, what,
                std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    lap("plan (+pinned staging)");
    const SlotNeed need{p.blob_bytes, (size_t)p.q_bytes + 16, (size_t)p.t_bytes + 16, sizeof(DpEnd) * (size_t)p.n,
                        sizeof(rsa_ext_result_t) * (size_t)p.n, (size_t)p.arena_bytes,
                        h->alninfo ? sizeof(rsa_ext_alninfo_t) * (size_t)p.n : 0, scratch_alloc_bytes(p.scratch_bytes)};
    if ((rc = ensure_slot(h, s, need))) return rc;
    lap("device buffers");

    if (p.dev) {
        // device planner: the chunk's raw offsets (16 B/pair, straight from the caller's arrays) + the per-length table
        const size_t n1 = (size_t)p.n + 1;
        CU_TRY(h, cudaMemcpyAsync(s.d_blob.p + p.off_in_q, h->qoff + p.lo, sizeof(int64_t) * n1, cudaMemcpyHostToDevice, h->s_h2d));
        if (h->win_off) {
            CU_TRY(h, cudaMemcpyAsync(s.d_blob.p + p.off_in_t, h->win_off + p.lo, sizeof(int64_t) * (size_t)p.n, cudaMemcpyHostToDevice, h->s_h2d));
            CU_TRY(h, cudaMemcpyAsync(s.d_blob.p + p.off_in_wlen, h->win_len + p.lo, sizeof(int32_t) * (size_t)p.n, cudaMemcpyHostToDevice, h->s_h2d));
        } else {
            CU_TRY(h, cudaMemcpyAsync(s.d_blob.p + p.off_in_t, h->toff + p.lo, sizeof(int64_t) * n1, cudaMemcpyHostToDevice, h->s_h2d));
        }
        CU_TRY(h, cudaMemcpyAsync(s.d_blob.p + p.off_qtab, h_blob, sizeof(PlanQlen) * kPlanQ, cudaMemcpyHostToDevice, h->s_h2d));
        CU_TRY(h, cudaEventRecord(s.ev_in, h->s_h2d));
        h->stats.h2d_bytes += (int64_t)(sizeof(int64_t) * n1 * 2 + sizeof(PlanQlen) * kPlanQ);
        if ((rc = enqueue_plan_kernels(h, s))) return rc;
    } else {
        CU_TRY(h, cudaMemcpyAsync(s.d_blob.p, h_blob, p.blob_bytes, cudaMemcpyHostToDevice, h->s_h2d));
        h->stats.h2d_bytes += (int64_t)p.blob_bytes;
    }
    if (p.q_bytes) CU_TRY(h, cudaMemcpyAsync(s.d_q.p, h->qbuf + h->qoff[p.lo], (size_t)p.q_bytes, cudaMemcpyHostToDevice, h->s_h2d));
    if (p.t_bytes) CU_TRY(h, cudaMemcpyAsync(s.d_t.p, h->tbuf + h->toff[p.lo], (size_t)p.t_bytes, cudaMemcpyHostToDevice, h->s_h2d));  // (0 in window form)
    CU_TRY(h, cudaEventRecord(s.ev_h2d, h->s_h2d));
    h->stats.h2d_bytes += p.q_bytes + p.t_bytes;
    lap("h2d enqueue");

    const bool serial = (h->cfg.flags & RSA_EXT_FLAG_SERIALIZE) != 0;
    cudaStream_t s_dp = (!serial && (h->dp_toggle++ & 1)) ? h->s_comp2 : h->s_comp;
    CU_TRY(h, cudaStreamWaitEvent(s_dp, s.ev_h2d, 0));
    if (p.dev) CU_TRY(h, cudaStreamWaitEvent(s_dp, s.ev_plan, 0));
    const uint8_t* d_targets = h->win_off ? h->ref->d : s.d_t.p;  // window form: meta.toff indexes the resident reference
    ChunkDev d{s.d_blob.p, s.d_q.p, d_targets, reinterpret_cast<DpEnd*>(s.d_ends.p),
               reinterpret_cast<rsa_ext_result_t*>(s.d_res.p), s.d_scratch.p, (uint64_t)s.d_scratch.cap, s.d_arena.p,
               s.d_arena_used, (uint64_t)s.d_arena.cap};
    if (h->win_off && !(h->cfg.flags & RSA_EXT_FLAG_ASCII_WINDOWS)) { d.tpack = h->ref->d_pack; d.tflag = h->ref->d_flag; }
    cudaStream_t s_trace = serial ? s_dp : h->s_tb;
    if ((rc = enqueue_compute(h, s_dp, s_trace, s.ev_mid, d, p, nullptr))) return rc;
    if (h->alninfo) {
        finish_kernel<<<(unsigned)((p.n + kFinishThreads - 1) / kFinishThreads), kFinishThreads, 0, s_trace>>>(
            s.d_q.p, d_targets, reinterpret_cast<const PairMeta*>(s.d_blob.p + p.off_meta),
            reinterpret_cast<const rsa_ext_result_t*>(s.d_res.p), (int)p.n, h->sc, h->end_bonus,
            reinterpret_cast<rsa_ext_alninfo_t*>(s.d_aln.p));
        h->stats.kernel_launches++;
    }
    CU_TRY(h, cudaEventRecord(s.ev_comp, s_trace));
    lap("kernel enqueue");

    CU_TRY(h, cudaStreamWaitEvent(h->s_d2h, s.ev_comp, 0));
    CU_TRY(h, cudaMemcpyAsync(h->results + p.lo, s.d_res.p, sizeof(rsa_ext_result_t) * p.n, cudaMemcpyDeviceToHost, h->s_d2h));
    if (h->alninfo) {
        CU_TRY(h, cudaMemcpyAsync(h->alninfo + p.lo, s.d_aln.p, sizeof(rsa_ext_alninfo_t) * p.n, cudaMemcpyDeviceToHost, h->s_d2h));
        h->stats.d2h_bytes += (int64_t)sizeof(rsa_ext_alninfo_t) * p.n;
    }
    CU_TRY(h, cudaMemcpyAsync(s.h_arena_used, s.d_arena_used, kSlotCounters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->s_d2h));
    CU_TRY(h, cudaEventRecord(s.ev_d2h, h->s_d2h));
    h->stats.d2h_bytes += (int64_t)sizeof(rsa_ext_result_t) * p.n + 8;
    lap("d2h enqueue");

    h->stats.pairs_fast += p.n_fast_pairs;
    h->stats.pairs_exact += p.n_exact[0] + p.n_exact[1] + p.n_exact[2];
    h->stats.pairs_failed += p.n_failed;
    h->stats.cells += p.cells;
    s.busy = true;
    h->next_pair = p.hi;
    h->inflight++;
    return RSA_EXT_OK;
}

// After a chunk's D2H finished: pull the long-CIGAR arena (rare) and file the byte strings by pair index.
int retire_chunk(rsa_ext* h, Slot& s) {
    CU_TRY(h, cudaEventSynchronize(s.ev_d2h));
    if (getenv("RSA_EXT_TRACE"))
        fprintf(stderr, "[rsa_ext %9.1f retire %p n=%lld]\n", since_load_ms(), (void*)h, (long long)s.plan.n);
    const unsigned long long used = *s.h_arena_used;
    if (used > 0) {
        // rare: some CIGARs of this chunk are longer than the record's inline bytes.  Their arena offsets were parked
        // in ends[pi].qend/tend by the traceback (kernels_tb.cuh), so the records themselves stay byte-deterministic.
        std::vector<uint8_t> host(used);
        std::vector<DpEnd> ends((size_t)s.plan.n);
        CU_TRY(h, cudaMemcpy(host.data(), s.d_arena.p, used, cudaMemcpyDeviceToHost));
        CU_TRY(h, cudaMemcpy(ends.data(), s.d_ends.p, sizeof(DpEnd) * (size_t)s.plan.n, cudaMemcpyDeviceToHost));
        for (int64_t i = s.plan.lo; i < s.plan.hi; ++i) {
            const rsa_ext_result_t& r = h->results[i];
            if (r.n_ops > RSA_EXT_RLE_INLINE && r.status == 0) {
                const DpEnd& e = ends[(size_t)(i - s.plan.lo)];
                const unsigned long long off = (unsigned long long)(uint32_t)e.qend | ((unsigned long long)(uint32_t)e.tend << 32);
                if (off + (unsigned long long)r.n_ops <= used)
                    h->overflow[i] = std::vector<uint8_t>(host.begin() + off, host.begin() + off + r.n_ops);
            }
        }
    }
    h->stats.pairs_redo += (int64_t)s.h_arena_used[2];
    if (s.h_arena_used[3] > 0) {  // never silently hand back records no kernel computed
        h->err = std::to_string(s.h_arena_used[3]) + " pairs of a chunk were not processed by any kernel (lost kernel launch)";
        s.busy = false;
        h->inflight--;
        if (h->plan_ahead) { std::lock_guard<std::mutex> lk(h->pa->m); h->pa->released++; h->pa->cv.notify_all(); }
        return RSA_EXT_ERR_CUDA;
    }
    if (s.h_arena_used[1] > 0)
        for (int64_t i = s.plan.lo; i < s.plan.hi; ++i)
            if (h->results[i].status == 4) h->retry.push_back(i);
    s.busy = false;
    h->inflight--;
    if (h->plan_ahead) {
        std::lock_guard<std::mutex> lk(h->pa->m);
        h->pa->released++;
        h->pa->cv.notify_all();
    }
    return RSA_EXT_OK;
}

// Targets either as a concatenated buffer with prefix offsets (tbuf/toff) or, window form, as (win_off, win_len)
// into the resident reference.
int submit_core_ex(rsa_ext* h, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf, const int64_t* toff,
                   const int64_t* win_off, const int32_t* win_len, rsa_ext_result_t* results) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is already pending"; return RSA_EXT_ERR_STATE; }
    const bool windows = win_off != nullptr;
    if (n <= 0 || !qbuf || !qoff || !results || (windows ? !win_len : (!tbuf || !toff))) { h->err = "bad argument"; return RSA_EXT_ERR_ARG; }
    // Large batches are validated chunk by chunk by the device planner's host pass (scan_chunk): a separate pass over a
    // million offsets would delay the first kernel by 1-2 ms.  An invalid pair then fails the batch at rsa_ext_wait.
    const bool dev_plan = n >= dev_plan_min_pairs() && (h->cfg.flags & RSA_EXT_FLAG_HOST_PLAN) == 0;
    if (windows) {
        if (!h->ref) { h->err = "no resident reference (rsa_ext_set_reference)"; return RSA_EXT_ERR_STATE; }
        for (int64_t i = 0; i < (dev_plan ? 0 : n); ++i)
            if (win_len[i] < 0 || win_off[i] < 0 || win_off[i] + win_len[i] > h->ref->len) {
                h->err = "window " + std::to_string(i) + " lies outside the resident reference";
                return RSA_EXT_ERR_ARG;
            }
    }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    // the reference validates the whole slice before touching the GPU (gasal2_ssw.cpp:74-89)
    for (int64_t i = 0; i < (dev_plan ? 0 : n); ++i) {
        const int64_t ql = qoff[i + 1] - qoff[i];
        if (ql > h->cfg.max_query_len) {
            h->err = "gasal2 : read size is too big, " + std::to_string(ql) + " > " + std::to_string(h->cfg.max_query_len);
            return RSA_EXT_ERR_QUERY_LEN;
        }
        if (ql < 0 || (!windows && toff[i + 1] < toff[i])) { h->err = "offsets are not monotone"; return RSA_EXT_ERR_ARG; }
    }
    if (h->resident_inflight) {  // an asynchronous resident run may still use the slots' scratch
        CU_TRY(h, cudaStreamSynchronize(h->s_comp));
        CU_TRY(h, cudaStreamSynchronize(h->s_tb));
        h->resident_inflight = false;
    }
    h->n = n; h->qbuf = qbuf; h->qoff = qoff; h->tbuf = tbuf; h->toff = toff; h->results = results;
    h->win_off = win_off; h->win_len = win_len;
    h->alninfo = h->alninfo_next;
    h->next_pair = 0; h->head = 0; h->tail = 0; h->inflight = 0; h->chunks_enqueued = 0;
    h->overflow.clear();
    h->retry.clear();
    h->stats = rsa_ext_stats_t{};
    h->deferred_rc = RSA_EXT_OK;
    h->pending = true;
    std::unique_lock<std::mutex> cold(g_cold_mutex, std::defer_lock);
    if (!h->warmed) cold.lock();
    h->dev_plan = dev_plan;
    if (n > kPlanAheadMinPairs && !h->dev_plan) {
        if (!h->pa) {
            h->pa = new PlanAhead();
            h->pa->th = std::thread(plan_ahead_main, h);
        }
        std::lock_guard<std::mutex> lk(h->pa->m);
        h->pa->lo = 0; h->pa->produced = h->pa->consumed = h->pa->released = 0;
        h->pa->plan_ms = 0;
        h->pa->cancel = false;
        h->pa->active = true;
        h->plan_ahead = true;
        h->pa->cv.notify_all();
    }
    while (h->next_pair < n && h->inflight < kSlots) {
        int rc = enqueue_chunk(h, h->slots[h->tail]);
        if (rc) {
            plan_ahead_finish(h);
            if (h->inflight > 0) { cudaDeviceSynchronize(); for (Slot& sl : h->slots) sl.busy = false; h->inflight = 0; }
            h->pending = false;
            return rc;
        }
        h->tail = (h->tail + 1) % kSlots;
    }
    h->warmed = true;
    return RSA_EXT_OK;
}

int submit_core(rsa_ext* h, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf, const int64_t* toff,
                rsa_ext_result_t* results) {
    return submit_core_ex(h, n, qbuf, qoff, tbuf, toff, nullptr, nullptr, results);
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------

extern "C" int rsa_ext_version(void) { return 1; }

// number of usable CUDA devices (0 when there is no driver/GPU)
extern "C" int rsa_ext_device_count(void) {
    std::lock_guard<std::mutex> cold(g_cold_mutex);
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" const char* rsa_ext_last_error(const rsa_ext_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int rsa_ext_create(const rsa_ext_config_t* cfg_in, rsa_ext_t** out) {
    if (!out) return RSA_EXT_ERR_ARG;
    *out = nullptr;
    rsa_ext_config_t cfg{};
    if (cfg_in) cfg = *cfg_in;
    if (cfg.max_query_len <= 0) cfg.max_query_len = 500;   // gasal2_ssw.h:24
    if (cfg.max_target_len <= 0) cfg.max_target_len = 2000; // gasal2_ssw.h:25
    if (cfg.match == 0 && cfg.mismatch == 0 && cfg.gap_open == 0 && cfg.gap_extend == 0) {
        cfg.match = 2; cfg.mismatch = 8; cfg.gap_open = 12; cfg.gap_extend = 1;  // cmdline.hpp:46-50
    }
    if (cfg.max_query_len > 512 || cfg.max_target_len > kMaxTargetLenCap) {
        g_create_error = "max_query_len must be <= 512 and max_target_len <= 8192";
        return RSA_EXT_ERR_ARG;
    }
    if (cfg.scratch_bytes <= 0) cfg.scratch_bytes = kDefaultScratch;
    std::lock_guard<std::mutex> cold(g_cold_mutex);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
        return RSA_EXT_ERR_CUDA;
    }
    if (cfg.device < 0 || cfg.device >= ndev) { g_create_error = "bad device ordinal"; return RSA_EXT_ERR_ARG; }
    rsa_ext* h = new rsa_ext();
    h->cfg = cfg;
    h->sc.match = cfg.match;
    h->sc.mismatch = cfg.mismatch;
    h->sc.gap_oe = (cfg.gap_open - 1) + cfg.gap_extend;  // gasal2_ssw.cpp:54, gasal_align.cu:332-337
    h->sc.gap_ext = cfg.gap_extend;
    h->scratch_per_slot = (size_t)cfg.scratch_bytes / kSlots;
    h->fk = make_fast_consts(h->sc);
    h->fast_ok = fast_scoring_ok(h->sc);
    auto fail = [&](const char* what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        rsa_ext_destroy(h);
        return RSA_EXT_ERR_CUDA;
    };
    const bool trace = getenv("RSA_EXT_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rsa_ext %9.1f create %p] %-28s %8.3f ms\n", since_load_ms(), (void*)h, what,
                std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    if ((e = cudaSetDevice(cfg.device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaFree(nullptr)) != cudaSuccess) return fail("context", e);
    lap("context");
    if ((e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_comp2, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_cls_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    for (int k = 0; k < kClsStreams; ++k) {
        if ((e = cudaStreamCreateWithFlags(&h->s_cls[k], cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_cls_join[k], cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    }
    {
        // the traceback stream outranks the DP stream: its small blocks are placed first whenever a DP block
        // retires, so the records of chunk k are not held back by the DP kernel of chunk k+1
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char* pe = getenv("RSA_EXT_TB_PRIO");  // experiment knob: 0 = same priority as the DP streams
        if ((e = cudaStreamCreateWithPriority(&h->s_tb, cudaStreamNonBlocking, (pe && atoi(pe) == 0) ? prio_lo : prio_hi)) != cudaSuccess) return fail("stream", e);
    }
    if ((e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if ((e = cudaStreamCreateWithPriority(&h->s_plan, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail("stream", e);
    }
    lap("7 streams");
    for (Slot& s : h->slots) {
        if ((e = cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaEventCreateWithFlags(&s.ev_mid, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaEventCreateWithFlags(&s.ev_plan, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
        if ((e = cudaHostAlloc(&s.h_arena_used, kSlotCounters * sizeof(unsigned long long), cudaHostAllocDefault)) != cudaSuccess) return fail("pinned", e);
        if ((e = cudaMalloc(&s.d_arena_used, kSlotCounters * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc", e);
    }
    lap("events + counters");
    // (cudaGetDeviceProperties costs ~10 ms per call; one attribute is all the engine needs)
    if ((e = cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, cfg.device)) != cudaSuccess)
        return fail("cudaDeviceGetAttribute", e);
    lap("device attribute");
    *out = h;
    return RSA_EXT_OK;
}

extern "C" void rsa_ext_destroy(rsa_ext_t* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->pa) {
        {
            std::lock_guard<std::mutex> lk(h->pa->m);
            h->pa->stop = true;
            h->pa->cv.notify_all();
        }
        h->pa->th.join();
        for (PlannedChunk& e : h->pa->ring)
            if (e.blob.p) cudaFreeHost(e.blob.p);
        delete h->pa;
        h->pa = nullptr;
    }
    if (h->s_comp) cudaStreamSynchronize(h->s_comp);
    if (h->s_comp2) cudaStreamSynchronize(h->s_comp2);
    if (h->s_tb) cudaStreamSynchronize(h->s_tb);
    if (h->s_h2d) cudaStreamSynchronize(h->s_h2d);
    if (h->s_plan) cudaStreamSynchronize(h->s_plan);
    if (h->s_d2h) cudaStreamSynchronize(h->s_d2h);
    for (Slot& s : h->slots) {
        for (DevBuf* b : {&s.d_slab, &s.d_scratch})
            if (b->p) cudaFree(b->p);
        if (s.h_blob.p) cudaFreeHost(s.h_blob.p);
        if (s.h_arena_used) cudaFreeHost(s.h_arena_used);
        if (s.d_arena_used) cudaFree(s.d_arena_used);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_mid) cudaEventDestroy(s.ev_mid);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_plan) cudaEventDestroy(s.ev_plan);
    }
    h->ref.reset();
    for (DevBuf* b : {&h->r_q, &h->r_t, &h->r_res, &h->r_blobs, &h->ham_q, &h->ham_t, &h->ham_off, &h->ham_out})
        if (b->p) cudaFree(b->p);
    for (cudaEvent_t ev : h->ham_ev) if (ev) cudaEventDestroy(ev);
    if (h->ham_pin_in.p) cudaFreeHost(h->ham_pin_in.p);
    if (h->ham_pin_out.p) cudaFreeHost(h->ham_pin_out.p);
    for (cudaEvent_t ev : h->r_events) cudaEventDestroy(ev);
    if (h->own_q.p) cudaFreeHost(h->own_q.p);
    if (h->own_t.p) cudaFreeHost(h->own_t.p);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_comp2) cudaStreamDestroy(h->s_comp2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_cls_fork) cudaEventDestroy(h->ev_cls_fork);
    for (int k = 0; k < kClsStreams; ++k) {
        if (h->s_cls[k]) { cudaStreamSynchronize(h->s_cls[k]); cudaStreamDestroy(h->s_cls[k]); }
        if (h->ev_cls_join[k]) cudaEventDestroy(h->ev_cls_join[k]);
    }
    if (h->s_tb) cudaStreamDestroy(h->s_tb);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->s_plan) cudaStreamDestroy(h->s_plan);
    delete h;
}

extern "C" int rsa_ext_submit(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf,
                              const int64_t* toff, rsa_ext_result_t* results) {
    return submit_core(h, n, qbuf, qoff, tbuf, toff, results);
}

extern "C" int rsa_ext_submit_ptrs(rsa_ext_t* h, int64_t n, const char* const* q, const int32_t* qlen,
                                   const char* const* t, const int32_t* tlen, rsa_ext_result_t* results) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is already pending"; return RSA_EXT_ERR_STATE; }
    if (n <= 0 || !q || !qlen || !t || !tlen || !results) { h->err = "bad argument"; return RSA_EXT_ERR_ARG; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    h->own_qoff.resize(n + 1);
    h->own_toff.resize(n + 1);
    int64_t qs = 0, ts = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (qlen[i] < 0 || tlen[i] < 0) { h->err = "negative length"; return RSA_EXT_ERR_ARG; }
        h->own_qoff[i] = qs; h->own_toff[i] = ts;
        qs += qlen[i]; ts += tlen[i];
    }
    h->own_qoff[n] = qs; h->own_toff[n] = ts;
    int rc;
    if ((rc = ensure_pin(h, h->own_q, (size_t)qs + 16))) return rc;
    if ((rc = ensure_pin(h, h->own_t, (size_t)ts + 16))) return rc;
    // the copy gasal_host_batch_fill does into its pinned pages (host_batch.cpp:137-146), minus the padding
    for (int64_t i = 0; i < n; ++i) {
        memcpy(h->own_q.p + h->own_qoff[i], q[i], (size_t)qlen[i]);
        memcpy(h->own_t.p + h->own_toff[i], t[i], (size_t)tlen[i]);
    }
    return submit_core(h, n, reinterpret_cast<const char*>(h->own_q.p), h->own_qoff.data(),
                       reinterpret_cast<const char*>(h->own_t.p), h->own_toff.data(), results);
}

// ---- SURVEY 8(f) rank 1, first half: windows by (offset, length) into a reference that lives in HBM -------------
//
// The reference builds every window as a std::string (references.sequences[ref_id].substr(...), src/pc.cpp:214-242)
// and the veneer copies those bytes to the GPU again.  With the reference sequence resident, a batch names its windows
// by offset and length: no window bytes are built, gathered or copied (250 of the 450 bytes per pair of a 150 bp batch).
extern "C" int rsa_ext_set_reference(rsa_ext_t* h, const char* seq, int64_t len) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    if (!seq || len <= 0 || len > (int64_t)0xFFFFFFFFll) { h->err = "reference must be 1..2^32-1 bytes"; return RSA_EXT_ERR_ARG; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    std::lock_guard<std::mutex> cold(g_cold_mutex);
    h->ref.reset();
    auto rb = std::make_shared<RefBuf>();
    rb->device = h->cfg.device;
    CU_TRY(h, cudaMalloc(&rb->d, (size_t)len + 16));
    CU_TRY(h, cudaMemcpy(rb->d, seq, (size_t)len, cudaMemcpyHostToDevice));
    // a pageable cudaMemcpy may return once the bytes are staged; the engine's streams are non-blocking, so nothing
    // would order their kernels behind the final DMA
    {   // packed planes: whole 64-base units, one spare unit behind the end (the staging loads whole units)
        const long long units = (len + 63) / 64 + 1;
        CU_TRY(h, cudaMalloc(&rb->d_pack, sizeof(uint4) * (size_t)units));
        CU_TRY(h, cudaMalloc(&rb->d_flag, sizeof(uint2) * (size_t)units));
        const long long n32 = units * 2;
        pack_reference_kernel<<<(unsigned)((n32 + 255) / 256), 256>>>(rb->d, (long long)len, n32, reinterpret_cast<uint32_t*>(rb->d_pack),
                                                                       reinterpret_cast<uint32_t*>(rb->d_flag));
        CU_TRY(h, cudaGetLastError());
    }
    CU_TRY(h, cudaDeviceSynchronize());
    rb->host = seq;
    rb->len = len;
    h->ref = rb;
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_packed_reference(rsa_ext_t* h, uint32_t* codes, uint32_t* flags, int64_t units) {
    if (!h || !codes || !flags) return RSA_EXT_ERR_ARG;
    if (!h->ref) { h->err = "no resident reference (rsa_ext_set_reference)"; return RSA_EXT_ERR_STATE; }
    if (units <= 0 || units > (h->ref->len + 63) / 64) { h->err = "bad unit count"; return RSA_EXT_ERR_ARG; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    CU_TRY(h, cudaMemcpy(codes, h->ref->d_pack, sizeof(uint4) * (size_t)units, cudaMemcpyDeviceToHost));
    CU_TRY(h, cudaMemcpy(flags, h->ref->d_flag, sizeof(uint2) * (size_t)units, cudaMemcpyDeviceToHost));
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_share_reference(rsa_ext_t* h, const rsa_ext_t* donor) {
    if (!h || !donor) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    if (!donor->ref) { h->err = "the donor handle has no resident reference"; return RSA_EXT_ERR_STATE; }
    if (donor->ref->device != h->cfg.device) { h->err = "the donor handle lives on another device"; return RSA_EXT_ERR_ARG; }
    h->ref = donor->ref;
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_submit_ref_windows(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff,
                                          const int64_t* win_off, const int32_t* win_len, rsa_ext_result_t* results) {
    return submit_core_ex(h, n, qbuf, qoff, nullptr, nullptr, win_off, win_len, results);
}

// ---- the Hamming shortcut (src/aln.cpp:391-404, src/aligner.cpp:219-302) on the device ----------
namespace {
int hamming_core(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf, const int64_t* toff,
                 const int64_t* win_off, int32_t end_bonus, int32_t* hamming, rsa_ext_alninfo_t* out) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    if (n <= 0 || !qbuf || !qoff || !hamming || !out) { h->err = "bad argument"; return RSA_EXT_ERR_ARG; }
    if (win_off) {
        if (!h->ref) { h->err = "no resident reference (rsa_ext_set_reference)"; return RSA_EXT_ERR_STATE; }
        for (int64_t i = 0; i < n; ++i) {
            if (win_off[i] < 0 || win_off[i] + (qoff[i + 1] - qoff[i]) > h->ref->len) {
                h->err = "window " + std::to_string(i) + " lies outside the resident reference";
                return RSA_EXT_ERR_ARG;
            }
        }
    } else if (!tbuf || !toff) { h->err = "bad argument"; return RSA_EXT_ERR_ARG; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t qbytes = (size_t)(qoff[n] - qoff[0]), tbytes = win_off ? 0 : (size_t)(toff[n] - toff[0]);
    const size_t offb = sizeof(int64_t) * (size_t)(n + 1);
    int rc;
    if ((rc = ensure_dev(h, h->ham_q, qbytes + 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_t, tbytes + 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_off, 2 * offb + 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_out, (sizeof(rsa_ext_alninfo_t) + sizeof(int32_t)) * (size_t)n + 16))) return rc;
    cudaStream_t st = h->s_comp;
    int64_t* d_qoff = reinterpret_cast<int64_t*>(h->ham_off.p);
    int64_t* d_toff = reinterpret_cast<int64_t*>(h->ham_off.p + offb);
    rsa_ext_alninfo_t* d_out = reinterpret_cast<rsa_ext_alninfo_t*>(h->ham_out.p);
    int32_t* d_ham = reinterpret_cast<int32_t*>(h->ham_out.p + sizeof(rsa_ext_alninfo_t) * (size_t)n);
    // Pageable caller memory goes through the handle's pinned bounce buffers.  A pageable cudaMemcpyAsync is a synchronous,
    // driver-staged copy; with 16 pipeline workers calling at once (integration/hamming_glue.cpp, sam_glue.cpp) those copies
    // made every GPU call of the process slow: BASELINE configs[1] at scale took 14.8 s instead of 6.7 s
    // (profiles/r2_e2e_reads_pe_5m_pairs_100mb_pageable_copies.json vs ..._full_device_path.json).  Pinned callers
    // (bench.py) are copied from / to directly.
    const size_t outb = sizeof(rsa_ext_alninfo_t) * (size_t)n, hamb = sizeof(int32_t) * (size_t)n;
    const size_t woffb = sizeof(int64_t) * (size_t)n;
    const bool out_pinned = host_is_pinned(out) && host_is_pinned(hamming);
    // every input array on its own: pinned ones are read in place, pageable ones take a slot of the bounce buffer
    struct In { const void* p; size_t bytes; bool pinned; size_t at; };
    In ins[4] = {{qbuf + qoff[0], qbytes, false, 0}, {qoff, offb, false, 0},
                 {win_off ? (const void*)win_off : (const void*)toff, win_off ? woffb : offb, false, 0},
                 {win_off ? nullptr : (const void*)(tbuf + toff[0]), tbytes, false, 0}};
    size_t bounce = 0;
    for (In& a : ins) {
        a.pinned = !a.p || !a.bytes || host_is_pinned(a.p);
        if (!a.pinned) { a.at = bounce; bounce += align_up(a.bytes, 16); }
    }
    if (bounce && (rc = ensure_pin(h, h->ham_pin_in, bounce + 16))) return rc;
    for (In& a : ins)
        if (!a.pinned) { memcpy(h->ham_pin_in.p + a.at, a.p, a.bytes); a.p = h->ham_pin_in.p + a.at; }
    const uint8_t *src_q = reinterpret_cast<const uint8_t*>(ins[0].p), *src_t = reinterpret_cast<const uint8_t*>(ins[3].p);
    const void *src_qoff = ins[1].p, *src_toff = ins[2].p;
    if (!out_pinned && (rc = ensure_pin(h, h->ham_pin_out, outb + hamb + 16))) return rc;
    CU_TRY(h, cudaMemcpyAsync(h->ham_q.p, src_q, qbytes, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(d_qoff, src_qoff, offb, cudaMemcpyHostToDevice, st));
    if (win_off) {
        CU_TRY(h, cudaMemcpyAsync(d_toff, src_toff, woffb, cudaMemcpyHostToDevice, st));
    } else {
        CU_TRY(h, cudaMemcpyAsync(h->ham_t.p, src_t, tbytes, cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(d_toff, src_toff, offb, cudaMemcpyHostToDevice, st));
    }
    const int64_t ham_rounds = (n + 31) / 32;   // a warp takes 32 pairs per round
    const int blocks = (int)std::min<int64_t>((ham_rounds + kHamWarpsPerBlock - 1) / kHamWarpsPerBlock, (int64_t)h->n_sms * 16);
    for (cudaEvent_t& ev : h->ham_ev)
        if (!ev) CU_TRY(h, cudaEventCreate(&ev));
    CU_TRY(h, cudaEventRecord(h->ham_ev[0], st));
    // the uploaded sequence slices start at the first pair's offset: rebase the pointers instead of the offsets
    hamming_kernel<<<blocks, 32 * kHamWarpsPerBlock, 0, st>>>(
        h->ham_q.p - qoff[0], d_qoff, win_off ? h->ref->d : h->ham_t.p - toff[0], win_off ? nullptr : d_toff,
        win_off ? d_toff : nullptr, (long long)n, h->sc.match, h->sc.mismatch, end_bonus, d_ham, d_out);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->ham_ev[1], st));
    // (d_out and d_ham are adjacent on the device: one copy when the destination is the bounce buffer)
    if (out_pinned) {
        CU_TRY(h, cudaMemcpyAsync(out, d_out, outb, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(hamming, d_ham, hamb, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
    } else {
        CU_TRY(h, cudaMemcpyAsync(h->ham_pin_out.p, d_out, outb + hamb, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
        memcpy(out, h->ham_pin_out.p, outb);
        memcpy(hamming, h->ham_pin_out.p + outb, hamb);
    }
    h->stats = rsa_ext_stats_t{};
    h->stats.kernel_launches = 1;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ham_ev[0], h->ham_ev[1]) == cudaSuccess) h->stats.dp_ms = ms;   // the Hamming kernel
    h->stats.h2d_bytes = (int64_t)(qbytes + tbytes + (win_off ? sizeof(int64_t) * (size_t)n + offb : 2 * offb));
    h->stats.d2h_bytes = (int64_t)((sizeof(rsa_ext_alninfo_t) + sizeof(int32_t)) * (size_t)n);
    return RSA_EXT_OK;
}
}  // namespace

extern "C" int rsa_ext_hamming_align(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff, const char* tbuf,
                                     const int64_t* toff, int32_t end_bonus, int32_t* hamming, rsa_ext_alninfo_t* out) {
    return hamming_core(h, n, qbuf, qoff, tbuf, toff, nullptr, end_bonus, hamming, out);
}

extern "C" int rsa_ext_hamming_ref_windows(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff,
                                           const int64_t* win_off, int32_t end_bonus, int32_t* hamming,
                                           rsa_ext_alninfo_t* out) {
    return hamming_core(h, n, qbuf, qoff, nullptr, nullptr, win_off, end_bonus, hamming, out);
}

// Allocate, now, what a batch of n pairs of (qlen x tlen) needs from slot 0 and the submit_ptrs staging, so that the
// first real batches of this shape allocate nothing (a cudaMalloc issued while 16 workers are enqueueing has been
// measured at 0.05-0.9 s, during which every other worker's enqueue stalls on the same driver lock).
extern "C" int rsa_ext_reserve(rsa_ext_t* h, int64_t n, int32_t qlen, int32_t tlen) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    if (n <= 0 || qlen <= 0 || tlen <= 0 || qlen > h->cfg.max_query_len || tlen > h->cfg.max_target_len) {
        h->err = "bad argument";
        return RSA_EXT_ERR_ARG;
    }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    std::lock_guard<std::mutex> cold(g_cold_mutex);
    const LenTables& LT = len_tables();
    uint64_t per_pair = (uint64_t)tlen * LT.exact_row_bytes_[qlen] + 16;  // same rule as the chunk budget (plan_chunk)
    if (fast_shape_ok(qlen, tlen, h->sc.match)) per_pair = std::max<uint64_t>(per_pair, (uint64_t)fast_tile_rows(tlen) * LT.fast_row_bytes[qlen]);
    ChunkPlan p;
    const size_t blob = blob_layout(p, n);
    const SlotNeed need{blob, (size_t)n * qlen + 16, (size_t)n * tlen + 16, sizeof(DpEnd) * (size_t)n,
                        sizeof(rsa_ext_result_t) * (size_t)n, (size_t)n * (qlen + tlen + 1) + 64,
                        sizeof(rsa_ext_alninfo_t) * (size_t)n, scratch_alloc_bytes((uint64_t)n * per_pair)};
    int rc;
    if ((rc = ensure_slot(h, h->slots[0], need))) return rc;
    if ((rc = ensure_pin(h, h->slots[0].h_blob, blob))) return rc;
    if ((rc = ensure_pin(h, h->own_q, (size_t)n * qlen + 16))) return rc;
    if ((rc = ensure_pin(h, h->own_t, (size_t)n * tlen + 16))) return rc;
    // the Hamming shortcut of the same batch shape (windows named by offset: no target bytes travel)
    if ((rc = ensure_dev(h, h->ham_q, (size_t)n * qlen + 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_t, 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_off, 2 * sizeof(int64_t) * (size_t)(n + 1) + 16))) return rc;
    if ((rc = ensure_dev(h, h->ham_out, (sizeof(rsa_ext_alninfo_t) + sizeof(int32_t)) * (size_t)n + 16))) return rc;
    if ((rc = ensure_pin(h, h->ham_pin_in, (size_t)n * qlen + 2 * sizeof(int64_t) * (size_t)(n + 1) + 128))) return rc;
    if ((rc = ensure_pin(h, h->ham_pin_out, (sizeof(rsa_ext_alninfo_t) + sizeof(int32_t)) * (size_t)n + 16))) return rc;
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_request_alninfo(rsa_ext_t* h, rsa_ext_alninfo_t* out, int32_t end_bonus) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    h->alninfo_next = out;
    h->end_bonus = end_bonus;
    return RSA_EXT_OK;
}

// Retire the head chunk and enqueue the next one.  block == false: only when the head chunk's copies have landed.
// Returns 1 = progress, 0 = nothing ready (non-blocking only) or nothing in flight, < 0 = error.
static int advance(rsa_ext* h, bool block) {
    if (h->inflight == 0) return 0;
    Slot& s = h->slots[h->head];
    if (!block) {
        const cudaError_t q = cudaEventQuery(s.ev_d2h);
        if (q == cudaErrorNotReady) { (void)cudaGetLastError(); return 0; }
    }
    int rc = retire_chunk(h, s);
    if (rc) return rc;
    h->head = (h->head + 1) % kSlots;
    if (h->next_pair < h->n) {
        if ((rc = enqueue_chunk(h, h->slots[h->tail]))) return rc;
        h->tail = (h->tail + 1) % kSlots;
    }
    return 1;
}

// Like gasal_is_aln_async_done in the reference's `while (poll) usleep` loop (src/gasal2_ssw.cpp:179), poll must make
// progress by itself: a batch of more chunks than slots only advances when finished chunks are retired and the next
// ones enqueued, so poll does exactly that (without blocking).  0 = every chunk has landed (rsa_ext_wait will not
// block on the GPU any more; it still has to be called: it finalises the batch and reports errors), 1 = still running.
extern "C" int rsa_ext_poll(rsa_ext_t* h) {
    if (!h || !h->pending) return 0;
    if (h->deferred_rc) return 0;  // rsa_ext_wait reports it
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 0;
    for (;;) {
        const int r = advance(h, false);
        if (r < 0) { h->deferred_rc = r; return 0; }
        if (r == 0) break;
    }
    return (h->inflight > 0 || h->next_pair < h->n) ? 1 : 0;
}

// Pairs the redo pass could not re-tile (status 4; only possible when a chunk holds more symbols outside
// ACGTN than the scratch head-room covers): run them again as their own exact-only batch and splice the
// records and long-CIGAR strings back.
static int run_retry(rsa_ext* h) {
    const std::vector<int64_t> idx = h->retry;
    const int64_t m = (int64_t)idx.size();
    // own staging: the pending batch's buffers may be the handle's submit_ptrs staging (own_q/own_t), which a nested
    // submit_ptrs would overwrite while reading from it
    std::vector<int64_t> qo((size_t)m + 1), to((size_t)m + 1);
    qo[0] = to[0] = 0;
    for (int64_t k = 0; k < m; ++k) {
        const int64_t i = idx[k];
        qo[k + 1] = qo[k] + (h->qoff[i + 1] - h->qoff[i]);
        to[k + 1] = to[k] + (h->win_off ? (int64_t)h->win_len[i] : h->toff[i + 1] - h->toff[i]);
    }
    std::vector<char> qs((size_t)qo[m] + 16), ts((size_t)to[m] + 16);
    for (int64_t k = 0; k < m; ++k) {
        const int64_t i = idx[k];
        memcpy(qs.data() + qo[k], h->qbuf + h->qoff[i], (size_t)(qo[k + 1] - qo[k]));
        const char* tsrc = h->win_off ? h->ref->host + h->win_off[i] : h->tbuf + h->toff[i];
        memcpy(ts.data() + to[k], tsrc, (size_t)(to[k + 1] - to[k]));
    }
    std::vector<rsa_ext_result_t> tmp(m);
    rsa_ext_result_t* results = h->results;
    rsa_ext_alninfo_t* const aln_out = h->alninfo;
    rsa_ext_alninfo_t* const aln_req = h->alninfo_next;
    std::vector<rsa_ext_alninfo_t> tmpa(aln_out ? m : 0);
    h->alninfo_next = aln_out ? tmpa.data() : nullptr;
    auto overflow = std::move(h->overflow);
    const rsa_ext_stats_t stats = h->stats;
    const int32_t flags = h->cfg.flags;
    h->cfg.flags |= RSA_EXT_FLAG_EXACT_ONLY;
    int rc = submit_core(h, m, qs.data(), qo.data(), ts.data(), to.data(), tmp.data());
    if (rc == RSA_EXT_OK) rc = rsa_ext_wait(h);
    h->cfg.flags = flags;
    h->alninfo_next = aln_req;
    h->alninfo = aln_out;
    if (rc != RSA_EXT_OK) return rc;
    for (int64_t k = 0; k < m; ++k) {
        results[idx[k]] = tmp[k];
        if (aln_out) aln_out[idx[k]] = tmpa[k];
        auto it = h->overflow.find(k);
        if (it != h->overflow.end()) overflow[idx[k]] = std::move(it->second);
    }
    h->overflow = std::move(overflow);
    rsa_ext_stats_t s2 = h->stats;
    h->stats = stats;
    h->stats.kernel_launches += s2.kernel_launches;
    h->stats.pairs_exact += s2.pairs_exact;
    h->stats.h2d_bytes += s2.h2d_bytes;
    h->stats.d2h_bytes += s2.d2h_bytes;
    h->results = results;
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_wait(rsa_ext_t* h) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (!h->pending) { h->err = "nothing submitted"; return RSA_EXT_ERR_STATE; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = h->deferred_rc;  // an error rsa_ext_poll ran into
    h->deferred_rc = RSA_EXT_OK;
    while (rc == RSA_EXT_OK && h->inflight > 0) {
        const int r = advance(h, true);
        if (r < 0) rc = r;
    }
    plan_ahead_finish(h);
    if (rc) {
        cudaDeviceSynchronize();
        for (Slot& s : h->slots) s.busy = false;
        h->inflight = 0;
    }
    h->pending = false;
    if (rc == RSA_EXT_OK && !h->retry.empty()) rc = run_retry(h);
    return rc;
}

extern "C" int rsa_ext_rle_overflow(rsa_ext_t* h, int64_t i, uint8_t* out, int32_t cap) {
    if (!h || !out) return RSA_EXT_ERR_ARG;
    auto it = h->overflow.find(i);
    if (it == h->overflow.end()) { h->err = "no overflow record for this pair"; return RSA_EXT_ERR_ARG; }
    const int n = (int)std::min<size_t>(it->second.size(), (size_t)std::max(cap, 0));
    memcpy(out, it->second.data(), (size_t)n);
    return n;
}

extern "C" int rsa_ext_rle_to_text(const uint8_t* rle, int32_t n_ops, char* out, int32_t cap) {
    // src/gasal2_ssw.cpp:184-243: bytes last-to-first, equal neighbours merged
    static const char opc[4] = {'M', 'X', 'D', 'I'};
    if (cap <= 0 || !out) return -1;
    int pos = 0;
    out[0] = 0;
    if (n_ops <= 0 || !rle) return 0;
    int last_op = rle[n_ops - 1] & 3;
    long count = rle[n_ops - 1] >> 2;
    auto put = [&](long cnt, int op) -> bool {
        char tmp[24];
        int w = snprintf(tmp, sizeof tmp, "%ld%c", cnt, opc[op]);
        if (pos + w >= cap) return false;
        memcpy(out + pos, tmp, (size_t)w);
        pos += w;
        out[pos] = 0;
        return true;
    };
    for (int u = n_ops - 2; u >= 0; --u) {
        const int op = rle[u] & 3;
        if (op == last_op) count += rle[u] >> 2;
        else {
            if (!put(count, last_op)) return -1;
            count = rle[u] >> 2;
        }
        last_op = op;
    }
    if (!put(count, last_op)) return -1;
    return pos;
}

extern "C" void* rsa_ext_stream(rsa_ext_t* h) { return h ? (void*)h->s_comp : nullptr; }

extern "C" int rsa_ext_get_stats(const rsa_ext_t* hc, rsa_ext_stats_t* out) {
    if (!hc || !out) return RSA_EXT_ERR_ARG;
    rsa_ext* h = const_cast<rsa_ext*>(hc);
    if (h->r_events_valid) {
        cudaSetDevice(h->cfg.device);
        cudaStreamSynchronize(h->s_comp);
        cudaStreamSynchronize(h->s_comp2);
        cudaStreamSynchronize(h->s_tb);
        double dp = 0, tb = 0;
        for (size_t c = 0; c < h->res_chunks.size(); ++c) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, h->r_events[4 * c], h->r_events[4 * c + 1]);
            cudaEventElapsedTime(&b, h->r_events[4 * c + 2], h->r_events[4 * c + 3]);
            dp += a; tb += b;
        }
        h->stats.dp_ms = dp;
        h->stats.tb_ms = tb;
        if (getenv("RSA_EXT_TRACE")) {  // timeline of the last resident run, ms after the first chunk's DP start
            for (size_t c = 0; c < h->res_chunks.size(); ++c) {
                float t[4];
                for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], h->r_events[0], h->r_events[4 * c + i]);
                fprintf(stderr, "[rsa_ext timeline] chunk %2zu n=%6lld  dp %7.3f..%7.3f  tb %7.3f..%7.3f\n", c,
                        (long long)h->res_chunks[c].plan.n, t[0], t[1], t[2], t[3]);
            }
        }
        unsigned long long c3[3] = {0, 0, 0};
        cudaMemcpy(c3, h->slots[0].d_arena_used, sizeof c3, cudaMemcpyDeviceToHost);
        h->stats.pairs_redo = (int64_t)c3[2];  // last chunk only
    }
    *out = h->stats;
    return RSA_EXT_OK;
}

// ---- resident legs ------------------------------------------------------------------------------------

extern "C" int rsa_ext_stage_resident(rsa_ext_t* h, int64_t n, const char* qbuf, const int64_t* qoff,
                                      const char* tbuf, const int64_t* toff) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    if (n <= 0 || !qbuf || !qoff || !tbuf || !toff) { h->err = "bad argument"; return RSA_EXT_ERR_ARG; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    CU_TRY(h, cudaDeviceSynchronize());
    h->res_chunks.clear();
    h->r_events_valid = false;
    h->r_n = n;
    h->stats = rsa_ext_stats_t{};
    PlanInput in{n, qoff, toff, qbuf, tbuf, h->cfg.max_query_len, h->cfg.max_target_len, h->scratch_per_slot,
                 (h->cfg.flags & RSA_EXT_FLAG_EXACT_ONLY) != 0 || !h->fast_ok};
    in.match = h->sc.match;
    std::vector<std::vector<uint8_t>> blobs;
    int64_t lo = 0;
    size_t blob_total = 0, max_pairs = 0, max_scratch = 0, max_arena = 0;
    while (lo < n) {
        ResidentChunk rc_;
        blobs.emplace_back();
        int rc = plan_chunk(h, in, lo, rc_.plan, &blobs.back(), nullptr);
        if (rc) return rc;
        if (const int64_t cap = balanced_tail_cap(n, lo, rc_.plan.hi)) {
            PlanInput in2 = in;
            in2.max_pairs = cap;
            blobs.back().clear();
            if ((rc = plan_chunk(h, in2, lo, rc_.plan, &blobs.back(), nullptr))) return rc;
        }
        rc_.q_base = (size_t)(qoff[lo] - qoff[0]);
        rc_.t_base = (size_t)(toff[lo] - toff[0]);
        blob_total += align_up(rc_.plan.blob_bytes, 256);
        max_pairs = std::max(max_pairs, (size_t)rc_.plan.n);
        max_scratch = std::max(max_scratch, (size_t)rc_.plan.scratch_bytes);
        max_arena = std::max(max_arena, (size_t)rc_.plan.arena_bytes);
        lo = rc_.plan.hi;
        h->res_chunks.push_back(rc_);
    }
    int rc;
    const size_t qbytes = (size_t)(qoff[n] - qoff[0]), tbytes = (size_t)(toff[n] - toff[0]);
    if ((rc = ensure_dev(h, h->r_q, qbytes + 16))) return rc;
    if ((rc = ensure_dev(h, h->r_t, tbytes + 16))) return rc;
    if ((rc = ensure_dev(h, h->r_res, sizeof(rsa_ext_result_t) * (size_t)n))) return rc;
    if ((rc = ensure_dev(h, h->r_blobs, blob_total))) return rc;
    // resident chunks rotate through the slots' scratch: with three, DP(c) never waits for the traceback of the
    // chunk that used its slot last (with two, DP(c) and DP(c+1) ran side by side and both waited: no overlap)
    for (int k = 0; k < (int)std::min<size_t>(h->res_chunks.size(), kSlots); ++k) {
        Slot& s = h->slots[k];
        h->r_need = SlotNeed{0, 0, 0, sizeof(DpEnd) * (size_t)max_pairs, 0, (size_t)max_arena, 0, scratch_alloc_bytes(max_scratch)};
        if ((rc = ensure_slot(h, s, h->r_need))) return rc;
    }
    CU_TRY(h, cudaMemcpy(h->r_q.p, qbuf + qoff[0], qbytes, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->r_t.p, tbuf + toff[0], tbytes, cudaMemcpyHostToDevice));
    size_t boff = 0;
    for (size_t c = 0; c < h->res_chunks.size(); ++c) {
        h->res_chunks[c].d_blob = h->r_blobs.p + boff;
        CU_TRY(h, cudaMemcpy(h->r_blobs.p + boff, blobs[c].data(), blobs[c].size(), cudaMemcpyHostToDevice));
        boff += align_up(h->res_chunks[c].plan.blob_bytes, 256);
        h->stats.pairs_fast += h->res_chunks[c].plan.n_fast_pairs;
        h->stats.pairs_exact += h->res_chunks[c].plan.n_exact[0] + h->res_chunks[c].plan.n_exact[1] + h->res_chunks[c].plan.n_exact[2];
        h->stats.pairs_failed += h->res_chunks[c].plan.n_failed;
        h->stats.cells += h->res_chunks[c].plan.cells;
    }
    h->stats.h2d_bytes = (int64_t)(qbytes + tbytes + blob_total);
    CU_TRY(h, cudaDeviceSynchronize());  // pageable uploads above vs the non-blocking streams (see rsa_ext_set_reference)
    while (h->r_events.size() < 4 * h->res_chunks.size()) {
        cudaEvent_t ev;
        CU_TRY(h, cudaEventCreate(&ev));
        h->r_events.push_back(ev);
    }
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_run_resident(rsa_ext_t* h) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (h->res_chunks.empty()) { h->err = "nothing staged"; return RSA_EXT_ERR_STATE; }
    if (h->pending) { h->err = "a batch is pending"; return RSA_EXT_ERR_STATE; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    h->stats.kernel_launches = 0;
    const int nslots = (int)std::min<size_t>(h->res_chunks.size(), kSlots);
    // a submit since rsa_ext_stage_resident may have re-carved the slots' buffers
    for (int k = 0; k < nslots; ++k) { int rc = ensure_slot(h, h->slots[k], h->r_need); if (rc) return rc; }
    h->resident_inflight = true;
    // whatever the caller recorded on the handle's stream (rsa_ext_stream) precedes the work on both DP streams
    CU_TRY(h, cudaEventRecord(h->ev_fork, h->s_comp));
    CU_TRY(h, cudaStreamWaitEvent(h->s_comp2, h->ev_fork, 0));
    for (size_t c = 0; c < h->res_chunks.size(); ++c) {
        const ResidentChunk& rcx = h->res_chunks[c];
        Slot& s = h->slots[c % nslots];
        const bool serial = (h->cfg.flags & RSA_EXT_FLAG_SERIALIZE) != 0;
        cudaStream_t s_dp = (!serial && (c & 1)) ? h->s_comp2 : h->s_comp;
        cudaStream_t s_trace = serial ? s_dp : h->s_tb;
        // the slot's scratch/ends are free once the traceback that last used them has finished
        CU_TRY(h, cudaStreamWaitEvent(s_dp, s.ev_comp, 0));
        ChunkDev d{rcx.d_blob, h->r_q.p + rcx.q_base, h->r_t.p + rcx.t_base, reinterpret_cast<DpEnd*>(s.d_ends.p),
                   reinterpret_cast<rsa_ext_result_t*>(h->r_res.p) + rcx.plan.lo, s.d_scratch.p, (uint64_t)s.d_scratch.cap,
                   s.d_arena.p, s.d_arena_used, (uint64_t)s.d_arena.cap};
        int rc = enqueue_compute(h, s_dp, s_trace, s.ev_mid, d, rcx.plan, &h->r_events[4 * c]);
        if (rc) return rc;
        CU_TRY(h, cudaEventRecord(s.ev_comp, s_trace));
    }
    // anything the caller records on the compute stream after this call comes after every traceback too
    for (int k = 0; k < nslots; ++k) CU_TRY(h, cudaStreamWaitEvent(h->s_comp, h->slots[k].ev_comp, 0));
    h->r_events_valid = true;
    return RSA_EXT_OK;
}

extern "C" int rsa_ext_fetch_resident(rsa_ext_t* h, rsa_ext_result_t* results) {
    if (!h || !results) return RSA_EXT_ERR_ARG;
    if (h->res_chunks.empty()) { h->err = "nothing staged"; return RSA_EXT_ERR_STATE; }
    CU_TRY(h, cudaSetDevice(h->cfg.device));
    CU_TRY(h, cudaStreamSynchronize(h->s_comp));
    CU_TRY(h, cudaStreamSynchronize(h->s_comp2));
    CU_TRY(h, cudaStreamSynchronize(h->s_tb));
    CU_TRY(h, cudaMemcpy(results, h->r_res.p, sizeof(rsa_ext_result_t) * (size_t)h->r_n, cudaMemcpyDeviceToHost));
    for (int k = 0; k < (int)std::min<size_t>(h->res_chunks.size(), kSlots); ++k) {  // counters of each slot's last chunk
        unsigned long long c[kSlotCounters] = {};
        CU_TRY(h, cudaMemcpy(c, h->slots[k].d_arena_used, sizeof c, cudaMemcpyDeviceToHost));
        if (c[3] > 0) {
            h->err = std::to_string(c[3]) + " pairs of a chunk were not processed by any kernel (lost kernel launch)";
            return RSA_EXT_ERR_CUDA;
        }
    }
    return RSA_EXT_OK;
}

// Host-only planning probe for tests (no CUDA call is made): plans the first chunk of a batch and reports
// how pairs were routed.  out[0]=pairs in chunk, [1]=fast pairs, [2]=exact pairs, [3]=failed, [4]=groups,
// [5]=scratch bytes, [6]=fast classes.
extern "C" int rsa_ext_plan_debug(int64_t n, const int64_t* qoff, const int64_t* toff, int64_t scratch_cap,
                                  int exact_only, int64_t* out) {
    rsa_ext h;
    PlanInput in{n, qoff, toff, nullptr, nullptr, 500, 2000, (size_t)scratch_cap, exact_only != 0};
    ChunkPlan p;
    std::vector<uint8_t> blob;
    int rc = plan_chunk(&h, in, 0, p, &blob, nullptr);
    if (rc) return rc;
    // out[7] (in/out): if > 0 on entry, re-plan that many times with warm buffers and report the mean ns
    if (out[7] > 0) {
        const int reps = (int)out[7];
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) plan_chunk(&h, in, 0, p, &blob, nullptr);
        out[7] = (int64_t)(std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / reps);
    }
    int groups = 0;
    for (auto& fc : p.fast) groups += fc.n_groups;
    out[0] = p.n; out[1] = p.n_fast_pairs; out[2] = p.n_exact[0] + p.n_exact[1] + p.n_exact[2];
    out[3] = p.n_failed; out[4] = groups; out[5] = (int64_t)p.scratch_bytes; out[6] = p.n_fast_classes;
    return RSA_EXT_OK;
}

// Host-only probe of the device planner's host pass (scan_chunk; no CUDA call): out[0]=pairs in the first chunk,
// [1]=packed candidates, [2]=exact pairs, [3]=not aligned, [4]=group slots, [5]=scratch bound, [6]=column classes,
// [7] in/out: repetitions -> mean ns per pass.
extern "C" int rsa_ext_scan_debug(int64_t n, const int64_t* qoff, const int64_t* toff, int64_t scratch_cap, int64_t* out) {
    rsa_ext h;
    PlanInput in{n, qoff, toff, nullptr, nullptr, 500, 2000, (size_t)scratch_cap, false};
    ChunkPlan p;
    std::vector<PlanQlen> qtab(kPlanQ);
    int rc = scan_chunk(&h, in, 0, p, qtab.data());
    if (rc) return rc;
    if (out[7] > 0) {
        const int reps = (int)out[7];
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) scan_chunk(&h, in, 0, p, qtab.data());
        out[7] = (int64_t)(std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / reps);
    }
    out[0] = p.n; out[1] = p.n_fast_pairs; out[2] = p.n_exact[0] + p.n_exact[1] + p.n_exact[2];
    out[3] = p.n_failed; out[4] = p.n_group_slots; out[5] = (int64_t)p.scratch_bytes; out[6] = p.n_fast_classes;
    return RSA_EXT_OK;
}
