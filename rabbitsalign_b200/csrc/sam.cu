// sam.cu -- device-side SAM record formatter behind include/rsa_sam.h.
//
// The reference appends every field of every record to a std::string on the worker thread (class Sam,
// /root/reference/src/sam.cpp:31-206).  Here a batch of record descriptors is formatted by two kernels:
//   sam_length_kernel   one thread per record: the exact byte length of its line (decimal widths, CIGAR text width
//                       with the M-operation merge of Cigar::to_m, '*' substitutions, tags, tail);
//   (exclusive scan of the lengths: three small kernels)
//   sam_write_kernel    one warp per record: the small fields digit by digit (lane k writes digit k), name / SEQ / QUAL
//                       as coalesced byte copies, reverse-complemented (src/revcomp.hpp:10-41) or reversed for
//                       reverse-strand records.
// HBM-bound byte work: ~350 B read and ~400 B written per 150-bp record.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/rsa_ext.h"
#include "../../include/rsa_sam.h"

namespace {

enum : uint32_t { F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8, F_REVERSE = 0x10, F_MREVERSE = 0x20, F_READ1 = 0x40,
                  F_READ2 = 0x80, F_SECONDARY = 0x100 };

struct SamCfg {
    const char* names;          // reference names, device
    const long long* names_off; // n_refs + 1
    int n_refs;
    int cigar_m, output_unmapped, show_details;
    int tail_len;
    char tail[64];              // "\n" or "\tRG:Z:<id>\n"
};

// literal pieces of the unmapped forms (src/sam.cpp:81,103-107; SAM_UNMAPPED_MAPQ_STRING = "0")
#define kUnmappedMiddle "\t*\t0\t0\t*\t*\t0\t0\t"
#define kMateMiddle "\t0\t*\t=\t"
#define kMateTail "\t0\t"
#define SAM_LIT(o, s) (o).lit((s), (int)sizeof(s) - 1)

__host__ __device__ inline int dec_width_u(unsigned long long v) {
    int w = 1;
    while (v >= 10ull) { v /= 10ull; ++w; }
    return w;
}
__host__ __device__ inline int dec_width_s(long long v) { return v < 0 ? 1 + dec_width_u((unsigned long long)(-v)) : dec_width_u((unsigned long long)v); }

// strip_suffix (src/sam.cpp:31-43)
__device__ inline uint32_t stripped_len(const char* name, uint32_t len) {
    if (len >= 2 && name[len - 2] == '/' && (name[len - 1] == '1' || name[len - 1] == '2')) return len - 2;
    return len;
}

// CIGAR as printed: Cigar::to_string (src/cigar.cpp:47-53); with cigar_m first Cigar::to_m (:6-18: = and X become M and
// equal neighbours merge through Cigar::push).  `emit(len, op_char)` is called once per printed operation.
template <class F>
__device__ inline void walk_cigar(const uint32_t* ops, uint32_t n, int cigar_m, F emit) {
    const char* letters = "MIDNSHP=X";
    if (!cigar_m) {
        for (uint32_t k = 0; k < n; ++k) emit(ops[k] >> 4, letters[ops[k] & 0xFu]);
        return;
    }
    uint32_t cur_op = 0xFFu;
    unsigned long long cur_len = 0;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t op = ops[k] & 0xFu;
        if (op == 7u || op == 8u) op = 0u;
        if (op == cur_op) { cur_len += ops[k] >> 4; continue; }
        if (cur_op != 0xFFu) emit((uint32_t)cur_len, letters[cur_op]);
        cur_op = op;
        cur_len = ops[k] >> 4;
    }
    if (cur_op != 0xFFu) emit((uint32_t)cur_len, letters[cur_op]);
}

__device__ inline int name_len_of(const SamCfg& c, int ref) {   // RNAME / RNEXT text width
    if (ref == RSA_SAM_REF_SAME || ref == RSA_SAM_REF_NONE || ref < 0 || ref >= c.n_refs) return 1;
    return (int)(c.names_off[ref + 1] - c.names_off[ref]);
}

__global__ void sam_length_kernel(SamCfg c, const rsa_sam_record_t* __restrict__ rec, long long n, const char* __restrict__ text,
                                  const uint32_t* __restrict__ cigars, unsigned long long* __restrict__ len_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rsa_sam_record_t r = rec[i];
    unsigned long long L = 0;
    const uint32_t nm = stripped_len(text + r.name_off, r.name_len);
    const int seq_w = r.seq_len ? (int)r.seq_len : 1, qual_w = r.qual_len ? (int)r.qual_len : 1;
    if (r.kind == RSA_SAM_UNMAPPED) {
        if (c.output_unmapped) L = nm + 1 + dec_width_u(r.flags) + (sizeof(kUnmappedMiddle) - 1) + seq_w + 1 + qual_w + c.tail_len;
    } else if (r.kind == RSA_SAM_UNMAPPED_MATE) {
        const int pw = dec_width_u((uint32_t)(r.mate_pos + 1u));
        L = nm + 1 + dec_width_u(r.flags) + 1 + name_len_of(c, r.ref_id) + 1 + pw + (sizeof(kMateMiddle) - 1) + pw + (sizeof(kMateTail) - 1) +
            seq_w + 1 + qual_w + c.tail_len;
    } else {
        unsigned long long cw = 0;
        if (r.n_cigar == 0) cw = 1;
        else walk_cigar(cigars + r.cigar_off, r.n_cigar, c.cigar_m, [&](uint32_t len, char) { cw += dec_width_u(len) + 1; });
        const bool secondary = r.flags & F_SECONDARY;
        L = nm + 1 + dec_width_u(r.flags) + 1 + name_len_of(c, r.ref_id) + 1 + dec_width_u((uint32_t)(r.pos + 1u)) + 1 +
            dec_width_u(r.mapq & 0xFFu) + 1 + cw + 1 + name_len_of(c, r.mate_ref) + 1 + dec_width_u((uint32_t)(r.mate_pos + 1u)) + 1 +
            dec_width_s(r.tlen) + 1 + (secondary ? 1 : seq_w) + 1 + (secondary ? 1 : qual_w);
        if (!(r.flags & F_UNMAP)) L += 6 + dec_width_s(r.edit_distance) + 6 + dec_width_s(r.score);   // "\tNM:i:" "\tAS:i:"
        if (c.show_details) {
            L += 6 + dec_width_u(r.details[0]) + 6 + dec_width_u(r.details[1]) + 6 + dec_width_u(r.details[2]) + 6 + dec_width_u(r.details[3]);
            if (r.flags & F_PAIRED) L += 6 + dec_width_u(r.details[4]);
        }
        L += c.tail_len;
    }
    len_out[i] = L;
}

// ---- exclusive scan of n 64-bit values: per-block sums, scan of the sums by one block, per-block rescan ----------
constexpr int kScanThreads = 256, kScanItems = 8;
__global__ void scan_block_sums(const unsigned long long* __restrict__ v, long long n, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long sh[kScanThreads];
    const long long base = (long long)blockIdx.x * kScanThreads * kScanItems;
    unsigned long long s = 0;
    for (int k = 0; k < kScanItems; ++k) { const long long i = base + (long long)threadIdx.x * kScanItems + k; if (i < n) s += v[i]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = kScanThreads / 2; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) sums[blockIdx.x] = sh[0];
}
__global__ void scan_of_sums(unsigned long long* sums, int nb, unsigned long long* total) {   // one thread: nb is small (n / 2048)
    unsigned long long run = 0;
    for (int b = 0; b < nb; ++b) { const unsigned long long s = sums[b]; sums[b] = run; run += s; }
    *total = run;
}
__global__ void scan_apply(const unsigned long long* __restrict__ v, long long n, const unsigned long long* __restrict__ sums,
                           unsigned long long* __restrict__ off) {
    __shared__ unsigned long long sh[kScanThreads];
    const long long base = (long long)blockIdx.x * kScanThreads * kScanItems;
    unsigned long long loc[kScanItems], s = 0;
    for (int k = 0; k < kScanItems; ++k) { const long long i = base + (long long)threadIdx.x * kScanItems + k; loc[k] = i < n ? v[i] : 0; s += loc[k]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < kScanThreads; o <<= 1) {   // Hillis-Steele inclusive scan of the thread sums
        unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned long long run = sums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < kScanItems; ++k) { const long long i = base + (long long)threadIdx.x * kScanItems + k; if (i < n) off[i] = run; run += loc[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) off[n] = sums[blockIdx.x] + sh[threadIdx.x];
}

// ---- writer: one warp per record; every lane keeps the same cursor -------------------------------------------------
struct Cursor {
    char* p;
    int lane;
    __device__ void ch(char c) { if (lane == 0) *p = c; ++p; }
    __device__ void lit(const char* s, int n) { if (lane < n) p[lane] = s[lane]; p += n; }   // n <= 32
    __device__ void bytes(const char* s, long long n) { for (long long i = lane; i < n; i += 32) p[i] = s[i]; p += n; }
    __device__ void dec_u(unsigned long long v) {
        const int w = dec_width_u(v);
        if (lane < w) { unsigned long long d = v; for (int k = 0; k < w - 1 - lane; ++k) d /= 10ull; p[lane] = (char)('0' + (int)(d % 10ull)); }
        p += w;
    }
    __device__ void dec_s(long long v) { if (v < 0) { ch('-'); dec_u((unsigned long long)(-v)); } else dec_u((unsigned long long)v); }
};

__device__ inline char revcomp_char(unsigned char ch) {   // revcomp_table (src/revcomp.hpp:10-27)
    switch (ch) {
        case 'A': case 'a': return 'T';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        case 'T': case 't': case 'U': case 'u': return 'A';
        default: return 'N';
    }
}

__device__ inline void put_ref_name(Cursor& o, const SamCfg& c, int ref) {
    if (ref == RSA_SAM_REF_SAME) o.ch('=');
    else if (ref < 0 || ref >= c.n_refs) o.ch('*');
    else o.bytes(c.names + c.names_off[ref], c.names_off[ref + 1] - c.names_off[ref]);
}

constexpr int kSamWarps = 4;
__global__ void __launch_bounds__(32 * kSamWarps)
sam_write_kernel(SamCfg c, const rsa_sam_record_t* __restrict__ rec, long long n, const char* __restrict__ text,
                 const uint32_t* __restrict__ cigars, const unsigned long long* __restrict__ off, char* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * kSamWarps + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * kSamWarps) {
        const rsa_sam_record_t r = rec[i];
        if (off[i + 1] == off[i]) continue;   // unmapped record with output_unmapped off
        Cursor o{out + off[i], lane};
        const char* name = text + r.name_off;
        o.bytes(name, stripped_len(name, r.name_len));
        o.ch('\t');
        o.dec_u(r.flags);
        const char* seq = text + r.seq_off;
        const char* qual = text + r.qual_off;
        if (r.kind == RSA_SAM_UNMAPPED) {
            SAM_LIT(o, kUnmappedMiddle);
            if (r.seq_len) o.bytes(seq, r.seq_len); else o.ch('*');
            o.ch('\t');
            if (r.qual_len) o.bytes(qual, r.qual_len); else o.ch('*');
        } else if (r.kind == RSA_SAM_UNMAPPED_MATE) {
            o.ch('\t');
            put_ref_name(o, c, r.ref_id);
            o.ch('\t');
            o.dec_u((uint32_t)(r.mate_pos + 1u));
            SAM_LIT(o, kMateMiddle);
            o.dec_u((uint32_t)(r.mate_pos + 1u));
            SAM_LIT(o, kMateTail);
            if (r.seq_len) o.bytes(seq, r.seq_len); else o.ch('*');
            o.ch('\t');
            if (r.qual_len) o.bytes(qual, r.qual_len); else o.ch('*');
        } else {
            o.ch('\t');
            put_ref_name(o, c, r.ref_id);
            o.ch('\t');
            o.dec_u((uint32_t)(r.pos + 1u));
            o.ch('\t');
            o.dec_u(r.mapq & 0xFFu);
            o.ch('\t');
            if (r.n_cigar == 0) o.ch('*');
            else walk_cigar(cigars + r.cigar_off, r.n_cigar, c.cigar_m, [&](uint32_t len, char opc) { o.dec_u(len); o.ch(opc); });
            o.ch('\t');
            put_ref_name(o, c, r.mate_ref);
            o.ch('\t');
            o.dec_u((uint32_t)(r.mate_pos + 1u));
            o.ch('\t');
            o.dec_s(r.tlen);
            o.ch('\t');
            const bool secondary = r.flags & F_SECONDARY, reverse = r.flags & F_REVERSE;
            if (secondary || r.seq_len == 0) o.ch('*');
            else if (reverse) {
                for (uint32_t k = lane; k < r.seq_len; k += 32) o.p[k] = revcomp_char((unsigned char)seq[r.seq_len - 1 - k]);
                o.p += r.seq_len;
            } else o.bytes(seq, r.seq_len);
            o.ch('\t');
            // (Sam::add_record, src/sam.cpp:176-195: the UNMAP branch prints the quality as it is)
            if ((secondary && !(r.flags & F_UNMAP)) || r.qual_len == 0) o.ch('*');
            else if (reverse && !(r.flags & F_UNMAP)) {
                for (uint32_t k = lane; k < r.qual_len; k += 32) o.p[k] = qual[r.qual_len - 1 - k];
                o.p += r.qual_len;
            } else o.bytes(qual, r.qual_len);
            if (!(r.flags & F_UNMAP)) {
                SAM_LIT(o, "\tNM:i:"); o.dec_s(r.edit_distance);
                SAM_LIT(o, "\tAS:i:"); o.dec_s(r.score);
            }
            if (c.show_details) {
                SAM_LIT(o, "\tna:i:"); o.dec_u(r.details[0]);
                SAM_LIT(o, "\tnr:i:"); o.dec_u(r.details[1]);
                SAM_LIT(o, "\tal:i:"); o.dec_u(r.details[2]);
                SAM_LIT(o, "\tga:i:"); o.dec_u(r.details[3]);
                if (r.flags & F_PAIRED) { SAM_LIT(o, "\tmr:i:"); o.dec_u(r.details[4]); }
            }
        }
        o.bytes(c.tail, c.tail_len);
    }
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct rsa_sam {
    int device = 0;
    SamCfg cfg{};
    char* d_names = nullptr;
    long long* d_names_off = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;   // around the length + scan kernels, around the writer
    double kernel_ms = 0;
    Buf arena, out;        // arena: records, text pool, CIGAR pool, line lengths, offsets, scan sums -- ONE allocation (a cudaMalloc
                           // issued while other pipeline workers are enqueueing stalls all of them)
    Buf pin_in, pin_out;   // pinned bounce buffers for callers that pass pageable memory
    std::string err;
};

static std::string g_sam_create_error;

#define SAM_TRY(h, call)                                                                  \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) { (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return RSA_EXT_ERR_CUDA; } \
    } while (0)

static int grow(rsa_sam* h, Buf& b, size_t need) {
    if (need <= b.cap) return RSA_EXT_OK;
    if (b.p) SAM_TRY(h, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t cap = ((need + need / 4) + 1048575) & ~(size_t)1048575;
    SAM_TRY(h, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return RSA_EXT_OK;
}

static int grow_pinned(rsa_sam* h, Buf& b, size_t need) {
    if (need <= b.cap) return RSA_EXT_OK;
    if (b.p) SAM_TRY(h, cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t cap = ((need + need / 2) + 1048575) & ~(size_t)1048575;
    SAM_TRY(h, cudaHostAlloc(&b.p, cap, cudaHostAllocDefault));
    b.cap = cap;
    return RSA_EXT_OK;
}

// page-locked host memory?  (pageable memory reports cudaMemoryTypeUnregistered)
static bool sam_host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" double rsa_sam_kernel_ms(const rsa_sam_t* h) { return h ? h->kernel_ms : 0.0; }

extern "C" const char* rsa_sam_last_error(const rsa_sam_t* h) { return h ? h->err.c_str() : g_sam_create_error.c_str(); }

extern "C" int rsa_sam_create(int32_t device, int32_t n_refs, const char* names_buf, const int64_t* names_off, int32_t cigar_m,
                              const char* read_group, int32_t output_unmapped, int32_t show_details, rsa_sam_t** out) {
    if (!out || n_refs < 0 || (n_refs > 0 && (!names_buf || !names_off))) { g_sam_create_error = "bad argument"; return RSA_EXT_ERR_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        g_sam_create_error = "no usable CUDA device (there is no CPU fallback in this library)";
        return RSA_EXT_ERR_CUDA;
    }
    rsa_sam* h = new rsa_sam();
    h->device = device;
    std::string tail = (read_group && read_group[0]) ? std::string("\tRG:Z:") + read_group + "\n" : std::string("\n");
    if (tail.size() > sizeof(h->cfg.tail)) { delete h; g_sam_create_error = "read group id too long"; return RSA_EXT_ERR_ARG; }
    memcpy(h->cfg.tail, tail.data(), tail.size());
    h->cfg.tail_len = (int)tail.size();
    h->cfg.n_refs = n_refs;
    h->cfg.cigar_m = cigar_m; h->cfg.output_unmapped = output_unmapped; h->cfg.show_details = show_details;
    auto fail = [&](const char* what, cudaError_t e) { g_sam_create_error = std::string(what) + ": " + cudaGetErrorString(e); rsa_sam_destroy(h); return RSA_EXT_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    for (cudaEvent_t* ev : {&h->ev0, &h->ev1, &h->ev2, &h->ev3})
        if ((e = cudaEventCreate(ev)) != cudaSuccess) return fail("cudaEventCreate", e);
    const size_t nb = n_refs ? (size_t)names_off[n_refs] : 0;
    if ((e = cudaMalloc(&h->d_names, nb + 16)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_names_off, sizeof(long long) * (size_t)(n_refs + 1))) != cudaSuccess) return fail("cudaMalloc", e);
    std::vector<long long> off((size_t)n_refs + 1, 0);
    for (int i = 0; i <= n_refs && n_refs > 0; ++i) off[i] = names_off[i];
    if (nb && (e = cudaMemcpy(h->d_names, names_buf, nb, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(h->d_names_off, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    h->cfg.names = h->d_names;
    h->cfg.names_off = h->d_names_off;
    *out = h;
    return RSA_EXT_OK;
}

extern "C" void rsa_sam_destroy(rsa_sam_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) { cudaStreamSynchronize(h->st); cudaStreamDestroy(h->st); }
    for (cudaEvent_t ev : {h->ev0, h->ev1, h->ev2, h->ev3}) if (ev) cudaEventDestroy(ev);
    for (Buf* b : {&h->arena, &h->out}) if (b->p) cudaFree(b->p);
    for (Buf* b : {&h->pin_in, &h->pin_out}) if (b->p) cudaFreeHost(b->p);
    if (h->d_names) cudaFree(h->d_names);
    if (h->d_names_off) cudaFree(h->d_names_off);
    delete h;
}

extern "C" int rsa_sam_format(rsa_sam_t* h, int64_t n, const rsa_sam_record_t* records, const char* text_pool, int64_t text_bytes,
                              const uint32_t* cigar_pool, int64_t n_cigar_ops, char* out, int64_t out_cap, int64_t* out_len,
                              int64_t* line_off) {
    if (!h) return RSA_EXT_ERR_ARG;
    if (n <= 0 || !records || !text_pool || text_bytes < 0 || n_cigar_ops < 0 || (n_cigar_ops > 0 && !cigar_pool) || !out_len) {
        h->err = "bad argument";
        return RSA_EXT_ERR_ARG;
    }
    for (int64_t i = 0; i < n; ++i) {   // the descriptors index caller memory: check them before the device follows them
        const rsa_sam_record_t& r = records[i];
        if (r.kind > RSA_SAM_UNMAPPED_MATE || r.name_off + r.name_len > (uint64_t)text_bytes || r.seq_off + r.seq_len > (uint64_t)text_bytes ||
            r.qual_off + r.qual_len > (uint64_t)text_bytes || (uint64_t)r.cigar_off + r.n_cigar > (uint64_t)n_cigar_ops) {
            h->err = "record " + std::to_string(i) + " points outside the pools";
            return RSA_EXT_ERR_ARG;
        }
    }
    SAM_TRY(h, cudaSetDevice(h->device));
    const int nb = (int)((n + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems));
    int rc;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_text = al(sizeof(rsa_sam_record_t) * (size_t)n), o_cig = o_text + al((size_t)text_bytes + 16),
                 o_len = o_cig + al(sizeof(uint32_t) * (size_t)n_cigar_ops + 16), o_off = o_len + al(sizeof(unsigned long long) * (size_t)n),
                 o_sums = o_off + al(sizeof(unsigned long long) * (size_t)(n + 1)), arena_need = o_sums + al(sizeof(unsigned long long) * (size_t)(nb + 1));
    if ((rc = grow(h, h->arena, arena_need))) return rc;
    // the text buffer is sized before the lengths are known (so that a first call allocates everything at once): the caller's
    // capacity, capped by a generous estimate; grown below in the rare case the estimate was short
    const size_t est_out = std::min<size_t>((size_t)(out_cap > 0 ? out_cap : 0), (size_t)text_bytes + (size_t)n * 256 + 12 * (size_t)n_cigar_ops + 64);
    if ((rc = grow(h, h->out, est_out))) return rc;
    char* const a_base = (char*)h->arena.p;
    void* const d_recs = a_base; void* const d_text = a_base + o_text; void* const d_cig = a_base + o_cig;
    cudaStream_t st = h->st;
    // Pageable caller memory (the pipeline's per-chunk collector, integration/sam_glue.cpp) goes through pinned bounce
    // buffers: pageable cudaMemcpyAsync calls are synchronous driver-staged copies, and sixteen workers formatting their
    // chunks through them slowed every GPU call of the process (see hamming_core in engine.cu for the measurement).
    // Pinned callers are copied from / to directly.
    const size_t recb = sizeof(rsa_sam_record_t) * (size_t)n, cigb = sizeof(uint32_t) * (size_t)n_cigar_ops;
    const void *src_rec = records, *src_text = text_pool, *src_cig = cigar_pool;
    if (!(sam_host_is_pinned(records) && sam_host_is_pinned(text_pool) && (!n_cigar_ops || sam_host_is_pinned(cigar_pool)))) {
        const size_t a_text = (recb + 15) & ~(size_t)15, a_cig = a_text + (((size_t)text_bytes + 15) & ~(size_t)15);
        if ((rc = grow_pinned(h, h->pin_in, a_cig + cigb + 16))) return rc;
        char* pin = (char*)h->pin_in.p;
        memcpy(pin, records, recb);
        if (text_bytes) memcpy(pin + a_text, text_pool, (size_t)text_bytes);
        if (n_cigar_ops) memcpy(pin + a_cig, cigar_pool, cigb);
        src_rec = pin; src_text = pin + a_text; src_cig = pin + a_cig;
    }
    SAM_TRY(h, cudaMemcpyAsync(d_recs, src_rec, recb, cudaMemcpyHostToDevice, st));
    if (text_bytes) SAM_TRY(h, cudaMemcpyAsync(d_text, src_text, (size_t)text_bytes, cudaMemcpyHostToDevice, st));
    if (n_cigar_ops) SAM_TRY(h, cudaMemcpyAsync(d_cig, src_cig, cigb, cudaMemcpyHostToDevice, st));
    const rsa_sam_record_t* d_rec = (const rsa_sam_record_t*)d_recs;
    unsigned long long* d_len = (unsigned long long*)(a_base + o_len);
    unsigned long long* d_off = (unsigned long long*)(a_base + o_off);
    unsigned long long* d_sums = (unsigned long long*)(a_base + o_sums);
    h->kernel_ms = 0;
    SAM_TRY(h, cudaEventRecord(h->ev0, st));
    sam_length_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->cfg, d_rec, (long long)n, (const char*)d_text, (const uint32_t*)d_cig, d_len);
    scan_block_sums<<<nb, kScanThreads, 0, st>>>(d_len, (long long)n, d_sums);
    scan_of_sums<<<1, 1, 0, st>>>(d_sums, nb, d_sums + nb);
    scan_apply<<<nb, kScanThreads, 0, st>>>(d_len, (long long)n, d_sums, d_off);
    SAM_TRY(h, cudaGetLastError());
    SAM_TRY(h, cudaEventRecord(h->ev1, st));
    // (holds the total; sized for the text right away when the destination is pageable: one allocation on a first call)
    const bool out_is_pinned = out && sam_host_is_pinned(out);
    if ((rc = grow_pinned(h, h->pin_out, out_is_pinned ? 64 : est_out + 64))) return rc;
    unsigned long long total = 0;
    SAM_TRY(h, cudaMemcpyAsync(h->pin_out.p, d_sums + nb, sizeof total, cudaMemcpyDeviceToHost, st));
    SAM_TRY(h, cudaStreamSynchronize(st));
    total = *(const unsigned long long*)h->pin_out.p;
    *out_len = (int64_t)total;
    if (line_off) {
        static_assert(sizeof(int64_t) == sizeof(unsigned long long), "offsets are copied as they are");
        SAM_TRY(h, cudaMemcpyAsync(line_off, d_off, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    }
    if ((int64_t)total > out_cap || (!out && total)) {
        SAM_TRY(h, cudaStreamSynchronize(st));
        h->err = "output buffer too small: " + std::to_string(total) + " bytes needed";
        return RSA_EXT_ERR_ARG;
    }
    if (total) {
        if ((rc = grow(h, h->out, (size_t)total))) return rc;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const long long want = (n + kSamWarps - 1) / kSamWarps;
        const int blocks = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
        SAM_TRY(h, cudaEventRecord(h->ev2, st));
        sam_write_kernel<<<blocks, 32 * kSamWarps, 0, st>>>(h->cfg, d_rec, (long long)n, (const char*)d_text, (const uint32_t*)d_cig, d_off, (char*)h->out.p);
        SAM_TRY(h, cudaGetLastError());
        SAM_TRY(h, cudaEventRecord(h->ev3, st));
        if (out_is_pinned) {
            SAM_TRY(h, cudaMemcpyAsync(out, h->out.p, (size_t)total, cudaMemcpyDeviceToHost, st));
            SAM_TRY(h, cudaStreamSynchronize(st));
        } else {
            if ((rc = grow_pinned(h, h->pin_out, (size_t)total))) return rc;
            SAM_TRY(h, cudaMemcpyAsync(h->pin_out.p, h->out.p, (size_t)total, cudaMemcpyDeviceToHost, st));
            SAM_TRY(h, cudaStreamSynchronize(st));
            memcpy(out, h->pin_out.p, (size_t)total);
        }
    }
    SAM_TRY(h, cudaStreamSynchronize(st));
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, h->ev0, h->ev1);                 // length + scan kernels
    if (total) cudaEventElapsedTime(&b, h->ev2, h->ev3);      // writer
    h->kernel_ms = (double)a + (double)b;
    return RSA_EXT_OK;
}

// ---- host helpers: the reference's flag / mate / template-length logic ----------------------------------------------
static void fill_read(rsa_sam_record_t* o, const rsa_sam_read_t* r) {
    o->name_off = r->name_off; o->name_len = r->name_len;
    o->seq_off = r->seq_off; o->seq_len = r->seq_len;
    o->qual_off = r->qual_off; o->qual_len = r->qual_len;
}

extern "C" void rsa_sam_unmapped(const rsa_sam_read_t* read, uint32_t flags, rsa_sam_record_t* out) {
    memset(out, 0, sizeof *out);
    out->kind = RSA_SAM_UNMAPPED;
    out->flags = flags;
    out->ref_id = RSA_SAM_REF_NONE; out->mate_ref = RSA_SAM_REF_NONE;
    fill_read(out, read);
}

static void aligned_record(rsa_sam_record_t* o, const rsa_sam_read_t* read, uint32_t flags, int32_t ref_id, uint32_t pos, uint32_t mapq,
                           const rsa_sam_alignment_t* a, int32_t mate_ref, uint32_t mate_pos, int32_t tlen, const uint32_t details[5]) {
    memset(o, 0, sizeof *o);
    o->kind = RSA_SAM_ALIGNED;
    o->flags = flags; o->ref_id = ref_id; o->pos = pos; o->mapq = mapq & 0xFFu;
    o->mate_ref = mate_ref; o->mate_pos = mate_pos; o->tlen = tlen;
    o->edit_distance = a->edit_distance; o->score = a->score;
    o->cigar_off = a->cigar_off; o->n_cigar = a->n_cigar;
    if (details) memcpy(o->details, details, sizeof o->details);
    fill_read(o, read);
}

// Sam::add (src/sam.cpp:117-139)
extern "C" void rsa_sam_single(const rsa_sam_alignment_t* a, const rsa_sam_read_t* read, uint32_t mapq, int32_t is_primary,
                               const uint32_t details[5], rsa_sam_record_t* out) {
    uint32_t flags = 0;
    if (!a->is_unaligned && a->is_rc) flags |= F_REVERSE;
    if (!is_primary) { flags |= F_SECONDARY; mapq = 255; }
    aligned_record(out, read, flags, a->ref_id, (uint32_t)a->ref_start, mapq, a, RSA_SAM_REF_NONE, (uint32_t)-1, 0, details);
}

// Sam::add_pair (src/sam.cpp:208-318)
extern "C" void rsa_sam_pair(const rsa_sam_alignment_t* a1, const rsa_sam_alignment_t* a2, const rsa_sam_read_t* r1,
                             const rsa_sam_read_t* r2, uint32_t mapq1, uint32_t mapq2, int32_t is_proper, int32_t is_primary,
                             const uint32_t details1[5], const uint32_t details2[5], rsa_sam_record_t out[2]) {
    uint32_t f1 = F_PAIRED | F_READ1, f2 = F_PAIRED | F_READ2;
    if (!is_primary) { f1 |= F_SECONDARY; f2 |= F_SECONDARY; }
    int template_len1 = 0;
    const bool both_aligned = !a1->is_unaligned && !a2->is_unaligned;
    if (both_aligned && a1->ref_id == a2->ref_id) {
        const int dist = a2->ref_start - a1->ref_start;
        template_len1 = dist > 0 ? dist + a2->length : dist - a1->length;
    }
    if (is_proper) { f1 |= F_PROPER; f2 |= F_PROPER; }
    int32_t name1, name2;
    int pos1 = a1->ref_start, pos2 = a2->ref_start;
    if (a1->is_unaligned) { f1 |= F_UNMAP; f2 |= F_MUNMAP; pos1 = -1; name1 = RSA_SAM_REF_NONE; }
    else { if (a1->is_rc) { f1 |= F_REVERSE; f2 |= F_MREVERSE; } name1 = a1->ref_id; }
    if (a2->is_unaligned) { f2 |= F_UNMAP; f1 |= F_MUNMAP; pos2 = -1; name2 = RSA_SAM_REF_NONE; }
    else { if (a2->is_rc) { f1 |= F_MREVERSE; f2 |= F_REVERSE; } name2 = a2->ref_id; }
    int32_t mate_name1 = name1, mate_name2 = name2;
    if ((both_aligned && a1->ref_id == a2->ref_id) || (a1->is_unaligned != a2->is_unaligned)) { mate_name1 = RSA_SAM_REF_SAME; mate_name2 = RSA_SAM_REF_SAME; }
    if (a1->is_unaligned != a2->is_unaligned) { if (a1->is_unaligned) pos1 = pos2; else pos2 = pos1; }
    auto unmapped_mate = [](rsa_sam_record_t* o, const rsa_sam_read_t* r, uint32_t flags, int32_t mate_reference, uint32_t mate_pos) {
        memset(o, 0, sizeof *o);
        o->kind = RSA_SAM_UNMAPPED_MATE;
        o->flags = flags; o->ref_id = mate_reference; o->mate_ref = RSA_SAM_REF_SAME; o->mate_pos = mate_pos;
        fill_read(o, r);
    };
    if (a1->is_unaligned) unmapped_mate(&out[0], r1, f1, name2, (uint32_t)pos2);
    else aligned_record(&out[0], r1, f1, name1, (uint32_t)a1->ref_start, mapq1, a1, mate_name2, (uint32_t)pos2, template_len1, details1);
    if (a2->is_unaligned) unmapped_mate(&out[1], r2, f2, name1, (uint32_t)pos1);
    else aligned_record(&out[1], r2, f2, name2, (uint32_t)a2->ref_start, mapq2, a2, mate_name1, (uint32_t)pos1, -template_len1, details2);
}
