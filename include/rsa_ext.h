/*
 * rsa_ext.h -- C ABI of the B200-native extension engine (librsa_ext.so).
 *
 * Drop-in boundary for the one GPU hot path of RabbitSAlign: the batched affine-gap local
 * Smith-Waterman + traceback + CIGAR step behind
 *
 *     void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res>&, std::vector<std::string>&,
 *                           std::vector<std::string>&, int match, int mismatch, int gap_open,
 *                           int gap_extend);                       (reference src/gasal2_ssw.h:46-47)
 *
 * Each entry point names the reference interface it replaces (paths relative to /root/reference).
 * Plain pointers and sizes only; no C++ or torch types.  All functions return RSA_EXT_OK (0) or a
 * negative status; rsa_ext_last_error() gives the text.  A handle is owned by one host thread at a
 * time (the reference keeps one GASAL stream + storage per worker `thread_id`,
 * src/gasal2_ssw.cpp:29,92-102); different handles are independent and may live on different GPUs.
 * A handle that has seen a batch of more than 32768 pairs owns one helper thread (it plans the chunks of
 * large batches ahead of the caller's thread; idle otherwise, joined by rsa_ext_destroy).  Handle creation
 * and each handle's first submit are serialised process-wide (driver-lock contention).
 * RSA_EXT_TRACE=1 in the environment prints host-side timing laps to stderr.
 *
 * There is NO CPU fallback inside this library: without a usable CUDA device rsa_ext_create fails.
 */
#ifndef RSA_EXT_H
#define RSA_EXT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSA_EXT_OK 0
#define RSA_EXT_ERR_ARG -1      /* bad argument (NULL, n <= 0, offsets not monotone, ...) */
#define RSA_EXT_ERR_CUDA -2     /* a CUDA call failed; the reference exit()s here (GASAL2/src/gasal.h:15-22) */
#define RSA_EXT_ERR_QUERY_LEN -3 /* a query is longer than max_query_len; the reference prints and exit(0)s
                                    (src/gasal2_ssw.cpp:84-87) */
#define RSA_EXT_ERR_STATE -4    /* wait without submit, submit while a batch is pending, ... */

#define RSA_EXT_RLE_INLINE 40   /* RLE bytes carried inside each result record */

typedef struct rsa_ext rsa_ext_t;

/* Replaces the compile-time knobs of src/gasal2_ssw.h:15-25 (MAX_QUERY_LEN 500, MAX_TARGET_LEN 2000,
 * STREAM_BATCH_SIZE 512 -- batches here may have any n >= 1) and the `Parameters` fake argv of
 * src/gasal2_ssw.cpp:37-44 plus gasal_copy_subst_scores (GASAL2/src/gasal_align.cu:329-339).
 * Scores are the strobealign values (src/cmdline.hpp:46-50: A=2 B=8 O=12 E=1); like
 * src/gasal2_ssw.cpp:52-56 the engine opens a gap with gap_open-1+gap_extend and extends with
 * gap_extend.  Scores are per handle, not global __constant__ state. */
typedef struct {
    int32_t device;         /* CUDA ordinal; the reference never calls cudaSetDevice (gasal2_ssw.cpp:34) */
    int32_t max_query_len;  /* 0 -> 500 */
    int32_t max_target_len; /* 0 -> 2000; longer windows get a failed record (score 0, n_ops 0), the
                               caller never consumes those (src/aligner.cpp:18-24) */
    int32_t match, mismatch, gap_open, gap_extend; /* all 0 -> 2, 8, 12, 1 */
    int32_t flags;          /* RSA_EXT_FLAG_* */
    int64_t scratch_bytes;  /* direction-bit scratch per handle; 0 -> default */
} rsa_ext_config_t;

#define RSA_EXT_FLAG_EXACT_ONLY 1 /* route every pair through the exact int32 kernel (testing) */
#define RSA_EXT_FLAG_ASCII_WINDOWS 8 /* window form: stage the windows from the ASCII copy of the resident reference
                                        instead of its packed planes (A/B tests; results are identical) */
#define RSA_EXT_FLAG_HOST_PLAN 4  /* plan every chunk on the host (the round-1 planner); default: batches of >= 4096 pairs
                                     are planned by device kernels from the raw offset arrays (same records) */
#define RSA_EXT_FLAG_SERIALIZE 2  /* one stream, no kernel overlap: per-kernel CUDA-event times (dp_ms, tb_ms) are then
                                     clean kernel durations (roofline measurements); slower than the default */

/* One result per pair: the integer fields of `struct gasal_tmp_res` (src/gasal2_ssw.h:31-38; starts and
 * ends 0-based inclusive, starts may be -1) plus the traceback's run-length bytes exactly as
 * GASAL2/src/kernels/get_tb.h:87-117 emits them: (count<<2)|op, count <= 63, op 0=M 1=X 2=D 3=I, in
 * END-TO-START order.  n_ops > RSA_EXT_RLE_INLINE: rle[] holds the first RSA_EXT_RLE_INLINE bytes and the full byte
 * string is fetched with rsa_ext_rle_overflow() (records are byte-deterministic: no internal offsets inside).  status: 0 ok; 1 window longer than max_target_len (not aligned; the caller never consumes
 * those, src/aligner.cpp:18-24); 3 empty query or window (the reference reads an unwritten tile there).  All other
 * records are bit-exact with the reference, including the ones its gasal_fail gate rejects.  (4 and 5 never reach the
 * caller: 4 = re-run internally by rsa_ext_wait; 5 = no kernel produced the pair, rsa_ext_wait returns
 * RSA_EXT_ERR_CUDA instead of handing such records back.) */
typedef struct {
    int32_t score;
    int32_t query_start;
    int32_t query_end;
    int32_t ref_start;
    int32_t ref_end;
    int16_t n_ops;
    int16_t status;
    uint8_t rle[RSA_EXT_RLE_INLINE];
} rsa_ext_result_t; /* 64 bytes */

/* gasal_init_gpu_storage_v + gasal_init_streams (GASAL2/src/ctors.cpp:26-128; call site
 * src/gasal2_ssw.cpp:92-102).  Storage grows on demand and is freed by rsa_ext_destroy (the reference
 * never frees: gasal2_ssw.cpp:252-255). */
int rsa_ext_create(const rsa_ext_config_t *cfg, rsa_ext_t **out);
void rsa_ext_destroy(rsa_ext_t *h);
const char *rsa_ext_last_error(const rsa_ext_t *h); /* h may be NULL: error of the last failed create */

/* gasal_host_batch_fill x 2n + gasal_op_fill + gasal_aln_async (src/gasal2_ssw.cpp:114-153,
 * GASAL2/src/host_batch.cpp:79-153, GASAL2/src/gasal_align.cu:29-307).  Asynchronous: returns once the
 * copies and kernels are enqueued.  Pair i is qbuf[qoff[i] .. qoff[i+1]) against tbuf[toff[i] ..
 * toff[i+1]) (raw ASCII, any case; batches of >= 16384 pairs are validated chunk by chunk while they run, so an
 * invalid pair (query longer than max_query_len, offsets not monotone) may be reported by rsa_ext_wait instead; bases are compared on `byte & 0xF` with 0xE as the zero-scoring
 * wildcard, GASAL2/src/kernels/pack_rc_seqs.h:13-53, gasal_kernels.h:48-51).  Buffers must stay valid
 * until rsa_ext_wait returns; pinned buffers are copied without staging. */
int rsa_ext_submit(rsa_ext_t *h, int64_t n, const char *qbuf, const int64_t *qoff, const char *tbuf,
                   const int64_t *toff, rsa_ext_result_t *results);

/* Same, from n separate strings (the shape of std::vector<std::string>). */
int rsa_ext_submit_ptrs(rsa_ext_t *h, int64_t n, const char *const *q, const int32_t *qlen,
                        const char *const *t, const int32_t *tlen, rsa_ext_result_t *results);

/* gasal_is_aln_async_done (GASAL2/src/gasal_align.cu:310-326): 0 = finished, 1 = still running.  Usable in the
 * reference's `while (poll) usleep(100)` loop (src/gasal2_ssw.cpp:179): poll itself retires finished chunks and
 * enqueues the next ones of a multi-chunk batch, without blocking.  After it returned 0, rsa_ext_wait no longer waits
 * for the GPU; it must still be called (it finalises the batch and reports any error). */
int rsa_ext_poll(rsa_ext_t *h);

/* The poll/usleep loop and result unpacking of src/gasal2_ssw.cpp:179-249: blocks until `results`
 * (given to submit) is filled. */
int rsa_ext_wait(rsa_ext_t *h);

/* ---- SURVEY 8(f) "next" row 1, first half: windows named by offset and length in a resident reference ------------
 *
 * The reference builds each window as a std::string (references.sequences[id].substr(...), src/pc.cpp:214-242) and
 * GASAL copies it to the GPU.  rsa_ext_set_reference uploads the (concatenated) reference once -- at most 2^32-1
 * bytes; `seq` must stay valid while windows are submitted (it is read again only when a pair is re-run exact-only).
 * rsa_ext_submit_ref_windows then takes targets as win_off[i], win_len[i] into that buffer; queries as in
 * rsa_ext_submit.  Results, rsa_ext_wait/poll and every status are exactly those of rsa_ext_submit on the same
 * bytes (tests/test_gpu_parity.py::test_reference_windows_equal_explicit_windows). */
int rsa_ext_set_reference(rsa_ext_t *h, const char *seq, int64_t len);
/* Let `h` use the reference `donor` uploaded (same device; the copy in HBM is shared and freed with its last user):
 * the pipeline's workers each own a handle but need the reference once per GPU (north_star: "index and reference
 * replicated per GPU"). */
int rsa_ext_share_reference(rsa_ext_t *h, const rsa_ext_t *donor);
/* The packed planes the upload built beside the ASCII copy and the packed DP kernel stages windows from (north_star:
 * "2-bit-packed reads and reference windows are staged into shared memory with vectorised, coalesced HBM loads"; the
 * reference packs to 4 bits on the device, GASAL2/src/kernels/pack_rc_seqs.h:13-53): per 64-base unit 16 bytes of
 * 2-bit codes (A C G T = 0..3, first base in the low bits) and 8 bytes of "not ACGT" bits (then code 0 = N, 1 = any
 * other symbol).  Copies `units` whole units out (diagnostics, tests). */
int rsa_ext_packed_reference(rsa_ext_t *h, uint32_t *codes, uint32_t *flags, int64_t units);
int rsa_ext_submit_ref_windows(rsa_ext_t *h, int64_t n, const char *qbuf, const int64_t *qoff,
                               const int64_t *win_off, const int32_t *win_len, rsa_ext_result_t *results);

/* Pre-allocate what batches of up to n pairs of (qlen x tlen) need (device buffers of the first chunk slot, pinned
 * staging), so the first real batches allocate nothing.  Optional; the reference allocates its GASAL storage for
 * STREAM_BATCH_SIZE x MAX lengths up front (gasal_init_streams, src/gasal2_ssw.cpp:92-102). */
int rsa_ext_reserve(rsa_ext_t *h, int64_t n, int32_t qlen, int32_t tlen);

/* Full RLE byte string of pair i of the last waited batch when n_ops > RSA_EXT_RLE_INLINE.
 * Returns the number of bytes written (<= cap) or a negative status. */
int rsa_ext_rle_overflow(rsa_ext_t *h, int64_t i, uint8_t *out, int32_t cap);

/* Host CIGAR text of src/gasal2_ssw.cpp:184-243 from RLE bytes: read last-to-first, merge equal
 * neighbours, "<count><M|X|D|I>".  Returns text length (no NUL counted) or -1 if cap is too small. */
int rsa_ext_rle_to_text(const uint8_t *rle, int32_t n_ops, char *out, int32_t cap);

/* ---- SURVEY 8(f) "next" rows 1 and 3: the host post-processing of every GPU record, on the device -----------
 *
 * What src/pc.cpp:735-744 does per pair on the host -- gasal_fail (src/pc.cpp:466-478), then Aligner::align_gpu
 * (src/aligner.cpp:13-112: CIGAR text -> ops, soft clips, M -> '=', NM = X+I+D via ext/ssw/ssw_cpp.cpp:54-90,
 * 212-247,396-429, and the greedy end-bonus extension to both read ends) -- computed by a kernel behind the
 * traceback.  The record mirrors `AlignmentInfo` (src/aligner.hpp:20-30); CIGAR ops are BAM-style
 * (len << 4 | op, op codes of src/cigar.hpp:11-21: 1 I, 2 D, 4 S, 7 =, 8 X).
 * status: 0 accepted (fields as align_gpu returns them); 1 gasal_fail -> the caller runs Aligner::align on the CPU
 * like the reference; 2 window > max_target_len -> the sentinel of src/aligner.cpp:18-24; 3 CIGAR longer than
 * RSA_EXT_CIGAR_INLINE ops -> use the rsa_ext_result_t record and the host path for this pair. */
#define RSA_EXT_CIGAR_INLINE 25
typedef struct {
    int32_t sw_score;
    int32_t edit_distance;
    int32_t ref_start, ref_end;     /* end exclusive */
    int32_t query_start, query_end; /* end exclusive */
    int16_t n_cigar;
    int16_t status;
    uint32_t cigar[RSA_EXT_CIGAR_INLINE];
} rsa_ext_alninfo_t; /* 128 bytes */

/* Ask the following rsa_ext_submit/_ptrs calls to also fill out[0..n) (valid after rsa_ext_wait); out == NULL
 * switches it off.  The pointer must cover the largest batch submitted while it is set.
 * end_bonus is strobealign's -L (src/cmdline.hpp:50, default 10). */
int rsa_ext_request_alninfo(rsa_ext_t *h, rsa_ext_alninfo_t *out, int32_t end_bonus);

/* ---- SURVEY 8(f) "next" row 3, first half: the Hamming shortcut of the seed extension, on the device -------
 *
 * What extend_seed_part (src/aln.cpp:374-431) does on the host for every candidate site whose projected window has the
 * read's length, before deciding to run Smith-Waterman: hamming_distance (src/aligner.hpp:54-67), the test
 * (float) distance / |query| < 0.05 (src/aln.cpp:395) and, if it holds, hamming_align (src/aligner.cpp:254-302, with
 * highest_scoring_segment :219-252).  Blocking call, one warp per pair.
 * hamming[i]: the distance, -1 for windows of another length than the read (and for reads longer than 512 bases, which
 * the engine does not extend either: status 1).
 * out[i].status: 0 the shortcut applies and the record is hamming_align's AlignmentInfo (ref_start/ref_end relative to
 * the window); 1 the pair needs the gapped path (submit it); 3 more than RSA_EXT_CIGAR_INLINE runs (host path).
 * Scores are the handle's; end_bonus is strobealign's -L.  Host arrays may be pinned or pageable: pinned ones are copied from / to
 * directly, pageable ones go through pinned bounce buffers of the handle (sized ahead by rsa_ext_reserve). */
int rsa_ext_hamming_align(rsa_ext_t *h, int64_t n, const char *qbuf, const int64_t *qoff, const char *tbuf,
                          const int64_t *toff, int32_t end_bonus, int32_t *hamming, rsa_ext_alninfo_t *out);
/* The same with window i = [win_off[i], win_off[i] + |query i|) of the resident reference (rsa_ext_set_reference). */
int rsa_ext_hamming_ref_windows(rsa_ext_t *h, int64_t n, const char *qbuf, const int64_t *qoff, const int64_t *win_off,
                                int32_t end_bonus, int32_t *hamming, rsa_ext_alninfo_t *out);

/* ---- device-resident legs (bench.py `value`, roofline): inputs already in HBM ---------------- */

/* Upload + plan a batch once; afterwards rsa_ext_run_resident() re-runs only the GPU kernels on the
 * handle's stream (no host<->device traffic) and rsa_ext_fetch_resident() copies the records out. */
int rsa_ext_stage_resident(rsa_ext_t *h, int64_t n, const char *qbuf, const int64_t *qoff,
                           const char *tbuf, const int64_t *toff);
int rsa_ext_run_resident(rsa_ext_t *h);
int rsa_ext_fetch_resident(rsa_ext_t *h, rsa_ext_result_t *results);

/* The handle's CUDA stream (cudaStream_t as void*) so callers can bracket launches with events. */
void *rsa_ext_stream(rsa_ext_t *h);

/* Counters of the last submit/run: kernels launched, pairs taken by each kernel family, DP cells. */
typedef struct {
    int64_t kernel_launches;
    int64_t pairs_fast;   /* packed s16x2 DPX kernel */
    int64_t pairs_exact;  /* int32 exact kernel (odd nibbles, ties, tiny/huge shapes) */
    int64_t pairs_failed; /* status != 0 */
    int64_t cells;        /* sum |q|*|t| */
    int64_t h2d_bytes, d2h_bytes;
    double dp_ms;         /* device time of the DP kernels of the last run_resident (CUDA events) */
    double tb_ms;         /* device time of the traceback kernels */
    int64_t pairs_redo;   /* pairs the packed kernel handed to the exact kernel at run time (ties on the maximum,
                             symbols outside ACGTN); resident runs: last chunk only */
    double host_plan_ms;  /* host time spent planning chunks during the last submit/wait */
} rsa_ext_stats_t;
int rsa_ext_get_stats(const rsa_ext_t *h, rsa_ext_stats_t *out);

int rsa_ext_version(void);

/* Test hook, no CUDA call: plans the first chunk of a batch on the host and reports how it would be routed
 * (out[0..7]: pairs, packed-kernel pairs, exact-kernel pairs, failed, groups, scratch bytes, column classes,
 * in/out: timing repetitions -> mean ns per plan). */
int rsa_ext_plan_debug(int64_t n, const int64_t *qoff, const int64_t *toff, int64_t scratch_cap, int exact_only,
                       int64_t *out);

/* Same for the host pass of the device planner (batches of >= 16384 pairs: the host only counts pairs per query length
 * and cuts the chunk; sorting, pairing and tile offsets are computed by kernels).  out[4] = group slots, out[5] = the
 * scratch bound. */
int rsa_ext_scan_debug(int64_t n, const int64_t *qoff, const int64_t *toff, int64_t scratch_cap, int64_t *out);

/* Number of usable CUDA devices (0 without a driver/GPU): lets a host pipeline spread its workers over the
 * GPUs of one box (the reference is single-device, src/gasal2_ssw.cpp:34). */
int rsa_ext_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RSA_EXT_H */
