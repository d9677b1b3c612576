/*
 * rsa_seed.h -- C ABI of the B200-native seeding path (part of librsa_ext.so): SURVEY.md 8(f) rank 2, the second kernel
 * family north_star names ("randstrobe seeding and NAM-merge ... ported ... if profiling shows it becomes the
 * bottleneck once extension is on the GPU" -- it is: 70-80 % of the pipeline's per-read host time).
 *
 * What it replaces, per read, in the reference's align_SE_read_part / align_PE_read_part (src/aln.cpp:1927-1958,
 * 2380-2400; paths relative to /root/reference):
 *
 *     query_randstrobes = randstrobes_query(seq, index_parameters)        src/randstrobes.cpp:207 (+ :57-127 syncmers,
 *                                                                         :151-176 randstrobe linking, src/hash.hpp xxh64)
 *     [fraction, nams]  = find_nams(query_randstrobes, index)             src/nam.cpp:771-922 (index lookup
 *                                                                         src/index.hpp:60-97, hit merge src/nam.cpp:368-510)
 *     if (rescue_level > 1 && (nams.empty() || fraction < 0.7))
 *         nams = find_nams_rescue(query_randstrobes, index, rescue_cutoff)   src/nam.cpp:955-1012 (+ :117-366 merge)
 *
 * for a whole batch of reads in one call, against the reference's own index arrays resident in HBM (replicated per GPU,
 * north_star).  Results are bit-exact: every field of every NAM (src/nam.hpp:11-38), the nonrepetitive fraction and the
 * rescue decision.  One thing is left to the caller: the reference appends the NAMs of one strand reference-id by
 * reference-id in the ITERATION ORDER OF ITS HASH MAP (robin_hood::unordered_map, src/nam.cpp:775), a property of that
 * container.  This library returns the groups of a strand in first-touch order (the order the reference inserts the
 * keys) and tags every NAM with its group; the binding re-orders groups with the reference's own container when a strand
 * touches more than one reference sequence (integration/seed_glue.hpp shows how; oracle/seed_ref_shim.cpp does it for
 * the tests).  nam_id is the NAM's index after that re-ordering.
 *
 * Plain pointers and sizes only.  Functions return RSA_SEED_OK (0) or a negative status; rsa_seed_last_error() has the
 * text.  There is NO CPU fallback in this library: a read whose intermediate lists exceed even the large scratch tier
 * comes back flagged RSA_SEED_READ_FAILED and the caller decides (the pipeline binding runs the reference's host code
 * for that one read).
 */
#ifndef RSA_SEED_H
#define RSA_SEED_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSA_SEED_OK 0
#define RSA_SEED_ERR_ARG -1
#define RSA_SEED_ERR_CUDA -2
#define RSA_SEED_ERR_STATE -4

typedef struct rsa_seed_index rsa_seed_index_t; /* the index resident on one GPU, shared by that GPU's workers */
typedef struct rsa_seed rsa_seed_t;             /* one host worker's stream + buffers */

/* IndexParameters (src/indexparameters.hpp:11-93) + what StrobemerIndex / MappingParameters add (src/index.hpp:49-185,
 * src/aln.hpp:58-75, src/main.cpp:415). */
typedef struct {
    int32_t device;
    int32_t k, s, t_syncmer;        /* SyncmerParameters */
    int32_t w_min, w_max, max_dist; /* RandstrobeParameters */
    uint64_t q;
    int32_t bits;                   /* StrobemerIndex::bits: top bits of the hash that select a bucket */
    uint32_t filter_cutoff;         /* StrobemerIndex::filter_cutoff */
    int32_t rescue_level;           /* MappingParameters::rescue_level (default 2) */
    uint32_t rescue_cutoff;         /* MappingParameters::rescue_cutoff */
} rsa_seed_config_t;

/* Upload the index: `randstrobes` = StrobemerIndex::randstrobes (n entries of RefRandstrobe: u64 hash, u32 position,
 * u32 packed = ref_index << 8 | strobe2_offset; src/randstrobes.hpp:21-50), `starts` = randstrobe_start_indices
 * ((1 << bits) + 1 entries of u64).  The arrays are copied; the host copies may be freed afterwards. */
int rsa_seed_index_upload(const rsa_seed_config_t *cfg, const void *randstrobes, int64_t n, const uint64_t *starts,
                          int64_t n_starts, rsa_seed_index_t **out);
void rsa_seed_index_free(rsa_seed_index_t *ix);

int rsa_seed_create(rsa_seed_index_t *ix, rsa_seed_t **out);
void rsa_seed_destroy(rsa_seed_t *h);
const char *rsa_seed_last_error(const rsa_seed_t *h); /* h may be NULL: error of the last failed upload/create */

/* One NAM: the fields of `struct Nam` (src/nam.hpp:11-38) minus nam_id, plus the group tag. */
typedef struct {
    int32_t query_start, query_end, query_prev_hit_startpos;
    int32_t ref_start, ref_end, ref_prev_hit_startpos;
    int32_t n_hits, ref_id;
    float score;
    uint32_t flags; /* bit 0: is_rc; bits 8..31: group = index of this NAM's reference id in the strand's first-touch order */
} rsa_seed_nam_t; /* 40 bytes */

#define RSA_SEED_READ_RESCUED 1u /* find_nams_rescue was run (Details::nam_rescue) */
#define RSA_SEED_READ_FAILED 2u  /* intermediate lists exceeded the large scratch tier: not seeded */

typedef struct {
    uint32_t nam_off; /* first NAM of this read in the NAM array */
    int32_t n_nams;
    float nonrepetitive_fraction; /* find_nams' first result (of the non-rescue pass, as the reference keeps it) */
    uint32_t flags;               /* RSA_SEED_READ_* */
} rsa_seed_read_t; /* 16 bytes */

/* Seed a batch: read i = reads[roff[i] .. roff[i+1]) (ASCII; anything outside ACGT/acgt/Uu restarts the syncmer window
 * like the reference's seq_nt4_table, src/randstrobes.cpp:13-30).  Blocking.  *per_read (n_reads entries) and *nams
 * (*n_nams entries; a read's NAMs are consecutive: strand 0 groups in first-touch order, then strand 1) point into
 * pinned buffers owned by the handle, valid until the next call on it. */
int rsa_seed_find_nams(rsa_seed_t *h, int64_t n_reads, const char *reads, const int64_t *roff,
                       const rsa_seed_read_t **per_read, const rsa_seed_nam_t **nams, int64_t *n_nams);

/* Counters of the last call. */
typedef struct {
    int64_t reads, nams, reads_rescued, reads_retried /* needed the large scratch tier */, reads_failed;
    int64_t h2d_bytes, d2h_bytes;
    double kernel_ms; /* device time of the seeding kernels (CUDA events) */
    int64_t kernel_launches;
    double kernel_ms_large; /* part of kernel_ms spent in the large scratch tier */
} rsa_seed_stats_t;
int rsa_seed_get_stats(const rsa_seed_t *h, rsa_seed_stats_t *out);

/* Device-resident leg for benchmarks: stage reads once, re-run the kernels only. */
int rsa_seed_stage(rsa_seed_t *h, int64_t n_reads, const char *reads, const int64_t *roff);
int rsa_seed_run_staged(rsa_seed_t *h);
void *rsa_seed_stream(rsa_seed_t *h);

#ifdef __cplusplus
}
#endif
#endif /* RSA_SEED_H */
