/*
 * rsa_sam.h -- C ABI of the device-side SAM record formatter (part of librsa_ext.so).
 *
 * SURVEY 8(f) "next" row 4: what the reference does per read with std::string appends on the worker threads
 * (class Sam, /root/reference/src/sam.cpp:31-206 and src/sam.hpp:81-135; callers src/aln.cpp, output src/pc.cpp:119-135)
 * -- QNAME with the /1 /2 suffix stripped, FLAG, RNAME, POS, MAPQ, CIGAR text (=/X or M operations), RNEXT, PNEXT,
 * TLEN, SEQ (reverse-complemented for reverse-strand records, src/revcomp.hpp:10-41), QUAL (reversed), NM, AS, the
 * optional detail tags and the read-group tail -- for a whole batch of records in two kernels (line lengths -> scan ->
 * text).  The text is byte-identical to what Sam::add_record / add_unmapped / add_unmapped_mate append.
 *
 * Record-level entry (rsa_sam_format) mirrors Sam::add_record's argument list; rsa_sam_single / rsa_sam_pair are host
 * helpers that fill records the way Sam::add (src/sam.cpp:117-139) and Sam::add_pair (src/sam.cpp:208-318) compute
 * flags, mate fields and the template length.  Plain pointers and sizes only.  Status codes are rsa_ext.h's.
 */
#ifndef RSA_SAM_H
#define RSA_SAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsa_sam rsa_sam_t;

#define RSA_SAM_ALIGNED 0        /* Sam::add_record        (src/sam.cpp:141-206) */
#define RSA_SAM_UNMAPPED 1       /* Sam::add_unmapped      (src/sam.cpp:73-86)   */
#define RSA_SAM_UNMAPPED_MATE 2  /* Sam::add_unmapped_mate (src/sam.cpp:88-110)  */

#define RSA_SAM_REF_NONE (-1)    /* "*" */
#define RSA_SAM_REF_SAME (-2)    /* "=" (RNEXT only) */

typedef struct {
    uint32_t kind;           /* RSA_SAM_* */
    uint32_t flags;          /* SAM FLAG (src/sam.hpp:47-60) */
    int32_t ref_id;          /* RNAME: index into the names given to rsa_sam_create, or RSA_SAM_REF_NONE;
                                UNMAPPED_MATE: the mapped mate's reference */
    uint32_t pos;            /* 0-based; printed as pos + 1 in 32-bit unsigned arithmetic like the reference */
    uint32_t mapq;           /* uint8_t in the reference */
    int32_t mate_ref;        /* RNEXT: index, RSA_SAM_REF_SAME or RSA_SAM_REF_NONE */
    uint32_t mate_pos;       /* printed as mate_pos + 1 (0xFFFFFFFF prints 0) */
    int32_t tlen;
    int32_t edit_distance;   /* NM:i */
    int32_t score;           /* AS:i */
    uint32_t cigar_off;      /* first op in the cigar pool; ops are BAM-style len << 4 | op (src/cigar.hpp:11-21) */
    uint32_t n_cigar;        /* 0 prints "*" */
    uint32_t name_len, seq_len, qual_len;
    uint32_t details[5];     /* na nr al ga mr (printed only with show_details; mr only for paired records) */
    uint64_t name_off, seq_off, qual_off; /* into the text pool: the read's name, sequence and quality AS READ from the
                                             FASTQ; the formatter reverse-complements / reverses them for REVERSE
                                             records and prints "*" for SECONDARY ones */
} rsa_sam_record_t; /* 96 bytes */

/* The fields of `Alignment` (src/sam.hpp:11-25) the SAM writer reads. */
typedef struct {
    int32_t ref_id, ref_start, edit_distance, score, length;
    int32_t is_rc, is_unaligned;
    uint32_t cigar_off, n_cigar;
} rsa_sam_alignment_t;

/* The read as the FASTQ parser delivers it (klibpp::KSeq name / seq / qual) as offsets into the text pool. */
typedef struct {
    uint64_t name_off, seq_off, qual_off;
    uint32_t name_len, seq_len, qual_len;
} rsa_sam_read_t;

/* names_buf/names_off: the reference names (References::names), n_refs + 1 offsets.
 * cigar_m != 0: print M operations (CigarOps::M: Cigar::to_m, src/cigar.cpp:6-18) instead of = and X.
 * read_group: NULL or "" for none, else the tail "\tRG:Z:<id>" (src/sam.hpp:96-101). */
int rsa_sam_create(int32_t device, int32_t n_refs, const char *names_buf, const int64_t *names_off, int32_t cigar_m,
                   const char *read_group, int32_t output_unmapped, int32_t show_details, rsa_sam_t **out);
void rsa_sam_destroy(rsa_sam_t *h);
const char *rsa_sam_last_error(const rsa_sam_t *h);

/* Format n records.  text_pool: names, sequences, qualities; cigar_pool: all CIGAR ops.  The lines are written back to
 * back into out[0..out_cap); line_off (n + 1 entries, may be NULL) receives every line's start.  *out_len = total bytes;
 * if it exceeds out_cap nothing is copied and RSA_EXT_ERR_ARG is returned (call again with a larger buffer). Blocking.
 * Host arrays may be pinned or pageable (pageable input pools and a pageable `out` go through pinned bounce buffers of the handle). */
int rsa_sam_format(rsa_sam_t *h, int64_t n, const rsa_sam_record_t *records, const char *text_pool, int64_t text_bytes,
                   const uint32_t *cigar_pool, int64_t n_cigar_ops, char *out, int64_t out_cap, int64_t *out_len,
                   int64_t *line_off);

/* Device time of the kernels of the last rsa_sam_format (CUDA events: length + scan kernels, writer). */
double rsa_sam_kernel_ms(const rsa_sam_t *h);

/* Sam::add (src/sam.cpp:117-139): one aligned single-end record. */
void rsa_sam_single(const rsa_sam_alignment_t *a, const rsa_sam_read_t *read, uint32_t mapq, int32_t is_primary,
                    const uint32_t details[5], rsa_sam_record_t *out);
/* Sam::add_pair (src/sam.cpp:208-318): the two records of a pair (either may be unaligned, not both). */
void rsa_sam_pair(const rsa_sam_alignment_t *a1, const rsa_sam_alignment_t *a2, const rsa_sam_read_t *r1,
                  const rsa_sam_read_t *r2, uint32_t mapq1, uint32_t mapq2, int32_t is_proper, int32_t is_primary,
                  const uint32_t details1[5], const uint32_t details2[5], rsa_sam_record_t out[2]);
/* Sam::add_unmapped (src/sam.cpp:73-86) / add_unmapped_pair (:112-115). */
void rsa_sam_unmapped(const rsa_sam_read_t *read, uint32_t flags, rsa_sam_record_t *out);

#ifdef __cplusplus
}
#endif
#endif
