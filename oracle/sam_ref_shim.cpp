// oracle/sam_ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's SAM writer (class Sam: /root/reference/src/sam.cpp, with src/cigar.cpp and src/revcomp.hpp) compiled
// from the reference's own sources where they lie (nothing copied; recipe in oracle/Makefile) behind a plain C entry:
// a list of "calls" -- Sam::add, Sam::add_pair, Sam::add_unmapped, Sam::add_unmapped_pair -- is replayed on one
// std::string exactly as the pipeline's workers make them (src/aln.cpp), and the text comes back.  The checker of
// rsa_sam_format (include/rsa_sam.h).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "sam.cpp"
#include "cigar.cpp"
#include "revcomp.hpp"

extern "C" {
struct SamRefAlignment {   // == rsa_sam_alignment_t
    int32_t ref_id, ref_start, edit_distance, score, length;
    int32_t is_rc, is_unaligned;
    uint32_t cigar_off, n_cigar;
};
struct SamRefRead {        // == rsa_sam_read_t
    uint64_t name_off, seq_off, qual_off;
    uint32_t name_len, seq_len, qual_len;
};
struct SamRefCall {
    int32_t kind;          // 0 add, 1 add_pair, 2 add_unmapped, 3 add_unmapped_pair
    int32_t is_primary, is_proper;
    uint32_t mapq1, mapq2, unmapped_flags;
    SamRefAlignment a1, a2;
    SamRefRead r1, r2;
    uint32_t details1[5], details2[5];
};

// returns the text length (or -needed if cap is too small)
long long sam_ref_replay(int n_refs, const char* names_buf, const int64_t* names_off, int cigar_m, const char* read_group,
                         int output_unmapped, int show_details, long long n_calls, const SamRefCall* calls, const char* text,
                         const uint32_t* cigars, char* out, long long cap) {
    std::vector<std::string> names, seqs;
    for (int i = 0; i < n_refs; ++i) { names.emplace_back(names_buf + names_off[i], names_buf + names_off[i + 1]); seqs.emplace_back(); }
    References references(seqs, names);
    std::string sam_string;
    Sam sam(sam_string, references, cigar_m ? CigarOps::M : CigarOps::EQX, read_group ? read_group : "", output_unmapped != 0, show_details != 0);
    auto kseq = [&](const SamRefRead& r) {
        klibpp::KSeq k;
        k.name.assign(text + r.name_off, r.name_len);
        k.seq.assign(text + r.seq_off, r.seq_len);
        k.qual.assign(text + r.qual_off, r.qual_len);
        return k;
    };
    auto aln = [&](const SamRefAlignment& a) {
        Alignment x;
        x.ref_id = a.ref_id; x.ref_start = a.ref_start; x.edit_distance = a.edit_distance; x.global_ed = a.edit_distance;
        x.score = a.score; x.length = a.length; x.is_rc = a.is_rc != 0; x.is_unaligned = a.is_unaligned != 0;
        x.cigar = Cigar(std::vector<uint32_t>(cigars + a.cigar_off, cigars + a.cigar_off + a.n_cigar));
        return x;
    };
    auto det = [&](const uint32_t d[5]) {
        Details x;
        x.nams = d[0]; x.nam_rescue = d[1] != 0; x.tried_alignment = d[2]; x.gapped = d[3]; x.mate_rescue = d[4];
        return x;
    };
    for (long long i = 0; i < n_calls; ++i) {
        const SamRefCall& c = calls[i];
        if (c.kind == 0) {
            const klibpp::KSeq rec = kseq(c.r1);
            const Read read(rec.seq);
            sam.add(aln(c.a1), rec, read.rc, (uint8_t)c.mapq1, c.is_primary != 0, det(c.details1));
        } else if (c.kind == 1) {
            const klibpp::KSeq rec1 = kseq(c.r1), rec2 = kseq(c.r2);
            const Read read1(rec1.seq), read2(rec2.seq);
            sam.add_pair(aln(c.a1), aln(c.a2), rec1, rec2, read1.rc, read2.rc, (uint8_t)c.mapq1, (uint8_t)c.mapq2, c.is_proper != 0,
                         c.is_primary != 0, std::array<Details, 2>{det(c.details1), det(c.details2)});
        } else if (c.kind == 2) {
            sam.add_unmapped(kseq(c.r1), (uint16_t)c.unmapped_flags);
        } else {
            sam.add_unmapped_pair(kseq(c.r1), kseq(c.r2));
        }
    }
    if ((long long)sam_string.size() > cap) return -(long long)sam_string.size();
    memcpy(out, sam_string.data(), sam_string.size());
    return (long long)sam_string.size();
}
}
