// integration/sam_glue.hpp -- binding of the device-side SAM formatter (include/rsa_sam.h) into the reference's host
// pipeline (the caller half).
//
// The reference formats every record on the worker thread: Sam::add / add_pair / add_unmapped* compute the flags and mate
// fields and end in one of three functions that append the text to the chunk's std::string (Sam::add_record
// src/sam.cpp:141-206, Sam::add_unmapped :73-86, Sam::add_unmapped_mate :88-110).  integration/patch_sam.py puts one call at
// the top of each of the three: while a chunk is being collected (sam_begin .. sam_flush, inserted into src/pc.cpp around the
// loops that call align_SE_read_last / align_PE_read_last) the call stores the function's ARGUMENTS as one rsa_sam_record_t
// (+ the read's name / sequence / quality and the CIGAR operations in two pools) and the function returns; sam_flush formats
// the chunk's records in one rsa_sam_format call and appends the text to the chunk's string.  Everything that decides WHAT is
// written stays the reference's code; only the number/CIGAR/reverse-complement formatting moved.
#ifndef RSA_SAM_GLUE_HPP
#define RSA_SAM_GLUE_HPP
#include <cstdint>
#include <string>

#include "cigar.hpp"
#include "refs.hpp"
#include "sam.hpp"

namespace rsa_glue {

// The writer configuration a record was produced under (the Sam object's private members, handed over by the patched
// member functions).
struct SamWriter {
    std::string& sam_string;
    const References& references;
    CigarOps cigar_ops;
    const std::string& tail;
    bool output_unmapped;
    bool show_details;
};

// Start / finish collecting the records appended to `sam_string` by this thread.
void sam_begin(int thread_id, std::string& sam_string);
void sam_flush(int thread_id, std::string& sam_string);

// true: the record was collected (the caller returns); false: no collection is active for this string (host code runs).
bool sam_collect_record(const SamWriter& w, const std::string& query_name, uint16_t flags, const std::string& reference_name,
                        uint32_t pos, uint8_t mapq, const Cigar& cigar, const std::string& mate_reference_name, uint32_t mate_pos,
                        int32_t template_len, const std::string& query_sequence, const std::string& qual, int ed, int aln_score,
                        const Details& details);
bool sam_collect_unmapped(const SamWriter& w, const klibpp::KSeq& record, uint16_t flags);
bool sam_collect_unmapped_mate(const SamWriter& w, const klibpp::KSeq& record, uint16_t flags, const std::string& mate_reference_name,
                               uint32_t mate_pos);

}  // namespace rsa_glue
#endif
