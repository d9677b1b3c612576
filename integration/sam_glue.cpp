// integration/sam_glue.cpp -- see sam_glue.hpp.
#include "sam_glue.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gasal2_ssw.h"
#include "rsa_ext.h"
#include "rsa_sam.h"

namespace rsa_glue {

namespace {

// One per worker thread: the chunk being collected and the device formatter (created with the first chunk's writer
// configuration; the pipeline uses one configuration per process).
struct Collector {
    std::string* target = nullptr;   // the chunk's string while collecting
    std::vector<rsa_sam_record_t> records;
    std::vector<char> text;          // names, sequences, qualities as read
    std::vector<uint32_t> cigars;
    size_t bound = 0;                // upper bound of the chunk's text
    size_t max_ref_name = 0;
    rsa_sam_t* h = nullptr;
    // configuration the handle was created with
    const References* references = nullptr;
    CigarOps cigar_ops = CigarOps::EQX;
    std::string tail;
    bool output_unmapped = true, show_details = false;
    ~Collector() { if (h) rsa_sam_destroy(h); }
};
thread_local Collector t_col;

[[noreturn]] void die(const char* what, const rsa_sam_t* h) {
    fprintf(stderr, "[RSA_EXT ERROR:] %s: %s\n", what, h ? rsa_sam_last_error(h) : "");
    exit(EXIT_FAILURE);
}

void ensure_handle(Collector& c, int thread_id, const SamWriter& w) {
    if (c.h) {
        if (c.references != &w.references || c.cigar_ops != w.cigar_ops || c.tail != w.tail || c.output_unmapped != w.output_unmapped ||
            c.show_details != w.show_details) {
            fprintf(stderr, "[RSA_EXT ERROR:] SAM glue: one writer configuration per process\n");
            exit(EXIT_FAILURE);
        }
        return;
    }
    std::string names;
    std::vector<int64_t> off;
    for (const std::string& n : w.references.names) {
        off.push_back((int64_t)names.size());
        names += n;
        if (n.size() > c.max_ref_name) c.max_ref_name = n.size();
    }
    off.push_back((int64_t)names.size());
    // tail = "\n" or "\tRG:Z:<id>\n" (src/sam.hpp:96-101)
    std::string rg;
    if (w.tail.size() > 7) rg = w.tail.substr(6, w.tail.size() - 7);
    if (rsa_sam_create(rsa_ext_veneer_device(thread_id), (int32_t)w.references.names.size(), names.data(), off.data(),
                       w.cigar_ops == CigarOps::M, rg.empty() ? nullptr : rg.c_str(), w.output_unmapped, w.show_details, &c.h) != RSA_EXT_OK)
        die("rsa_sam_create", c.h);
    c.references = &w.references; c.cigar_ops = w.cigar_ops; c.tail = w.tail;
    c.output_unmapped = w.output_unmapped; c.show_details = w.show_details;
}

thread_local int t_thread_id = 0;

// RNAME / RNEXT as the formatter wants them: an index into References::names, "*" or "=".  The reference passes
// references.names[id] itself (src/sam.cpp:139, :262-283), so the index is the element's position.
int32_t ref_index(const References& references, const std::string& name) {
    const std::vector<std::string>& names = references.names;
    if (!names.empty() && &name >= names.data() && &name < names.data() + names.size()) return (int32_t)(&name - names.data());
    if (name == "*") return RSA_SAM_REF_NONE;
    if (name == "=") return RSA_SAM_REF_SAME;
    for (size_t i = 0; i < names.size(); ++i)
        if (names[i] == name) return (int32_t)i;
    fprintf(stderr, "[RSA_EXT ERROR:] SAM glue: unknown reference name %s\n", name.c_str());
    exit(EXIT_FAILURE);
}

void put_read(Collector& c, rsa_sam_record_t& r, const std::string& name, const std::string& seq, const std::string& qual) {
    r.name_off = c.text.size(); r.name_len = (uint32_t)name.size();
    c.text.insert(c.text.end(), name.begin(), name.end());
    r.seq_off = c.text.size(); r.seq_len = (uint32_t)seq.size();
    c.text.insert(c.text.end(), seq.begin(), seq.end());
    r.qual_off = c.text.size(); r.qual_len = (uint32_t)qual.size();
    c.text.insert(c.text.end(), qual.begin(), qual.end());
    c.bound += name.size() + seq.size() + qual.size() + 2 * c.max_ref_name + c.tail.size() + 256;
}

}  // namespace

void sam_begin(int thread_id, std::string& sam_string) {
    Collector& c = t_col;
    if (getenv("RSA_EXT_HOST_SAM")) return;
    c.target = &sam_string;
    c.records.clear(); c.text.clear(); c.cigars.clear();
    c.bound = 0;
    t_thread_id = thread_id;
}

bool sam_collect_record(const SamWriter& w, const std::string& query_name, uint16_t flags, const std::string& reference_name,
                        uint32_t pos, uint8_t mapq, const Cigar& cigar, const std::string& mate_reference_name, uint32_t mate_pos,
                        int32_t template_len, const std::string& query_sequence, const std::string& qual, int ed, int aln_score,
                        const Details& details) {
    Collector& c = t_col;
    if (c.target != &w.sam_string) return false;
    ensure_handle(c, t_thread_id, w);
    rsa_sam_record_t r;
    memset(&r, 0, sizeof r);
    r.kind = RSA_SAM_ALIGNED;
    r.flags = flags;
    r.ref_id = ref_index(w.references, reference_name);
    r.pos = pos;
    r.mapq = mapq;
    r.mate_ref = ref_index(w.references, mate_reference_name);
    r.mate_pos = mate_pos;
    r.tlen = template_len;
    r.edit_distance = ed;
    r.score = aln_score;
    r.cigar_off = (uint32_t)c.cigars.size();
    r.n_cigar = (uint32_t)cigar.m_ops.size();
    c.cigars.insert(c.cigars.end(), cigar.m_ops.begin(), cigar.m_ops.end());
    r.details[0] = (uint32_t)details.nams; r.details[1] = (uint32_t)details.nam_rescue; r.details[2] = (uint32_t)details.tried_alignment;
    r.details[3] = (uint32_t)details.gapped; r.details[4] = (uint32_t)details.mate_rescue;
    put_read(c, r, query_name, query_sequence, qual);
    c.bound += 11 * (size_t)r.n_cigar;
    c.records.push_back(r);
    return true;
}

bool sam_collect_unmapped(const SamWriter& w, const klibpp::KSeq& record, uint16_t flags) {
    Collector& c = t_col;
    if (c.target != &w.sam_string) return false;
    if (!w.output_unmapped) return true;   // src/sam.cpp:74-76
    ensure_handle(c, t_thread_id, w);
    rsa_sam_record_t r;
    memset(&r, 0, sizeof r);
    r.kind = RSA_SAM_UNMAPPED;
    r.flags = flags;
    r.ref_id = RSA_SAM_REF_NONE; r.mate_ref = RSA_SAM_REF_NONE;
    put_read(c, r, record.name, record.seq, record.qual);
    c.records.push_back(r);
    return true;
}

bool sam_collect_unmapped_mate(const SamWriter& w, const klibpp::KSeq& record, uint16_t flags, const std::string& mate_reference_name,
                               uint32_t mate_pos) {
    Collector& c = t_col;
    if (c.target != &w.sam_string) return false;
    ensure_handle(c, t_thread_id, w);
    rsa_sam_record_t r;
    memset(&r, 0, sizeof r);
    r.kind = RSA_SAM_UNMAPPED_MATE;
    r.flags = flags;
    r.ref_id = ref_index(w.references, mate_reference_name);
    r.mate_ref = RSA_SAM_REF_SAME;
    r.mate_pos = mate_pos;
    put_read(c, r, record.name, record.seq, record.qual);
    c.records.push_back(r);
    return true;
}

void sam_flush(int thread_id, std::string& sam_string) {
    (void)thread_id;
    Collector& c = t_col;
    if (c.target != &sam_string) return;
    c.target = nullptr;
    if (c.records.empty()) return;
    const size_t before = sam_string.size();
    sam_string.resize(before + c.bound);
    int64_t len = 0;
    c.text.resize(c.text.size() + 16);
    if (rsa_sam_format(c.h, (int64_t)c.records.size(), c.records.data(), c.text.data(), (int64_t)c.text.size() - 16, c.cigars.data(),
                       (int64_t)c.cigars.size(), &sam_string[before], (int64_t)c.bound, &len, nullptr) != RSA_EXT_OK)
        die("rsa_sam_format", c.h);
    sam_string.resize(before + (size_t)len);
}

}  // namespace rsa_glue
