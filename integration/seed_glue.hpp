// integration/seed_glue.hpp -- binding of the GPU seeding path (include/rsa_seed.h) into the reference's host pipeline.
//
// The reference seeds read by read on the worker thread (src/aln.cpp:1937-1958 in align_PE_read_part, :2384-2400 in
// align_SE_read_part): randstrobes_query -> find_nams -> find_nams_rescue.  With this glue a worker seeds a whole chunk
// in ONE GPU call before its per-read loop (rsa_glue::seed_chunk_se / _pe, inserted into src/pc.cpp by
// integration/patch_seed.py), and the three calls inside the two *_part functions are redirected (same patcher) to the
// functions below, which hand out the precomputed results in read order.  Everything after the NAM list (sort by score,
// shuffle, candidate selection, SAM) is the reference's unmodified host code, so the SAM stays byte-identical.
//
// The one host-side step: the reference appends the NAMs of one strand per reference id in the iteration order of its
// robin_hood::unordered_map (src/nam.cpp:775); the GPU returns the groups in first-touch order, and group_order() asks
// the reference's own container for its order when a strand touched more than one reference sequence.
#ifndef RSA_SEED_GLUE_HPP
#define RSA_SEED_GLUE_HPP
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "aln.hpp"
#include "index.hpp"
#include "nam.hpp"
#include "randstrobes.hpp"

namespace rsa_glue {

// Seed the reads `seqs` (in the order the per-read loop will consume them) on the GPU of worker `thread_id`.
void seed_chunk(int thread_id, const std::vector<const std::string*>& seqs, const IndexParameters& index_parameters,
                const StrobemerIndex& index, const MappingParameters& map_param);

template <class Rec>
void seed_chunk_se(int thread_id, const std::vector<Rec>& records, const IndexParameters& ip, const StrobemerIndex& index,
                   const MappingParameters& mp) {
    std::vector<const std::string*> seqs;
    seqs.reserve(records.size());
    for (const Rec& r : records) seqs.push_back(&r.seq);
    seed_chunk(thread_id, seqs, ip, index, mp);
}

template <class Rec>
void seed_chunk_pe(int thread_id, const std::vector<Rec>& records1, const std::vector<Rec>& records2, const IndexParameters& ip,
                   const StrobemerIndex& index, const MappingParameters& mp) {
    std::vector<const std::string*> seqs;
    seqs.reserve(2 * records1.size());
    for (size_t i = 0; i < records1.size(); ++i) {  // align_PE_read_part seeds record1, then record2
        seqs.push_back(&records1[i].seq);
        seqs.push_back(&records2[i].seq);
    }
    seed_chunk(thread_id, seqs, ip, index, mp);
}

// Stand-ins for the three calls inside align_SE_read_part / align_PE_read_part.
QueryRandstrobeVector randstrobes_query(const std::string_view seq, const IndexParameters& parameters);
std::pair<float, std::vector<Nam>> find_nams(const QueryRandstrobeVector& query_randstrobes, const StrobemerIndex& index);
std::vector<Nam> find_nams_rescue(const QueryRandstrobeVector& query_randstrobes, const StrobemerIndex& index, unsigned int rescue_cutoff);

}  // namespace rsa_glue
#endif
