// common.cuh -- shared device/host declarations of the extension engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/rsa_ext.h"

namespace rsa {

// Per-pair record uploaded with every chunk.  Offsets are relative to the chunk's ASCII slices.
struct PairMeta {
    uint32_t qoff;
    uint32_t toff;
    uint16_t qlen;
    uint16_t tlen;
};

// Scoring of one handle (reference: __constant__ globals, GASAL2/src/gasal_kernels.h:29-33).
struct Scoring {
    int match;     // +match
    int mismatch;  // -mismatch
    int gap_oe;    // (gap_open-1) + gap_extend, cost of the first gap base (gasal_align.cu:332-337)
    int gap_ext;   // cost of every further gap base
};

// Internal per-pair DP result handed from the DP kernels to the traceback kernel.
struct DpEnd {
    int32_t score;
    int32_t qend;
    int32_t tend;
    uint32_t flags;  // DPF_*
};
enum : uint32_t {
    DPF_DONE = 1u,       // a DP kernel produced score/end and direction bits for this pair
    DPF_NEED_EXACT = 2u, // the fast kernel declined (tie on the maximum, unsupported symbol, ...)
    DPF_LAYOUT_FAST = 4u, // direction bits are in the fast kernel's layout
    DPF_NO_SCRATCH = 8u,  // the redo pass ran out of scratch for this pair: status 4, the host re-submits it
    DPF_TRACED = 16u      // the packed kernel already traced this pair back and wrote its record
};

// Direction-nibble code, identical to the reference's (local_kernel_template.h:45-60):
//   bits 1..0  0 = diagonal/match, 1 = diagonal/mismatch, 2 = from E (deletion), 3 = from F (insertion)
//   bit 2      E of the next row extends the gap (1) or opens it from the diagonal (0)
//   bit 3      F of the next column extends (1) or opens (0)
// Exact-kernel layout: row-major [target row][query column], two columns per byte (even column in the
// low nibble), row stride `dir_stride` bytes.

__host__ __device__ inline uint32_t nibble_of(uint8_t ascii) { return ascii & 0xFu; }  // pack_rc_seqs.h:13-53
constexpr uint32_t kWildcard = 0xEu;                                                    // 'N' & 0xF

}  // namespace rsa
