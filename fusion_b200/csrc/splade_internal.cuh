// Internal interface between sparse.cu (tail-code kernel, shares the posting-list walkers) and splade.cu (pipeline).
#pragma once

#include "common.cuh"

namespace fz {

struct TailCodeArgs {
    fz_postings_t ix;            // TAIL terms only (head terms have empty lists), no dense rows, tile_docs % 256 == 0
    const int32_t* q_ptr;
    const int32_t* q_term;
    const float* q_weight;
    int n_queries;
    int q_pad;                   // n_queries rounded up to 128
    long long r_lo, r_hi;        // doc range of the round; r_lo % 256 == 0
    const float2* qparam;        // [n_queries] (gh, g)
    uint32_t* codes;             // [((r_hi - r_lo + 255) / 256 * 8 + chunk) * q_pad + q] x 4 words of 8 codes
    int32_t* status;             // FZ_STATUS_FALLBACK when a tail sum left the code range
};

// one CTA per (query, group of FZ_COARSE_TILES tail tiles) of the round
int launch_tail_codes(const TailCodeArgs& A, cudaStream_t stream);

}  // namespace fz
