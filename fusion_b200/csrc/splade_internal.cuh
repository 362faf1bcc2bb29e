// Internal interface between sparse.cu (tail-code kernel, shares the posting-list walkers) and splade.cu (pipeline).
#pragma once

#include "common.cuh"
#include "topk_state.cuh"

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
    uint32_t* codes;             // [((doc - r_lo) / 256) * q_pad + q][8] x 4 words of 8 codes (GemmArgs::codes)
    int32_t* status;             // FZ_STATUS_FALLBACK when a tail sum left the code range
    int debug;                   // timing experiments only (FZ_DEBUG_TAIL): 1 = no global stores, 2 = no posting walk
};

// one CTA per (query, group of FZ_COARSE_TILES tail tiles) of the round
int launch_tail_codes(const TailCodeArgs& A, cudaStream_t stream);

// Threshold bootstrap: the general inverted-index kernel over a small index of the shard's FIRST documents (all terms),
// geometric rounds + cand_select into an initialised candidate state.  Leaves the k best of those documents (scores of the
// fixed-point kernel when `fixed_point`: the caller rescores them) and their k-th score as the threshold.
int sparse_bootstrap_f32(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, int k, const CandState<float>& st, bool fixed_point, cudaStream_t stream);

}  // namespace fz
