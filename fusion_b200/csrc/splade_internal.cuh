// Internal interface between sparse.cu (general inverted-index kernels) and splade.cu (head / tail pipeline).
#pragma once

#include "common.cuh"
#include "topk_state.cuh"

namespace fz {

// Threshold bootstrap: the general inverted-index kernel over a small index of the shard's FIRST documents (all terms),
// geometric rounds + cand_select into an initialised candidate state.  Leaves the k best of those documents (scores of the
// fixed-point kernel when `fixed_point`: the caller rescores them) and their k-th score as the threshold.
int sparse_bootstrap_f32(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, int k, const CandState<float>& st, bool fixed_point, cudaStream_t stream);

}  // namespace fz
