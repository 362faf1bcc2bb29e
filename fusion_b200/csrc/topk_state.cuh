// Per-query candidate buffers shared by the threshold-filter top-k pipelines (dense K1, sparse K2).
//
// A scoring kernel appends every (score, doc) pair that beats the query's running threshold `tau`;
// between rounds `cand_select` sorts the buffer, tightens `tau` to the k-th best (minus a margin) and
// compacts the survivors.  Scores are never materialised as a [Q, N] matrix.
#pragma once

#include "common.cuh"

namespace fz {

template <typename ST>
struct CandState {
    ST* score;        // [n_queries, cap]
    int32_t* id;      // [n_queries, cap]   local doc rows
    int32_t* cnt;     // [n_queries]        appended so far (may exceed cap => overflow)
    ST* tau;          // [n_queries]        emit iff score > tau
    int32_t* status;  // [n_queries]        FZ_STATUS_* bits (caller-provided output)
    int32_t* npos;    // [n_queries]        docs with score > 0 seen (sparse only)
    int32_t* nneg;    // [n_queries]        docs with score < 0 seen (sparse only)
    int cap;
};

template <typename ST>
inline size_t cand_state_bytes(int n_queries, int cap) {
    size_t b = 0;
    b += align_up((size_t)n_queries * cap * sizeof(ST), 256);
    b += align_up((size_t)n_queries * cap * sizeof(int32_t), 256);
    b += 3 * align_up((size_t)n_queries * sizeof(int32_t), 256);
    b += align_up((size_t)n_queries * sizeof(ST), 256);
    return b;
}

template <typename ST>
inline CandState<ST> cand_state_carve(void* ws, int n_queries, int cap, int32_t* status) {
    char* p = (char*)ws;
    CandState<ST> s;
    s.score = (ST*)p;      p += align_up((size_t)n_queries * cap * sizeof(ST), 256);
    s.id = (int32_t*)p;    p += align_up((size_t)n_queries * cap * sizeof(int32_t), 256);
    s.cnt = (int32_t*)p;   p += align_up((size_t)n_queries * sizeof(int32_t), 256);
    s.npos = (int32_t*)p;  p += align_up((size_t)n_queries * sizeof(int32_t), 256);
    s.nneg = (int32_t*)p;  p += align_up((size_t)n_queries * sizeof(int32_t), 256);
    s.tau = (ST*)p;
    s.status = status;
    s.cap = cap;
    return s;
}

#ifdef __CUDACC__
// -0.0 is folded into +0.0 first: Python's sorted() / torch treat them as equal, so they must tie (by doc id)
__device__ __forceinline__ uint64_t score_key(float s) { return (uint64_t)ord32(s + 0.0f); }
__device__ __forceinline__ uint64_t score_key(double s) { return ord64(s + 0.0); }
__device__ __forceinline__ void key_score(uint64_t k, float& s) { s = unord32((uint32_t)k); }
__device__ __forceinline__ void key_score(uint64_t k, double& s) { s = unord64(k); }

// Append one candidate (any thread, any time).
template <typename ST>
__device__ __forceinline__ void cand_append(const CandState<ST>& st, int q, ST score, int32_t doc) {
    int idx = atomicAdd(&st.cnt[q], 1);
    if (idx < st.cap) {
        st.score[(size_t)q * st.cap + idx] = score;
        st.id[(size_t)q * st.cap + idx] = doc;
    }
}
#endif

// host-side launchers (select.cu)
template <typename ST>
int cand_init(const CandState<ST>& st, int n_queries, cudaStream_t stream);

// Sort each query's buffer; if it holds >= k entries set tau = (k-th best score) - margin and drop
// everything below tau.  With `final_out`, also write the best k as (score, doc_base + row) rows padded
// with (-inf, -1) and the number written to out_n (may be NULL).
template <typename ST>
int cand_select(const CandState<ST>& st, int n_queries, int k, ST margin, bool final_out, int64_t doc_base,
                ST* out_scores, int32_t* out_ids, int32_t* out_n, cudaStream_t stream, const ST* floor = nullptr);

// (rank-th best score of every query's buffer) - margin -> out [n_queries]; the buffers are not modified
template <typename ST>
int cand_kth_score(const CandState<ST>& st, int n_queries, int rank, ST margin, ST* out, cudaStream_t stream);

// Cross-shard threshold exchange of a sharded corpus (fz_shard_sync_t): between rounds every shard publishes the
// ceil(k / n_shards)-th best score it holds, the caller's hook all-reduces (MIN) the buffer, and the result becomes a
// floor under every shard's threshold.  Returns the floor to hand to cand_select, or nullptr when there is no hook.
template <typename ST>
inline const ST* shard_floor(const fz_shard_sync_t* sync, const CandState<ST>& st, int n_queries, int k, ST margin,
                             cudaStream_t stream, int* rc) {
    *rc = FZ_OK;
    if (!sync || !sync->hook) return nullptr;
    const int rank = sync->floor_rank > 0 ? sync->floor_rank : (k + sync->n_shards - 1) / sync->n_shards;
    *rc = cand_kth_score<ST>(st, n_queries, rank, margin, (ST*)sync->exchange, stream);
    if (*rc) return nullptr;
    if (sync->hook(sync->user) != 0) {
        set_error("shard sync hook failed");
        *rc = FZ_ERR_ARG;
        return nullptr;
    }
    return (const ST*)sync->exchange;
}

}  // namespace fz
