// K2c: SPLADE top-k as HEAD GEMM (tcgen05) + TAIL bound (4-bit codes) + EXACT fp32 rescoring.
//
// The reference scores SPLADE as a dense [Q, V] x [V, N] cosine (src/retrievers/hybrid.py:101-103,
// src/retrievers/splade/base.py:186-251).  On a Zipfian vocabulary > 95 % of all (query term, posting) pairs belong to
// the ~200 most frequent terms: walking those posting lists once per query is an L2-bandwidth problem (round 1:
// 1.6 TB through L2 -> SM per pass), while as a [Q x 192] x [192 x N] bf16 GEMM they cost a few milliseconds of tensor
// time.  The pipeline per round of documents:
//   1. tail_codes_kernel (sparse.cu)      tail terms' postings -> fixed-point sums -> a 4-bit upper bound per (query, doc)
//   2. filter_gemm_kernel<true>           head scores on the tensor cores; epilogue: head * gh + decode(code) > threshold?
//                                         survivors appended to the query's candidate buffer, flagged "pending"
//   3. sparse_rescore_kernel              pending candidates: exact fp32 sparse dot product from the doc-major CSR copy
//   4. cand_select (select.cu)            k-th best EXACT score -> the query's threshold for the next round
// The result holds exact fp32 scores only; the bf16 head and the quantised tail decide nothing but which docs get rescored,
// and both err on the side of keeping a doc (the bound is an upper bound: weights are non-negative).
#include "filter_gemm.cuh"
#include "splade_internal.cuh"

namespace fz {

// ----------------------------------------------------------------------------------- query preparation
// One CTA per query: the head part of the query as a bf16 row of the GEMM's A operand, and the two gains:
//   g  = kCodeTop / (sum over tail terms of w_q * max_d w_d): the largest tail the query can reach maps to the top code
//   gh = g * (1 + c): c covers the bf16 rounding of both operands (2^-8 relative on a sum of non-negative products)
//        and the tensor core's fp32 accumulation, so head * gh >= exact head * g.
constexpr int kPrepThreads = 128;
__global__ void __launch_bounds__(kPrepThreads)
splade_query_prep_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term, const float* __restrict__ q_weight,
                         const int32_t* __restrict__ term_head, const float* __restrict__ term_max, int n_terms, int head_dim,
                         __nv_bfloat16* __restrict__ qh, float2* __restrict__ qparam, int32_t* __restrict__ status) {
    __shared__ float hq[256];
    __shared__ float s_tail[kPrepThreads / 32];
    __shared__ int s_neg;
    const int q = blockIdx.x;
    for (int j = threadIdx.x; j < head_dim; j += kPrepThreads) hq[j] = 0.f;
    if (threadIdx.x == 0) s_neg = 0;
    __syncthreads();
    float tail = 0.f;
    for (int i = q_ptr[q] + threadIdx.x; i < q_ptr[q + 1]; i += kPrepThreads) {
        const int term = q_term[i];
        if (term < 0 || term >= n_terms) continue;
        const float w = q_weight ? q_weight[i] : 1.0f;
        if (!(w >= 0.f)) s_neg = 1;
        const int h = term_head[term];
        if (h >= 0) atomicAdd(&hq[h], w);
        else tail = fmaf(w, term_max[term], tail);
    }
    tail = warp_sum(tail);
    if ((threadIdx.x & 31) == 0) s_tail[threadIdx.x >> 5] = tail;
    __syncthreads();
    for (int j = threadIdx.x; j < head_dim; j += kPrepThreads) qh[(size_t)q * head_dim + j] = __float2bfloat16_rn(hq[j]);
    if (threadIdx.x == 0) {
        float ts = 0.f;
        for (int w = 0; w < kPrepThreads / 32; ++w) ts += s_tail[w];
        float g = ts > 0.f ? kCodeTop / (ts * 1.0002f) : 1.0f;
        g = fminf(g, 1048576.0f);
        const float c = 0.00390625f + 0.00004f + (float)head_dim * 2.4e-7f;
        qparam[q] = make_float2(g * (1.0f + c), g);
        if (s_neg) atomicOr(&status[q], FZ_STATUS_FALLBACK);
    }
}

// ----------------------------------------------------------------------------------- exact rescoring
// One CTA per query.  The query's terms sit in a small open-addressing table in shared memory; a warp takes one pending
// candidate (id flagged by the filter), streams the doc's (term, weight) pairs with coalesced 8-byte loads and sums
// w_q * w_d over the shared terms in fp32.  ~0.75 KB per candidate; a few thousand candidates per query and pass.
constexpr int kRsWarps = 8;
constexpr int kRsSlots = 512;       // >= 4 x FZ_MAX_QUERY_TERMS
__global__ void __launch_bounds__(kRsWarps * 32)
sparse_rescore_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term, const float* __restrict__ q_weight,
                      const int64_t* __restrict__ doc_ptr, const uint2* __restrict__ doc_post, CandState<float> st, int force_all) {
    __shared__ int s_key[kRsSlots];
    __shared__ float s_val[kRsSlots];
    const int q = blockIdx.x;
    const int n = min(st.cnt[q], st.cap);
    const size_t off = (size_t)q * st.cap;
    for (int i = threadIdx.x; i < kRsSlots; i += blockDim.x) { s_key[i] = -1; s_val[i] = 0.f; }
    __syncthreads();
    const int qb = q_ptr[q], qe = min(q_ptr[q + 1], qb + FZ_MAX_QUERY_TERMS);
    for (int i = qb + threadIdx.x; i < qe; i += blockDim.x) {
        const int term = q_term[i];
        if (term < 0) continue;
        const float w = q_weight ? q_weight[i] : 1.0f;
        unsigned slot = ((unsigned)term * 2654435761u) >> 23;      // top 9 bits
        while (true) {
            const int prev = atomicCAS(&s_key[slot], -1, term);
            if (prev == -1 || prev == term) { atomicAdd(&s_val[slot], w); break; }      // a repeated query term adds up
            slot = (slot + 1) & (kRsSlots - 1);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Software pipeline over this warp's candidates: the id of candidate i + 2 and the posting range of candidate i + 1 are
    // in flight while candidate i's postings are summed (each step of the chain id -> doc_ptr -> postings is a dependent
    // global load of ~1 us under load).
    auto load_id = [&](int i) { return i < n ? st.id[off + i] : 0; };
    auto is_pending = [&](int32_t id) { return force_all != 0 || id < 0; };
    int32_t id1 = load_id(warp), id2 = load_id(warp + kRsWarps);
    long long a1 = 0, b1 = 0;
    if (warp < n && is_pending(id1)) {
        const int d = (int)((uint32_t)id1 & ~kPendingBit);
        a1 = __ldg(doc_ptr + d);
        b1 = __ldg(doc_ptr + d + 1);
    }
    for (int i = warp; i < n; i += kRsWarps) {
        const int32_t id = id1;
        const long long p0 = a1, p1 = b1;
        id1 = id2;
        id2 = load_id(i + 2 * kRsWarps);
        if (i + kRsWarps < n && is_pending(id1)) {
            const int d = (int)((uint32_t)id1 & ~kPendingBit);
            a1 = __ldg(doc_ptr + d);
            b1 = __ldg(doc_ptr + d + 1);
        }
        if (!is_pending(id)) continue;                              // rescored in an earlier round
        float acc = 0.f;
        for (long long p = p0 + lane; p < p1; p += 128) {           // up to four independent loads in flight per lane
            uint2 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = p + 32 * u < p1 ? __ldg(doc_post + p + 32 * u) : make_uint2(0xffffffffu, 0u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (e[u].x == 0xffffffffu) continue;
                unsigned slot = (e[u].x * 2654435761u) >> 23;
                while (true) {
                    const int key = s_key[slot];
                    if (key == (int)e[u].x) { acc = fmaf(__uint_as_float(e[u].y), s_val[slot], acc); break; }
                    if (key == -1) break;
                    slot = (slot + 1) & (kRsSlots - 1);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            st.score[off + i] = acc;
            st.id[off + i] = (int32_t)((uint32_t)id & ~kPendingBit);
        }
    }
}

// after the final select: a query whose threshold is not positive has fewer than k positive-score docs (zero-score docs
// would have to fill up in doc-id order: the general path does that)
__global__ void splade_finalize_kernel(CandState<float> st, int n_queries) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_queries && !(st.tau[q] > 0.f)) st.status[q] |= FZ_STATUS_FALLBACK;
}

struct SpladeWs {
    CandState<float> st;
    __nv_bfloat16* qh;
    float2* qparam;
    uint32_t* codes;
    size_t code_bytes;
};

static size_t splade_fixed_bytes(int n_queries, int cap, int head_dim) {
    const size_t q_pad = (size_t)ceil_div(n_queries, kBM) * kBM;
    return align_up(cand_state_bytes<float>(n_queries, cap), 1024) + align_up(q_pad * head_dim * 2, 1024) +
           align_up((size_t)n_queries * sizeof(float2), 1024);
}

}  // namespace fz

using namespace fz;

extern "C" {

size_t fz_splade_topk_workspace_bytes(int n_queries, int k, int cap, int head_dim, int64_t max_round_docs) {
    (void)k;
    const size_t q_pad = (size_t)ceil_div(n_queries, kBM) * kBM;
    const size_t tiles = (size_t)ceil_div<long long>(max_round_docs > 256 ? max_round_docs : 256, 256);
    return splade_fixed_bytes(n_queries, cap, head_dim) + tiles * q_pad * 128;
}

int fz_splade_topk(const fz_postings_t* tail, const fz_splade_head_t* head, const fz_postings_t* boot, const int32_t* q_ptr,
                   const int32_t* q_term,
                   const float* q_weight, int n_queries, int k, int64_t doc_base, int cap, int growth, float* out_scores,
                   int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync,
                   fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(tail && head && q_ptr && q_term && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE(head->head_bf16 && head->term_head && head->term_max && head->doc_ptr && head->doc_post, "null head pointer");
    FZ_REQUIRE(head->head_dim >= 64 && head->head_dim <= 256 && head->head_dim % 64 == 0, "head_dim=%d must be 64, 128, 192 or 256",
               head->head_dim);
    FZ_REQUIRE(head->n_docs == tail->n_docs && head->n_terms == tail->n_terms, "head and tail index describe different shards");
    FZ_REQUIRE(head->n_docs >= 1 && head->n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192, "need 1 <= k, 2k <= cap <= 8192 (k=%d cap=%d)", k, cap);
    FZ_REQUIRE(growth >= 2 && growth <= 64, "growth=%d out of range", growth);
    if (n_queries == 0) return FZ_OK;
    const int head_dim = head->head_dim;
    const long long N = head->n_docs;
    const int q_pad = ceil_div(n_queries, kBM) * kBM;
    const size_t fixed = splade_fixed_bytes(n_queries, cap, head_dim);
    FZ_REQUIRE(ws && ws_bytes >= fixed + (size_t)q_pad * 128, "workspace too small");
    char* p = (char*)ws;
    CandState<float> st = cand_state_carve<float>(p, n_queries, cap, out_status);
    p += align_up(cand_state_bytes<float>(n_queries, cap), 1024);
    __nv_bfloat16* qh = (__nv_bfloat16*)p;
    p += align_up((size_t)q_pad * head_dim * 2, 1024);
    float2* qparam = (float2*)p;
    p += align_up((size_t)n_queries * sizeof(float2), 1024);
    uint32_t* codes = (uint32_t*)p;
    const long long max_round = (long long)((ws_bytes - fixed) / ((size_t)q_pad * 128)) * 256;

    int rc = cand_init<float>(st, n_queries, stream);
    if (rc) return rc;
    {
        ProfScope prof("splade_query_prep", stream);
        splade_query_prep_kernel<<<n_queries, kPrepThreads, 0, stream>>>(q_ptr, q_term, q_weight, head->term_head, head->term_max,
                                                                          head->n_terms, head_dim, qh, qparam, out_status);
        FZ_LAUNCH_CHECK();
    }
    CUtensorMap tmap_q, tmap_d;
    rc = make_bf16_tile_map(&tmap_q, qh, (uint64_t)n_queries, (uint64_t)head_dim, kBM);
    if (rc) return rc;
    rc = make_bf16_tile_map(&tmap_d, head->head_bf16, (uint64_t)N, (uint64_t)head_dim, kBN / kPair);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(filter_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        FZ_CUDA(cudaFuncSetAttribute(filter_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        attr = true;
    }
    GemmArgs G;
    memset(&G, 0, sizeof(G));
    G.n_queries = n_queries;
    G.num_k_blocks = head_dim / kBK;
    G.m_tiles = ceil_div(n_queries, kBM);
    G.st = st;
    G.stats = (unsigned long long*)g_debug_stats;
    if (const char* e = getenv("FZ_DEBUG_GEMM")) G.debug = atoi(e);
    G.codes = (const uint4*)codes;
    G.q_pad = q_pad;
    G.qparam = qparam;
    TailCodeArgs T;
    memset(&T, 0, sizeof(T));
    T.ix = *tail;
    T.q_ptr = q_ptr;
    T.q_term = q_term;
    T.q_weight = q_weight;
    T.n_queries = n_queries;
    T.q_pad = q_pad;
    T.qparam = qparam;
    T.codes = codes;
    T.status = out_status;
    { const char* e = getenv("FZ_DEBUG_TAIL"); T.debug = e ? atoi(e) : 0; }

    const bool synced = sync && sync->hook;
    FZ_REQUIRE(!synced || (sync->exchange && sync->n_shards >= 1 && sync->sched_docs >= N), "bad shard sync");
    const long long SN = synced ? (long long)sync->sched_docs : N;     // the schedule every shard follows
    auto rescore = [&](int force_all) -> int {
        ProfScope prof("splade_rescore", stream);
        sparse_rescore_kernel<<<n_queries, kRsWarps * 32, 0, stream>>>(q_ptr, q_term, q_weight, head->doc_ptr,
                                                                        (const uint2*)head->doc_post, st, force_all);
        FZ_LAUNCH_CHECK();
        return FZ_OK;
    };
    long long lo = 0, hi;
    if (boot) {
        // threshold bootstrap on the shard's first documents with the general kernel (no rescoring per emitted doc), then
        // its survivors are rescored exactly so that every score in the buffer - and the threshold - is an exact one
        FZ_REQUIRE(boot->n_docs >= 256 && boot->n_docs % 256 == 0 && boot->n_docs <= N && boot->n_terms == head->n_terms,
                   "bootstrap index must cover the first n (multiple of 256) docs of the shard");
        rc = sparse_bootstrap_f32(boot, q_ptr, q_term, q_weight, n_queries, k, st, (head->flags & FZ_SPLADE_UNIT_ROWS) != 0, stream);
        if (rc) return rc;
        rc = rescore(1);
        if (rc) return rc;
        rc = cand_select<float>(st, n_queries, k, 0.f, false, doc_base, nullptr, nullptr, nullptr, stream, nullptr);
        if (rc) return rc;
        lo = boot->n_docs;
        hi = lo * growth;
        if (hi - lo > max_round) hi = lo + max_round;
    } else {
        // the first round takes every doc that shares a term with the query (no threshold yet): it must fit the buffer
        hi = (long long)(cap / 256) * 256;
        if (hi > 2048 && 2048 >= 2 * k) hi = 2048;
        if (hi < 256) hi = 256;
        if (hi > max_round) hi = max_round;
        FZ_REQUIRE(hi >= 256 && hi <= cap, "cap=%d too small for a first round of 256 docs", cap);
    }
    if (hi > SN) hi = SN;
    while (true) {
        const long long r_lo = lo < N ? lo : N, r_hi = hi < N ? hi : N;
        if (r_hi > r_lo) {
            T.r_lo = r_lo;
            T.r_hi = r_hi;
            rc = launch_tail_codes(T, stream);
            if (rc) return rc;
            G.r_lo = r_lo;
            G.r_hi = r_hi;
            G.n_tiles = (int)ceil_div<long long>(r_hi - r_lo, kBN);
            const long long pair_tiles = (long long)ceil_div(G.m_tiles, kPair) * G.n_tiles;
            const int max_clusters = num_sms() / kPair;
            const int grid = kPair * (int)(pair_tiles < max_clusters ? pair_tiles : max_clusters);
            {
                ProfScope prof("splade_head_gemm", stream);
                if (gemm_two_cta()) filter_gemm_kernel<true, true><<<grid, kGemmThreads, kGemmSmem, stream>>>(tmap_q, tmap_d, G);
                else filter_gemm_kernel<true, false><<<grid, kGemmThreads, kGemmSmem, stream>>>(tmap_q, tmap_d, G);
                FZ_LAUNCH_CHECK();
            }
            rc = rescore(0);
            if (rc) return rc;
        }
        const bool last = hi >= SN;
        const float* floor = last ? nullptr : shard_floor<float>(sync, st, n_queries, k, 0.f, stream, &rc);
        if (rc) return rc;
        rc = cand_select<float>(st, n_queries, k, 0.f, last, doc_base, out_scores, out_ids, nullptr, stream, floor);
        if (rc) return rc;
        if (last) break;
        lo = hi;
        hi = hi * growth;
        if (hi - lo > max_round) hi = lo + max_round;
        if (hi > SN) hi = SN;
    }
    splade_finalize_kernel<<<ceil_div(n_queries, 256), 256, 0, stream>>>(st, n_queries);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
