// K2: BM25 / TF-IDF / ATIRE and SPLADE scoring as sparse accumulation over a term-major CSR inverted index.
//
// The document axis is cut into tiles whose score accumulators live in shared memory.  One CTA scores one
// (doc tile, query) pair: it walks the query's terms IN QUERY ORDER (fp64 lexical scores are summed left to
// right like src/retrievers/bm25.py:152-155; postings of one term hit distinct docs, so there are no atomics and
// the sum is deterministic), then scans the tile and appends every doc that beats the query's running threshold
// to the candidate buffer (topk_state.cuh).  CTAs are ordered tile-major, so all queries stream the same slice of
// the posting lists at the same time and the slice is served from L2 after its first read.
//
// Lexical postings carry the precomputed fp64 impact of the (term, doc) pair — the reference's per-pair formula
// evaluated once at index time with the same operation order and no FMA contraction (fz_lexical_impacts).
#include "common.cuh"
#include "topk_state.cuh"

#include <limits>

namespace fz {

constexpr int kSparseThreads = 128;      // small CTAs: the per-term barrier only couples 4 warps, 16 CTAs per SM hide latency
constexpr int kMaxTermsPerPass = 64;

// ------------------------------------------------------------------------------------ index-time helpers
__global__ void lexical_impacts_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_doc,
                                       const int32_t* __restrict__ post_tf, const int32_t* __restrict__ doc_len,
                                       const double* __restrict__ idf, int n_terms, long long nnz, double avgdl,
                                       double k1, double k1p1, double one_minus_b, double b, int variant,
                                       double* __restrict__ out) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_terms;     // last t with term_ptr[t] <= p
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (term_ptr[mid] <= p) lo = mid; else hi = mid;
        }
        const double w = idf[lo];
        const double tf = (double)post_tf[p];
        double v;
        if (variant == FZ_LEX_TFIDF) {
            v = __dmul_rn(tf, w);                                                     // bm25.py:114
        } else {
            const double dl = (double)doc_len[post_doc[p]];
            const double num = __dmul_rn(w, __dmul_rn(tf, k1p1));                     // idf * (tf * (k1 + 1))
            const double inner = __dadd_rn(one_minus_b, __ddiv_rn(__dmul_rn(b, dl), avgdl));   // 1 - b + b*dl/avgdl
            const double den = __dadd_rn(tf, __dmul_rn(k1, inner));                   // tf + k1 * (...)
            v = __ddiv_rn(num, den);                                                  // bm25.py:155
        }
        out[p] = v;
    }
}

__global__ void long_tile_offsets_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_doc,
                                         const int32_t* __restrict__ long_terms, int n_long, int tile_docs,
                                         int n_tiles, uint32_t* __restrict__ out) {
    const long long total = (long long)n_long * (n_tiles + 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / (n_tiles + 1)), t = (int)(i % (n_tiles + 1));
        const int term = long_terms[r];
        const long long base = term_ptr[term], end = term_ptr[term + 1];
        const long long first_doc = (long long)t * tile_docs;
        long long lo = base, hi = end;   // first posting with doc >= first_doc
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if ((long long)post_doc[mid] < first_doc) lo = mid + 1; else hi = mid;
        }
        out[i] = (uint32_t)(lo - base);
    }
}

// ------------------------------------------------------------------------------------ tile scoring
template <typename AccT>
struct SparseArgs {
    fz_postings_t ix;
    const int32_t* q_ptr;
    const int32_t* q_term;
    const float* q_weight;    // f32 only
    int n_queries;
    int tile_lo;              // first tile of this launch
    long long r_lo, r_hi;     // doc range of this round
    int sign_mode;            // +1: emit score > 0, -1: emit score < 0
    CandState<AccT> st;
    AccT* out_full;           // full mode: [n_queries, n_docs]
    // zero-fill
    int k;
    long long doc_base;
    AccT* out_scores;
    int32_t* out_ids;
};

template <typename AccT> struct ValOf;
template <> struct ValOf<double> { using type = double; };
template <> struct ValOf<float> { using type = float; };

// Accumulate one query's postings that fall into docs [d_lo, d_hi) into acc[0 .. d_hi-d_lo) (shared memory).
// Terms are applied one after the other: the barrier between terms keeps every doc's sum in query order and makes
// atomics unnecessary (one term never hits a doc twice).  (A register-prefetch of the next term's first chunk was
// measured slower: the kernel is bound by L2 traffic and issue slots, not by exposed latency.)
template <typename AccT>
__device__ __forceinline__ void accumulate_tile(const SparseArgs<AccT>& A, int q, int tile, long long d_lo,
                                                long long d_hi, AccT* acc, long long* t_lo, long long* t_hi,
                                                float* t_w) {
    using ValT = typename ValOf<AccT>::type;
    const ValT* __restrict__ vals = reinterpret_cast<const ValT*>(A.ix.post_val);
    const int32_t* __restrict__ docs = A.ix.post_doc;
    const int n = (int)(d_hi - d_lo);
    const int dl = (int)d_lo;
    {   // zero the tile's accumulators with 16-byte stores (the buffer is 16-byte aligned and padded)
        int4* z = reinterpret_cast<int4*>(acc);
        const int n16 = (n * (int)sizeof(AccT) + 15) / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_int4(0, 0, 0, 0);
    }
    const int qb = A.q_ptr[q], qe = A.q_ptr[q + 1];
    for (int pass = qb; pass < qe; pass += kMaxTermsPerPass) {
        const int nt = min(kMaxTermsPerPass, qe - pass);
        __syncthreads();
        // resolve every term's posting range for this tile in parallel (one latency chain, not one per term)
        if ((int)threadIdx.x < nt) {
            const int t = A.q_term[pass + threadIdx.x];
            long long lo = 0, hi = 0;
            if (t >= 0 && t < A.ix.n_terms) {
                const long long base = A.ix.term_ptr[t];
                const int lr = A.ix.long_row[t];
                if (lr >= 0) {
                    const uint32_t* o = A.ix.long_tile_off + (size_t)lr * (A.ix.n_tiles + 1) + tile;
                    lo = base + o[0];
                    hi = base + o[1];
                } else {
                    lo = base;
                    hi = A.ix.term_ptr[t + 1];
                }
            }
            t_lo[threadIdx.x] = lo;
            t_hi[threadIdx.x] = hi;
            t_w[threadIdx.x] = A.q_weight ? A.q_weight[pass + threadIdx.x] : 1.0f;
        }
        __syncthreads();
        for (int j = 0; j < nt; ++j) {
            const int32_t* __restrict__ dj = docs + t_lo[j];
            const ValT* __restrict__ vj = vals + t_lo[j];
            const int len = (int)(t_hi[j] - t_lo[j]);
            const float w = t_w[j];
            // two postings per thread and iteration, doc and value loads issued together (no load behind a branch)
            for (int p = threadIdx.x; p < len; p += 2 * blockDim.x) {
                const int p1 = p + blockDim.x;
                const bool ok1 = p1 < len;
                const int d0 = __ldg(dj + p);
                const ValT v0 = __ldg(vj + p);
                const int d1 = ok1 ? __ldg(dj + p1) : dl - 1;
                const ValT v1 = ok1 ? __ldg(vj + p1) : (ValT)0;
                const unsigned o0 = (unsigned)(d0 - dl), o1 = (unsigned)(d1 - dl);   // one unsigned compare covers both bounds
                if constexpr (std::is_same<AccT, double>::value) {
                    if (o0 < (unsigned)n) acc[o0] = __dadd_rn(acc[o0], v0);
                    if (o1 < (unsigned)n) acc[o1] = __dadd_rn(acc[o1], v1);
                } else {
                    if (o0 < (unsigned)n) acc[o0] = __fmaf_rn(v0, w, acc[o0]);
                    if (o1 < (unsigned)n) acc[o1] = __fmaf_rn(v1, w, acc[o1]);
                }
            }
            __syncthreads();   // the next term may hit the same docs
        }
    }
    __syncthreads();
}

template <typename AccT, int MODE>   // MODE 0: threshold emit, 1: store every score
__global__ void __launch_bounds__(kSparseThreads) sparse_tile_kernel(const SparseArgs<AccT> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AccT* acc = reinterpret_cast<AccT*>(smem_raw);
    __shared__ long long t_lo[kMaxTermsPerPass], t_hi[kMaxTermsPerPass];
    __shared__ float t_w[kMaxTermsPerPass];

    const int tile = A.tile_lo + blockIdx.x / A.n_queries;
    const int q = blockIdx.x % A.n_queries;
    long long d_lo = (long long)tile * A.ix.tile_docs;
    long long d_hi = min(d_lo + A.ix.tile_docs, (long long)A.ix.n_docs);
    if (MODE == 0) {
        d_lo = max(d_lo, A.r_lo);
        d_hi = min(d_hi, A.r_hi);
    }
    if (d_hi <= d_lo) return;
    accumulate_tile<AccT>(A, q, tile, d_lo, d_hi, acc, t_lo, t_hi, t_w);
    const int n = (int)(d_hi - d_lo);
    if (MODE == 1) {
        AccT* out = A.out_full + (size_t)q * A.ix.n_docs + d_lo;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = acc[i];
        return;
    }
    const AccT tau = A.st.tau[q];
    // scan the tile 16 bytes at a time; survivors are rare once tau has risen, so the common case is one compare
    constexpr int V = 16 / (int)sizeof(AccT);
    const AccT lim = A.sign_mode > 0 ? (tau > (AccT)0 ? tau : (AccT)0) : tau;
    for (int i0 = threadIdx.x * V; i0 < n; i0 += blockDim.x * V) {
        AccT v[V];
        *reinterpret_cast<int4*>(v) = *reinterpret_cast<const int4*>(acc + i0);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const AccT sc = v[u];
            const bool want = A.sign_mode > 0 ? (sc > lim) : (sc < (AccT)0 && sc > lim);
            if (want && i0 + u < n) cand_append<AccT>(A.st, q, sc, (int32_t)(d_lo + i0 + u));
        }
    }
}

// Queries with fewer than k positive-score docs: append zero-score docs in ascending doc-id order
// (the reference ranks every document; unmatched ones score exactly 0.0 and tie by index, bm25.py:103-105).
template <typename AccT>
__global__ void __launch_bounds__(kSparseThreads) sparse_zero_fill_kernel(const SparseArgs<AccT> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AccT* acc = reinterpret_cast<AccT*>(smem_raw);
    __shared__ long long t_lo[kMaxTermsPerPass], t_hi[kMaxTermsPerPass];
    __shared__ float t_w[kMaxTermsPerPass];
    __shared__ int s_warp[kSparseThreads / 32];
    __shared__ int s_found;

    const int q = blockIdx.x;
    // after the final select cnt[q] is the number of positive-score docs kept: all of them when there are < k
    const int n_have = A.st.cnt[q];
    if (n_have >= A.k || A.sign_mode < 0) return;           // negatives never need zero fill
    const int need = A.k - n_have;
    if (threadIdx.x == 0) {
        s_found = 0;
        A.st.status[q] |= FZ_STATUS_NEED_ZERO;
    }
    __syncthreads();
    for (int tile = 0; tile < A.ix.n_tiles; ++tile) {
        const long long d_lo = (long long)tile * A.ix.tile_docs;
        const long long d_hi = min(d_lo + A.ix.tile_docs, (long long)A.ix.n_docs);
        accumulate_tile<AccT>(A, q, tile, d_lo, d_hi, acc, t_lo, t_hi, t_w);
        const int n = (int)(d_hi - d_lo);
        for (int c0 = 0; c0 < n; c0 += blockDim.x) {
            const int i = c0 + threadIdx.x;
            const int z = (i < n) && (acc[i] == (AccT)0);
            const unsigned bal = __ballot_sync(0xffffffffu, z);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            __syncthreads();
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < kSparseThreads / 32; ++w) {
                const int v = s_warp[w];
                if (w < warp) before += v;
                total += v;
            }
            const int base = s_found;
            if (z) {
                const int r = base + before + __popc(bal & ((1u << lane) - 1));
                if (r < need) {
                    A.out_scores[(size_t)q * A.k + n_have + r] = (AccT)0;
                    A.out_ids[(size_t)q * A.k + n_have + r] = (int32_t)(A.doc_base + d_lo + i);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_found = base + total;
            __syncthreads();
            if (s_found >= need) return;
        }
    }
    if (threadIdx.x == 0 && s_found < need) A.st.status[q] |= FZ_STATUS_NEED_NEG;
}

static int check_index(const fz_postings_t* ix, size_t acc_bytes) {
    FZ_REQUIRE(ix && ix->term_ptr && ix->post_doc && ix->post_val && ix->long_row, "null index pointer");
    FZ_REQUIRE(ix->n_long == 0 || ix->long_tile_off, "long_tile_off missing");
    FZ_REQUIRE(ix->tile_docs >= 256 && ix->tile_docs % 256 == 0, "tile_docs=%d must be a positive multiple of 256",
               ix->tile_docs);
    FZ_REQUIRE((size_t)ix->tile_docs * acc_bytes <= 200 * 1024, "tile_docs=%d does not fit shared memory", ix->tile_docs);
    FZ_REQUIRE(ix->n_docs >= 1 && ix->n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(ix->n_tiles == (int)ceil_div<long long>(ix->n_docs, ix->tile_docs), "n_tiles inconsistent with n_docs");
    return FZ_OK;
}

template <typename AccT>
static int set_smem_attrs() {
    static bool done = false;
    if (!done) {
        FZ_CUDA(cudaFuncSetAttribute(sparse_tile_kernel<AccT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(sparse_tile_kernel<AccT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(sparse_zero_fill_kernel<AccT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    return FZ_OK;
}

template <typename AccT>
static int sparse_topk(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                       int n_queries, int k, int64_t doc_base, int cap, int growth, int sign_mode, AccT* out_scores,
                       int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    int rc = check_index(ix, sizeof(AccT));
    if (rc) return rc;
    FZ_REQUIRE(q_ptr && q_term && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192, "need 1 <= k, 2k <= cap <= 8192 (k=%d cap=%d)", k, cap);
    FZ_REQUIRE(growth >= 1 && growth <= 64, "growth=%d out of range", growth);
    FZ_REQUIRE(sign_mode == 1 || sign_mode == -1, "sign_mode must be +1 or -1");
    FZ_REQUIRE(ws && ws_bytes >= cand_state_bytes<AccT>(n_queries, cap), "workspace too small");
    if (n_queries == 0) return FZ_OK;
    rc = set_smem_attrs<AccT>();
    if (rc) return rc;

    SparseArgs<AccT> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = q_ptr;
    A.q_term = q_term;
    A.q_weight = q_weight;
    A.n_queries = n_queries;
    A.sign_mode = sign_mode;
    A.st = cand_state_carve<AccT>(ws, n_queries, cap, out_status);
    A.k = k;
    A.doc_base = doc_base;
    A.out_scores = out_scores;
    A.out_ids = out_ids;
    rc = cand_init<AccT>(A.st, n_queries, stream);
    if (rc) return rc;

    const size_t smem = (size_t)ix->tile_docs * sizeof(AccT);
    const long long N = ix->n_docs;
    long long lo = 0, hi = N < cap ? N : cap;
    while (true) {
        A.r_lo = lo;
        A.r_hi = hi;
        A.tile_lo = (int)(lo / ix->tile_docs);
        const int tile_hi = (int)ceil_div<long long>(hi, ix->tile_docs);
        const long long blocks = (long long)(tile_hi - A.tile_lo) * n_queries;
        FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
        {
            ProfScope prof(sizeof(AccT) == 8 ? "sparse_tile_f64" : "sparse_tile_f32", stream);
            sparse_tile_kernel<AccT, 0><<<(unsigned)blocks, kSparseThreads, smem, stream>>>(A);
        }
        FZ_LAUNCH_CHECK();
        const bool last = hi >= N;
        rc = cand_select<AccT>(A.st, n_queries, k, (AccT)0, last, doc_base, out_scores, out_ids, nullptr, stream);
        if (rc) return rc;
        if (last) break;
        lo = hi;
        hi = growth >= 2 ? hi * growth : hi + (cap - k);
        if (hi > N) hi = N;
    }
    ProfScope prof("sparse_zero_fill", stream);
    sparse_zero_fill_kernel<AccT><<<n_queries, kSparseThreads, smem, stream>>>(A);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

template <typename AccT>
static int sparse_scores(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, AccT* out, cudaStream_t stream) {
    int rc = check_index(ix, sizeof(AccT));
    if (rc) return rc;
    FZ_REQUIRE(q_ptr && q_term && out, "null pointer");
    if (n_queries == 0) return FZ_OK;
    rc = set_smem_attrs<AccT>();
    if (rc) return rc;
    SparseArgs<AccT> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = q_ptr;
    A.q_term = q_term;
    A.q_weight = q_weight;
    A.n_queries = n_queries;
    A.out_full = out;
    const long long blocks = (long long)ix->n_tiles * n_queries;
    FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
    sparse_tile_kernel<AccT, 1><<<(unsigned)blocks, kSparseThreads, (size_t)ix->tile_docs * sizeof(AccT), stream>>>(A);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz

using namespace fz;

extern "C" {

int fz_lexical_impacts(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* post_tf, const int32_t* doc_len,
                       const double* idf, int32_t n_terms, int64_t nnz, double avgdl, double k1, double b, int variant,
                       double* out_impact, fz_stream_t stream) {
    FZ_REQUIRE(term_ptr && post_doc && post_tf && idf && out_impact, "null pointer");
    FZ_REQUIRE(variant == FZ_LEX_TFIDF || variant == FZ_LEX_BM25, "unknown lexical variant %d", variant);
    FZ_REQUIRE(variant == FZ_LEX_TFIDF || doc_len, "BM25 needs doc_len");
    if (nnz == 0) return FZ_OK;
    int blocks = (int)(ceil_div<long long>(nnz, 256) < 148 * 32 ? ceil_div<long long>(nnz, 256) : 148 * 32);
    lexical_impacts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(term_ptr, post_doc, post_tf, doc_len, idf, n_terms,
                                                                     nnz, avgdl, k1, k1 + 1, 1 - b, b, variant, out_impact);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_long_tile_offsets(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* long_terms, int32_t n_long,
                         int32_t tile_docs, int32_t n_tiles, uint32_t* out_long_tile_off, fz_stream_t stream) {
    if (n_long == 0) return FZ_OK;
    FZ_REQUIRE(term_ptr && post_doc && long_terms && out_long_tile_off, "null pointer");
    long long total = (long long)n_long * (n_tiles + 1);
    int blocks = (int)(ceil_div<long long>(total, 256) < 148 * 32 ? ceil_div<long long>(total, 256) : 148 * 32);
    long_tile_offsets_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(term_ptr, post_doc, long_terms, n_long, tile_docs,
                                                                       n_tiles, out_long_tile_off);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

size_t fz_sparse_topk_workspace_bytes(int n_queries, int k, int cap, int is_f64) {
    (void)k;
    return is_f64 ? cand_state_bytes<double>(n_queries, cap) : cand_state_bytes<float>(n_queries, cap);
}

int fz_sparse_topk_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries, int k,
                       int64_t doc_base, int cap, int growth, int sign_mode, double* out_scores, int32_t* out_ids,
                       int32_t* out_status, void* ws, size_t ws_bytes, fz_stream_t stream) {
    return sparse_topk<double>(index, q_ptr, q_term, nullptr, n_queries, k, doc_base, cap, growth, sign_mode,
                               out_scores, out_ids, out_status, ws, ws_bytes, (cudaStream_t)stream);
}

int fz_sparse_topk_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                       int n_queries, int k, int64_t doc_base, int cap, int growth, int sign_mode, float* out_scores,
                       int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, fz_stream_t stream) {
    return sparse_topk<float>(index, q_ptr, q_term, q_weight, n_queries, k, doc_base, cap, growth, sign_mode,
                              out_scores, out_ids, out_status, ws, ws_bytes, (cudaStream_t)stream);
}

int fz_sparse_scores_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries,
                         double* out_scores, fz_stream_t stream) {
    return sparse_scores<double>(index, q_ptr, q_term, nullptr, n_queries, out_scores, (cudaStream_t)stream);
}

int fz_sparse_scores_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, float* out_scores, fz_stream_t stream) {
    return sparse_scores<float>(index, q_ptr, q_term, q_weight, n_queries, out_scores, (cudaStream_t)stream);
}

}  // extern "C"
