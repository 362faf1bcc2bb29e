// K1: exhaustive dense inner-product scoring with the top-k fused into the tcgen05 GEMM epilogue.
//
// scores[q, n] = <Q[q,:], D[n,:]> is a [Qn x dim] x [dim x Nn] GEMM (9.5e13 FLOP for 6,980 x 8.8M x 768) whose
// 245 GB result must never reach HBM.  A warp-specialised persistent kernel (TMA producer warp, single-thread
// tcgen05.mma issuer, four epilogue warps reading the fp32 accumulators back from TMEM) filters every score
// against the query's running threshold and appends the few survivors to per-query candidate buffers
// (topk_state.cuh).  The corpus is swept in rounds of geometrically growing doc ranges; between rounds
// cand_select tightens the thresholds.  In the exact mode the bf16 tensor-core scores only FILTER: everything
// within `margin` of the running k-th score is kept and rescored in fp32 on CUDA cores from the fp32 originals,
// and the final top-k is taken from the exact scores.
//
// Replaces: sentence_transformers.util.semantic_search (src/retrievers/hybrid.py:103),
// BaseModel.search (src/retrievers/splade/base.py:199-251), compute_metrices scoring loop
// (src/utils/sentence_transformers.py:334-364).
#include "filter_gemm.cuh"

namespace fz {

// ----------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols] (128 bytes wide), 128-byte swizzle.
int make_bf16_tile_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    FZ_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
    FZ_REQUIRE(((uintptr_t)base & 15) == 0 && (cols * 2) % 16 == 0, "tensor base / row pitch must be 16-byte aligned");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FZ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FZ_OK;
}

// ----------------------------------------------------------------------------------- fp32 rescoring
// One warp per (query, kept candidate): exact fp32 dot product from the fp32 originals, in place.
constexpr int kRescoreWarps = 8;
__global__ void __launch_bounds__(kRescoreWarps * 32)
rescore_kernel(const float* __restrict__ qf, const float* __restrict__ df, int dim, CandState<float> st,
               const float* __restrict__ tau_floor) {
    extern __shared__ __align__(16) float qs[];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) qs[i] = qf[(size_t)q * dim + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = min(st.cnt[q], st.cap);
    const size_t off = (size_t)q * st.cap;
    // corpus-sharded runs: a candidate below the best shard's threshold cannot be in the global top-k - no need to
    // fetch its 3 KB row
    const float floor_q = tau_floor ? tau_floor[q] : -std::numeric_limits<float>::infinity();
    for (int i = warp; i < n; i += kRescoreWarps) {
        if (st.score[off + i] < floor_q) {
            if (lane == 0) st.score[off + i] = -std::numeric_limits<float>::infinity();
            continue;
        }
        const float* __restrict__ d = df + (size_t)st.id[off + i] * dim;
        float acc = 0.f;
        if ((dim & 3) == 0) {
            const float4* d4 = reinterpret_cast<const float4*>(d);
            const float4* q4 = reinterpret_cast<const float4*>(qs);
            for (int j = lane; j < dim / 4; j += 32) {
                const float4 a = __ldg(d4 + j), b = q4[j];
                acc = fmaf(a.x, b.x, acc);
                acc = fmaf(a.y, b.y, acc);
                acc = fmaf(a.z, b.z, acc);
                acc = fmaf(a.w, b.w, acc);
            }
        } else {
            for (int j = lane; j < dim; j += 32) acc = fmaf(d[j], qs[j], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) st.score[off + i] = acc;
    }
}

// ----------------------------------------------------------------------------------- exact fp32 scores (full mode)
// Plain SIMT tiled GEMM: out[q, n] = sum_k Q[q,k] * D[n,k], fp32 FMA in k order.  Used for the reference's
// "rank every document" mode on small corpora (src/retrievers/hybrid.py:103 with top_k = N), where the whole
// score matrix is wanted anyway.
constexpr int kFT = 64, kFK = 16;
__global__ void __launch_bounds__(256) dense_scores_kernel(const float* __restrict__ Q, const float* __restrict__ D,
                                                           int nq, long long nd, int dim, float* __restrict__ out) {
    __shared__ float qs[kFK][kFT + 1];
    __shared__ float ds[kFK][kFT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long n0 = (long long)blockIdx.x * kFT;
    const int q0 = blockIdx.y * kFT;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < dim; k0 += kFK) {
        for (int i = threadIdx.x; i < kFT * kFK; i += 256) {
            const int r = i / kFK, c = i - r * kFK;
            const int k = k0 + c;
            qs[c][r] = (q0 + r < nq && k < dim) ? Q[(size_t)(q0 + r) * dim + k] : 0.f;
            ds[c][r] = (n0 + r < nd && k < dim) ? D[(size_t)(n0 + r) * dim + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kFK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = qs[k][ty * 4 + i]; b[i] = ds[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int q = q0 + ty * 4 + i;
            const long long n = n0 + tx * 4 + j;
            if (q < nq && n < nd) out[(size_t)q * nd + n] = acc[i][j];
        }
}

// ----------------------------------------------------------------------------------- row-wise dot products
// out[i] = <Q[i,:], D[i,:]> in fp32 (compute_pairwise_similarity, src/retrievers/splade/base.py:173-184): one warp per row
__global__ void pairwise_dot_kernel(const float* __restrict__ q, const float* __restrict__ d, long long n_rows, int dim,
                                    float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    float acc = 0.f;
    for (int j = lane; j < dim; j += 32) acc = fmaf(q[(size_t)row * dim + j], d[(size_t)row * dim + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

// ----------------------------------------------------------------------------------- row normalisation
__global__ void normalize_rows_kernel(const float* __restrict__ x, long long n_rows, int dim, int normalize,
                                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float* r = x + (size_t)row * dim;
    float nrm = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int j = lane; j < dim; j += 32) ss = fmaf(r[j], r[j], ss);
        ss = warp_sum(ss);
        nrm = fmaxf(sqrtf(ss), 1e-12f);       // x / max(||x||, eps)
    }
    for (int j = lane; j < dim; j += 32) {
        const float v = normalize ? __fdiv_rn(r[j], nrm) : r[j];
        if (out_f32) out_f32[(size_t)row * dim + j] = v;
        if (out_bf16) out_bf16[(size_t)row * dim + j] = __float2bfloat16_rn(v);
    }
}

}  // namespace fz

using namespace fz;

extern "C" {

size_t fz_dense_topk_workspace_bytes(int n_queries, int k, int cap) {
    (void)k;
    return cand_state_bytes<float>(n_queries, cap);
}

static int dense_filter_phase(const void* q_bf16, const void* d_bf16, bool exact, int n_queries, int64_t n_docs, int dim,
                              int k, float margin, int64_t doc_base, int cap, int growth, float* out_scores,
                              int32_t* out_ids, int32_t* out_status, void* ws, const fz_shard_sync_t* sync,
                              cudaStream_t stream) {
    CUtensorMap tmap_q, tmap_d;
    int rc = make_bf16_tile_map(&tmap_q, q_bf16, (uint64_t)n_queries, (uint64_t)dim, kBM);
    if (rc) return rc;
    rc = make_bf16_tile_map(&tmap_d, d_bf16, (uint64_t)n_docs, (uint64_t)dim, kBN / kPair);    // each CTA of a pair loads half
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(filter_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        FZ_CUDA(cudaFuncSetAttribute(filter_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        attr = true;
    }
    GemmArgs G;
    memset(&G, 0, sizeof(G));
    G.n_queries = n_queries;
    G.num_k_blocks = dim / kBK;
    G.m_tiles = ceil_div(n_queries, kBM);
    G.st = cand_state_carve<float>(ws, n_queries, cap, out_status);
    G.stats = (unsigned long long*)g_debug_stats;
    if (const char* e = getenv("FZ_DEBUG_GEMM")) G.debug = atoi(e);
    rc = cand_init<float>(G.st, n_queries, stream);
    if (rc) return rc;
    const bool synced = sync && sync->hook;
    FZ_REQUIRE(!synced || (sync->exchange && sync->n_shards >= 1 && sync->sched_docs >= n_docs), "bad shard sync");
    const long long SN = synced ? (long long)sync->sched_docs : n_docs;     // the schedule every shard follows
    long long lo = 0, hi = SN < cap ? SN : cap;
    while (true) {
        G.r_lo = lo < n_docs ? lo : n_docs;
        G.r_hi = hi < n_docs ? hi : n_docs;
        if (G.r_hi > G.r_lo) {
            G.n_tiles = (int)ceil_div<long long>(G.r_hi - G.r_lo, kBN);
            const long long pair_tiles = (long long)ceil_div(G.m_tiles, kPair) * G.n_tiles;
            const int max_clusters = num_sms() / kPair;
            const int grid = kPair * (int)(pair_tiles < max_clusters ? pair_tiles : max_clusters);
            ProfScope prof("dense_filter_gemm", stream);
            if (gemm_two_cta()) filter_gemm_kernel<false, true><<<grid, kGemmThreads, kGemmSmem, stream>>>(tmap_q, tmap_d, G);
                else filter_gemm_kernel<false, false><<<grid, kGemmThreads, kGemmSmem, stream>>>(tmap_q, tmap_d, G);
        }
        FZ_LAUNCH_CHECK();
        const bool last = hi >= SN;
        const float* floor = last ? nullptr : shard_floor<float>(sync, G.st, n_queries, k, margin, stream, &rc);
        if (rc) return rc;
        rc = cand_select<float>(G.st, n_queries, k, margin, last && !exact, doc_base, out_scores, out_ids, nullptr, stream, floor);
        if (rc) return rc;
        if (last) break;
        lo = hi;
        hi = growth >= 2 ? hi * growth : hi + (cap - k);
        if (hi > SN) hi = SN;
    }
    return FZ_OK;
}

static int dense_finish_phase(const float* q_f32, const float* d_f32, const float* tau_floor, int n_queries, int dim, int k,
                              int64_t doc_base, int cap, float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws,
                              cudaStream_t stream) {
    CandState<float> st = cand_state_carve<float>(ws, n_queries, cap, out_status);
    {
        ProfScope prof("dense_rescore_f32", stream);
        rescore_kernel<<<n_queries, kRescoreWarps * 32, (size_t)dim * sizeof(float), stream>>>(q_f32, d_f32, dim, st, tau_floor);
    }
    FZ_LAUNCH_CHECK();
    return cand_select<float>(st, n_queries, k, 0.f, true, doc_base, out_scores, out_ids, nullptr, stream);
}

static int dense_check(const void* q_bf16, const void* d_bf16, const float* q_f32, const float* d_f32, int n_queries,
                       int64_t n_docs, int dim, int k, float margin, int cap, int growth, const float* out_scores,
                       const int32_t* out_ids, const int32_t* out_status, const void* ws, size_t ws_bytes) {
    FZ_REQUIRE(q_bf16 && d_bf16 && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE((q_f32 == nullptr) == (d_f32 == nullptr), "q_f32 and d_f32 must both be given (exact) or both NULL (bf16)");
    FZ_REQUIRE(dim >= 64 && dim % 64 == 0, "dim=%d must be a positive multiple of 64 (pad with zeros)", dim);
    FZ_REQUIRE(n_docs >= 1 && n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192, "need 1 <= k, 2k <= cap <= 8192 (k=%d cap=%d)", k, cap);
    FZ_REQUIRE(growth >= 1 && growth <= 64, "growth=%d out of range", growth);
    FZ_REQUIRE(margin >= 0.f, "margin must be >= 0");
    FZ_REQUIRE(ws && ws_bytes >= cand_state_bytes<float>(n_queries, cap), "workspace too small");
    return FZ_OK;
}

int fz_dense_topk(const void* q_bf16, const void* d_bf16, const float* q_f32, const float* d_f32, int n_queries,
                  int64_t n_docs, int dim, int k, float margin, int64_t doc_base, int cap, int growth,
                  float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes,
                  fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = dense_check(q_bf16, d_bf16, q_f32, d_f32, n_queries, n_docs, dim, k, margin, cap, growth, out_scores, out_ids,
                         out_status, ws, ws_bytes);
    if (rc) return rc;
    if (n_queries == 0) return FZ_OK;
    const bool exact = d_f32 != nullptr;
    rc = dense_filter_phase(q_bf16, d_bf16, exact, n_queries, n_docs, dim, k, margin, doc_base, cap, growth, out_scores,
                            out_ids, out_status, ws, nullptr, stream);
    if (rc || !exact) return rc;
    return dense_finish_phase(q_f32, d_f32, nullptr, n_queries, dim, k, doc_base, cap, out_scores, out_ids, out_status, ws, stream);
}

int fz_dense_topk_filter(const void* q_bf16, const void* d_bf16, int n_queries, int64_t n_docs, int dim, int k, float margin,
                         int64_t doc_base, int cap, int growth, int floor_rank, float* out_tau, int32_t* out_status, void* ws,
                         size_t ws_bytes, const fz_shard_sync_t* sync, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(out_tau, "null pointer");
    int rc = dense_check(q_bf16, d_bf16, nullptr, nullptr, n_queries, n_docs, dim, k, margin, cap, growth, out_tau,
                         out_status, out_status, ws, ws_bytes);
    if (rc) return rc;
    if (n_queries == 0) return FZ_OK;
    rc = dense_filter_phase(q_bf16, d_bf16, true, n_queries, n_docs, dim, k, margin, doc_base, cap, growth, nullptr, nullptr,
                            out_status, ws, sync, stream);
    if (rc) return rc;
    FZ_REQUIRE(floor_rank >= 1 && floor_rank <= k, "floor_rank=%d must be in [1, k]", floor_rank);
    const CandState<float> st = cand_state_carve<float>(ws, n_queries, cap, out_status);
    return cand_kth_score<float>(st, n_queries, floor_rank, margin, out_tau, stream);
}

int fz_dense_topk_finish(const float* q_f32, const float* d_f32, const float* tau_floor, int n_queries, int dim, int k,
                         int64_t doc_base, int cap, float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws,
                         size_t ws_bytes, fz_stream_t stream_) {
    FZ_REQUIRE(q_f32 && d_f32 && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192 && dim >= 1, "bad sizes");
    FZ_REQUIRE(ws && ws_bytes >= cand_state_bytes<float>(n_queries, cap), "workspace too small");
    if (n_queries == 0) return FZ_OK;
    return dense_finish_phase(q_f32, d_f32, tau_floor, n_queries, dim, k, doc_base, cap, out_scores, out_ids, out_status, ws,
                              (cudaStream_t)stream_);
}

int fz_dense_scores_f32(const float* q_f32, const float* d_f32, int n_queries, int64_t n_docs, int dim,
                        float* out_scores, fz_stream_t stream) {
    FZ_REQUIRE(q_f32 && d_f32 && out_scores, "null pointer");
    FZ_REQUIRE(dim >= 1 && n_docs >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    dim3 grid((unsigned)ceil_div<long long>(n_docs, kFT), (unsigned)ceil_div(n_queries, kFT));
    dense_scores_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q_f32, d_f32, n_queries, n_docs, dim, out_scores);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_pairwise_dot_f32(const float* q_f32, const float* d_f32, int64_t n_rows, int dim, float* out, fz_stream_t stream) {
    FZ_REQUIRE(q_f32 && d_f32 && out, "null pointer");
    FZ_REQUIRE(dim >= 1, "bad dim");
    if (n_rows == 0) return FZ_OK;
    pairwise_dot_kernel<<<(unsigned)ceil_div<long long>(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(q_f32, d_f32, n_rows, dim, out);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_normalize_rows(const float* x, int64_t n_rows, int dim, int normalize, float* out_f32, void* out_bf16,
                      fz_stream_t stream) {
    FZ_REQUIRE(x && (out_f32 || out_bf16), "null pointer");
    FZ_REQUIRE(dim >= 1, "bad dim");
    if (n_rows == 0) return FZ_OK;
    const int warps = 8;
    normalize_rows_kernel<<<(unsigned)ceil_div<long long>(n_rows, warps), warps * 32, 0, (cudaStream_t)stream>>>(
        x, n_rows, dim, normalize, out_f32, (__nv_bfloat16*)out_bf16);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
