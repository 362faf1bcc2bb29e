// Retrieval metrics on ranked id lists and the linear-fusion weight sweep.
//
//   fz_rank_metrics  recall@k / map@k / mrr@k / ndcg@k / R-precision of [Q, n] ranked lists against gold id lists, summed
//                    over the queries - src/utils/metrics.py:40-162 with run_evaluation's metric set (hybrid.py:24-27).
//   fz_fuse_sweep    src/retrievers/hybrid.py:404-426: for EVERY weight vector of a grid, fuse the systems' (already
//                    normalised) lists by weighted sum and evaluate the fused ranking.  Normalisation does not depend on
//                    the weights, so one CTA per query builds the union table once, and per weight vector only computes
//                    the fused scores and, for each gold doc, how many union docs beat it (score desc, ties by first
//                    insertion = the reference's stable sort).  No sort, no [W, Q, union] materialisation.
//
// Every metric follows from the sorted ranks r_1 < r_2 < ... of the gold docs that were retrieved:
//   recall@k = #{r_j <= k} / |gold|          map@k = sum_{r_j <= k} (j / r_j) / |gold|       mrr@k = 1 / r_1 if r_1 <= k
//   ndcg@k   = sum_{r_j <= k} (r_j == 1 ? 1 : 1 / log2(r_j)) / (1 + sum_{i=1}^{|gold|-1} 1 / log2(i + 1))   (:98-111)
//   r-precision = #{r_j <= |gold|} / |gold|
#include "common.cuh"

#include <limits>

namespace fz {

constexpr int kMaxGold = 256;     // distinct gold documents per query (more: the outputs are poisoned with NaN, never truncated)
constexpr int kMaxKs = 8;         // cut-offs per metric family
constexpr int kMetricThreads = 256;
constexpr int kSweepThreads = 512;
constexpr int kSweepMaxSys = 4;
constexpr int kNoRank = 0x7fffffff;

struct MetricCfg {
    int n_recall, n_map, n_mrr, n_ndcg;
    int recall_k[kMaxKs], map_k[kMaxKs], mrr_k[kMaxKs], ndcg_k[kMaxKs];
    __host__ __device__ int count() const { return n_recall + n_map + n_mrr + n_ndcg + 1; }
};

// ranks: 1-based ranks of the retrieved gold docs, ascending, n_found of them; n_gold = len(gold) as the reference counts it
__device__ void metrics_from_ranks(const int* ranks, int n_found, int n_gold, const MetricCfg& C, double* out) {
    int m = 0;
    const double ng = (double)n_gold;
    for (int i = 0; i < C.n_recall; ++i) {
        int c = 0;
        for (int j = 0; j < n_found; ++j) c += ranks[j] <= C.recall_k[i];
        out[m++] = __ddiv_rn((double)c, ng);
    }
    for (int i = 0; i < C.n_map; ++i) {
        double s = 0.0;
        for (int j = 0; j < n_found; ++j)
            if (ranks[j] <= C.map_k[i]) s = __dadd_rn(s, __ddiv_rn((double)(j + 1), (double)ranks[j]));
        out[m++] = __ddiv_rn(s, ng);
    }
    for (int i = 0; i < C.n_mrr; ++i)
        out[m++] = (n_found > 0 && ranks[0] <= C.mrr_k[i]) ? __ddiv_rn(1.0, (double)ranks[0]) : 0.0;
    double idcg = 1.0;
    {
        double t = 0.0;
        for (int i = 1; i < n_gold; ++i) t = __dadd_rn(t, __ddiv_rn(1.0, log2((double)(i + 1))));
        idcg = __dadd_rn(1.0, t);
    }
    for (int i = 0; i < C.n_ndcg; ++i) {
        double rel0 = 0.0, t = 0.0;
        for (int j = 0; j < n_found; ++j) {
            if (ranks[j] > C.ndcg_k[i]) continue;
            if (ranks[j] == 1) rel0 = 1.0;
            else t = __dadd_rn(t, __ddiv_rn(1.0, log2((double)ranks[j])));
        }
        out[m++] = __ddiv_rn(__dadd_rn(rel0, t), idcg);
    }
    int c = 0;
    for (int j = 0; j < n_found; ++j) c += ranks[j] <= n_gold;
    out[m++] = __ddiv_rn((double)c, ng);
}

// distinct gold ids of query q into s_gold (first occurrence order); returns their number, n_gold = list length
__device__ int load_gold(const int32_t* gold_ptr, const int32_t* gold_ids, int q, int* s_gold, int& n_gold) {
    const int b = gold_ptr[q], e = gold_ptr[q + 1];
    n_gold = e - b;
    int n = 0;
    for (int i = b; i < e; ++i) {
        const int g = gold_ids[i];
        bool dup = false;
        for (int j = 0; j < n; ++j) dup |= s_gold[j] == g;
        if (dup) continue;
        if (n == kMaxGold) return -1;        // too many relevant docs for the shared-memory tables: the caller poisons the result
        s_gold[n++] = g;
    }
    return n;
}

__device__ void sort_ranks(int* r, int n) {      // insertion sort, n <= kMaxGold, unfound ranks (kNoRank) go last
    for (int i = 1; i < n; ++i) {
        const int v = r[i];
        int j = i - 1;
        while (j >= 0 && r[j] > v) { r[j + 1] = r[j]; --j; }
        r[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------------ metrics of ranked lists
__global__ void __launch_bounds__(kMetricThreads)
rank_metrics_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ lens, int stride, const int32_t* gold_ptr,
                    const int32_t* gold_ids, MetricCfg C, double* __restrict__ out_sum, double* __restrict__ per_query) {
    __shared__ int s_gold[kMaxGold], s_rank[kMaxGold];
    __shared__ int s_n, s_ngold;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) {
        int ng;
        s_n = load_gold(gold_ptr, gold_ids, q, s_gold, ng);
        s_ngold = ng;
    }
    for (int i = threadIdx.x; i < kMaxGold; i += blockDim.x) s_rank[i] = kNoRank;
    __syncthreads();
    const int n = s_n;
    if (n < 0) {            // more than kMaxGold distinct gold ids
        if (threadIdx.x == 0)
            for (int i = 0; i < C.count(); ++i) {
                if (per_query) per_query[(size_t)q * C.count() + i] = nan("");
                else atomicAdd(&out_sum[i], nan(""));
            }
        return;
    }
    const int len = lens ? min(lens[q], stride) : stride;
    const int32_t* row = ids + (size_t)q * stride;
    for (int p = threadIdx.x; p < len; p += blockDim.x) {
        const int d = row[p];
        if (d < 0) continue;
        for (int j = 0; j < n; ++j)
            if (s_gold[j] == d) atomicMin(&s_rank[j], p + 1);      // a repeated id counts at its first position
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m[4 * kMaxKs + 1];
        for (int i = 0; i < C.count(); ++i) m[i] = 0.0;
        if (s_ngold > 0) {
            sort_ranks(s_rank, n);
            int nf = 0;
            while (nf < n && s_rank[nf] != kNoRank) ++nf;
            metrics_from_ranks(s_rank, nf, s_ngold, C, m);
        }
        for (int i = 0; i < C.count(); ++i) {
            if (per_query) per_query[(size_t)q * C.count() + i] = m[i];
            else if (s_ngold > 0) atomicAdd(&out_sum[i], m[i]);
        }
    }
}

// out_sum[m] = sum over the queries in query order: the same bits every run (statistics.mean-style accumulation order)
__global__ void metrics_reduce_kernel(const double* __restrict__ per_query, int n_queries, int n_metrics, double* __restrict__ out_sum) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_metrics) return;
    double s = 0.0;
    for (int q = 0; q < n_queries; ++q) s = __dadd_rn(s, per_query[(size_t)q * n_metrics + m]);
    out_sum[m] = s;
}

// ------------------------------------------------------------------------------------------ weight sweep
struct SweepParams {
    const int32_t* ids[kSweepMaxSys];      // normalised lists: [n_queries, stride[s]], no repeated ids inside a list
    const double* val[kSweepMaxSys];       // normalised scores (fp32 values when f32_path)
    const int32_t* lens[kSweepMaxSys];
    int stride[kSweepMaxSys];
    int n_sys, n_queries, n_weights, table_slots, f32_path;
    const double* weights;                 // [n_weights, n_sys]
    const int32_t* gold_ptr;
    const int32_t* gold_ids;
    MetricCfg cfg;
    double* out_sum;                       // [n_weights, n_metrics]
};

__device__ __forceinline__ int sweep_hash(int key, int mask) {
    return (int)((((uint32_t)key * 2654435761u) >> 7) & (uint32_t)mask);
}

// fused score of one union entry: sum over the systems that hold it, in system order; fp32 (torch-normalised) or fp64
// (normalization 'none') values, always weighted and summed in fp64
template <bool F32>
__device__ __forceinline__ double fused_score(const float* const* t32, const double* const* t64, uint32_t meta, int slot,
                                              const double* w, int n_sys) {
    if (F32) {
        // the sweep's weights come from np.arange: np.float64 scalars, so np.float32 score * weight is float64 under every
        // NumPy version and aggregate_scores sums doubles (hybrid.py:291,302,405-409)
        double a = 0.0;
        for (int s = 0; s < n_sys; ++s)
            if (meta & (1u << (28 + s))) a = __dadd_rn(a, __dmul_rn((double)t32[s][slot], w[s]));
        return a;
    }
    double a = 0.0;
    for (int s = 0; s < n_sys; ++s)
        if (meta & (1u << (28 + s))) a = __dadd_rn(a, __dmul_rn(t64[s][slot], w[s]));
    return a;
}

template <bool F32>
__global__ void __launch_bounds__(kSweepThreads) fuse_sweep_kernel(const SweepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = P.table_slots, mask = H - 1;
    // layout: key[H] i32 | meta[H] u32 (bits 0-27 first-insertion order, bits 28-31 systems holding the id) | t[S][H]
    int* h_key = reinterpret_cast<int*>(smem_raw);
    uint32_t* h_meta = reinterpret_cast<uint32_t*>(h_key + H);
    unsigned char* tbase = reinterpret_cast<unsigned char*>(h_meta + H);
    const float* t32[kSweepMaxSys];
    const double* t64[kSweepMaxSys];
    for (int s = 0; s < kSweepMaxSys; ++s) {
        t32[s] = reinterpret_cast<const float*>(tbase) + (size_t)s * H;
        t64[s] = reinterpret_cast<const double*>(tbase) + (size_t)s * H;
    }
    __shared__ int s_gold[kMaxGold], s_slot[kMaxGold], s_cnt[kMaxGold], s_rank[kMaxGold];
    __shared__ double s_gscore[kMaxGold];
    __shared__ uint32_t s_gorder[kMaxGold];
    __shared__ double s_w[kSweepMaxSys];
    __shared__ int s_n, s_ngold, s_base;

    const int q = blockIdx.x;
    constexpr int kEmpty = (int)0x80000000;
    for (int i = threadIdx.x; i < H; i += blockDim.x) { h_key[i] = kEmpty; h_meta[i] = 0x0fffffffu; }
    if (threadIdx.x == 0) {
        int ng;
        s_n = load_gold(P.gold_ptr, P.gold_ids, q, s_gold, ng);
        s_ngold = ng;
        s_base = 0;
    }
    __syncthreads();
    if (s_n < 0) {          // more than kMaxGold distinct gold ids: poison every row, never truncate
        const int nm = P.cfg.count();
        for (int i = threadIdx.x; i < P.n_weights * nm; i += blockDim.x) atomicAdd(&P.out_sum[i], nan(""));
        return;
    }

    // ---- union table: systems one after the other, so "first insertion" = (system order, rank inside the system)
    for (int s = 0; s < P.n_sys; ++s) {
        const int n = P.lens[s] ? min(P.lens[s][q], P.stride[s]) : P.stride[s];
        const int32_t* ids = P.ids[s] + (size_t)q * P.stride[s];
        const double* val = P.val[s] + (size_t)q * P.stride[s];
        const int base = s_base;
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const int key = ids[p];
            if (key < 0) continue;
            int slot = sweep_hash(key, mask);
            while (true) {
                const int cur = h_key[slot];
                if (cur == key) break;
                if (cur == kEmpty) {
                    const int old = atomicCAS(&h_key[slot], kEmpty, key);
                    if (old == kEmpty || old == key) break;
                }
                slot = (slot + 1) & mask;
            }
            // one entry per (system, id): this thread is the only writer of the slot in this phase
            uint32_t m = h_meta[slot];
            if ((m & 0x0fffffffu) == 0x0fffffffu) m = (m & 0xf0000000u) | (uint32_t)(base + p);
            h_meta[slot] = m | (1u << (28 + s));
            if (F32) const_cast<float*>(t32[s])[slot] = (float)val[p];
            else const_cast<double*>(t64[s])[slot] = val[p];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + n;
        __syncthreads();
    }
    // the systems-present bits of empty slots stay 0; their order field is unused
    const int n_gold_u = s_n;
    for (int j = threadIdx.x; j < n_gold_u; j += blockDim.x) {
        const int key = s_gold[j];
        int slot = sweep_hash(key, mask), found = -1;
        while (true) {
            const int cur = h_key[slot];
            if (cur == key) { found = slot; break; }
            if (cur == kEmpty) break;
            slot = (slot + 1) & mask;
        }
        s_slot[j] = found;
    }
    __syncthreads();
    if (s_ngold == 0) return;

    const int n_metrics = P.cfg.count();
    for (int wi = 0; wi < P.n_weights; ++wi) {
        if (threadIdx.x < P.n_sys) s_w[threadIdx.x] = P.weights[(size_t)wi * P.n_sys + threadIdx.x];
        if (threadIdx.x < kMaxGold) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        if (threadIdx.x < n_gold_u) {
            const int slot = s_slot[threadIdx.x];
            if (slot >= 0) {
                const uint32_t m = h_meta[slot];
                s_gscore[threadIdx.x] = fused_score<F32>(t32, t64, m, slot, s_w, P.n_sys);
                s_gorder[threadIdx.x] = m & 0x0fffffffu;
            }
        }
        __syncthreads();
        // how many union entries come before each gold doc in the reference's stable descending sort
        for (int g0 = 0; g0 < n_gold_u; g0 += 8) {
            int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < H; i += blockDim.x) {
                const uint32_t m = h_meta[i];
                if (!(m >> 28)) continue;
                const double sc = fused_score<F32>(t32, t64, m, i, s_w, P.n_sys);
                const uint32_t ord = m & 0x0fffffffu;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = g0 + u;
                    if (j < n_gold_u && s_slot[j] >= 0)
                        c[u] += (sc > s_gscore[j]) || (sc == s_gscore[j] && ord < s_gorder[j]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int tot = warp_sum(c[u]);
                if ((threadIdx.x & 31) == 0 && tot && g0 + u < n_gold_u) atomicAdd(&s_cnt[g0 + u], tot);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int j = 0; j < n_gold_u; ++j) s_rank[j] = s_slot[j] >= 0 ? s_cnt[j] + 1 : kNoRank;
            sort_ranks(s_rank, n_gold_u);
            int nf = 0;
            while (nf < n_gold_u && s_rank[nf] != kNoRank) ++nf;
            double m[4 * kMaxKs + 1];
            metrics_from_ranks(s_rank, nf, s_ngold, P.cfg, m);
            for (int i = 0; i < n_metrics; ++i) atomicAdd(&P.out_sum[(size_t)wi * n_metrics + i], m[i]);
        }
        __syncthreads();
    }
}

static int fill_cfg(MetricCfg& C, const int32_t* recall_k, int n_recall, const int32_t* map_k, int n_map, const int32_t* mrr_k,
                    int n_mrr, const int32_t* ndcg_k, int n_ndcg) {
    FZ_REQUIRE(n_recall >= 0 && n_recall <= kMaxKs && n_map >= 0 && n_map <= kMaxKs && n_mrr >= 0 && n_mrr <= kMaxKs &&
                   n_ndcg >= 0 && n_ndcg <= kMaxKs, "at most %d cut-offs per metric", kMaxKs);
    memset(&C, 0, sizeof(C));
    C.n_recall = n_recall; C.n_map = n_map; C.n_mrr = n_mrr; C.n_ndcg = n_ndcg;
    for (int i = 0; i < n_recall; ++i) C.recall_k[i] = recall_k[i];
    for (int i = 0; i < n_map; ++i) C.map_k[i] = map_k[i];
    for (int i = 0; i < n_mrr; ++i) C.mrr_k[i] = mrr_k[i];
    for (int i = 0; i < n_ndcg; ++i) C.ndcg_k[i] = ndcg_k[i];
    return FZ_OK;
}

}  // namespace fz

using namespace fz;

extern "C" {

int fz_rank_metrics(const int32_t* ids, const int32_t* lens, int n_queries, int stride, const int32_t* gold_ptr,
                    const int32_t* gold_ids, const int32_t* recall_k_h, int n_recall, const int32_t* map_k_h, int n_map,
                    const int32_t* mrr_k_h, int n_mrr, const int32_t* ndcg_k_h, int n_ndcg, double* out_sum,
                    double* ws_per_query, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(ids && gold_ptr && gold_ids && out_sum, "null pointer");
    FZ_REQUIRE(stride >= 1 && n_queries >= 0, "bad sizes");
    MetricCfg C;
    int rc = fill_cfg(C, recall_k_h, n_recall, map_k_h, n_map, mrr_k_h, n_mrr, ndcg_k_h, n_ndcg);
    if (rc) return rc;
    FZ_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double) * C.count(), stream));
    if (n_queries == 0) return FZ_OK;
    rank_metrics_kernel<<<n_queries, kMetricThreads, 0, stream>>>(ids, lens, stride, gold_ptr, gold_ids, C, out_sum, ws_per_query);
    if (ws_per_query) metrics_reduce_kernel<<<1, 64, 0, stream>>>(ws_per_query, n_queries, C.count(), out_sum);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_fuse_sweep(const int32_t* const* ids_h, const double* const* vals_h, const int32_t* const* lens_h,
                  const int32_t* list_stride_h, int n_sys, int n_queries, int values_are_f32, const double* weights,
                  int n_weights, const int32_t* gold_ptr, const int32_t* gold_ids, const int32_t* recall_k_h, int n_recall,
                  const int32_t* map_k_h, int n_map, const int32_t* mrr_k_h, int n_mrr, const int32_t* ndcg_k_h, int n_ndcg,
                  double* out_sum, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(ids_h && vals_h && lens_h && list_stride_h && weights && gold_ptr && gold_ids && out_sum, "null pointer");
    FZ_REQUIRE(n_sys >= 1 && n_sys <= kSweepMaxSys, "1 .. %d systems", kSweepMaxSys);
    FZ_REQUIRE(n_weights >= 1 && n_queries >= 0, "bad sizes");
    SweepParams P;
    memset(&P, 0, sizeof(P));
    int rc = fill_cfg(P.cfg, recall_k_h, n_recall, map_k_h, n_map, mrr_k_h, n_mrr, ndcg_k_h, n_ndcg);
    if (rc) return rc;
    long long total = 0;
    for (int s = 0; s < n_sys; ++s) {
        FZ_REQUIRE(ids_h[s] && vals_h[s] && list_stride_h[s] >= 1, "system %d: null list", s);
        P.ids[s] = ids_h[s];
        P.val[s] = vals_h[s];
        P.lens[s] = lens_h[s];
        P.stride[s] = list_stride_h[s];
        total += list_stride_h[s];
    }
    FZ_REQUIRE(total < (1 << 27), "lists too long");
    int H = 64;
    while (H < 2 * total) H <<= 1;
    const size_t smem = (size_t)H * 8 + (size_t)H * n_sys * (values_are_f32 ? 4 : 8);
    FZ_REQUIRE(smem <= 212 * 1024, "the union of the lists (%lld entries, %d systems) does not fit shared memory: sweep top-k "
               "lists, not full rankings", total, n_sys);
    P.n_sys = n_sys;
    P.n_queries = n_queries;
    P.n_weights = n_weights;
    P.table_slots = H;
    P.f32_path = values_are_f32;
    P.weights = weights;
    P.gold_ptr = gold_ptr;
    P.gold_ids = gold_ids;
    P.out_sum = out_sum;
    FZ_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double) * (size_t)n_weights * P.cfg.count(), stream));
    if (n_queries == 0) return FZ_OK;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(fuse_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(fuse_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        attr = true;
    }
    ProfScope prof("fuse_sweep", stream);
    if (values_are_f32) fuse_sweep_kernel<true><<<n_queries, kSweepThreads, smem, stream>>>(P);
    else fuse_sweep_kernel<false><<<n_queries, kSweepThreads, smem, stream>>>(P);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
