// K4: rank fusion as one segmented kernel — one CTA per query walks the systems' lists, normalises, weights,
// union-sums through an open-addressing table and sorts the union.
//
// Semantics follow src/retrievers/hybrid.py (Aggregator, :166-307) including its quirks:
//   * a repeated id inside one list keeps its first position and its last score (convert2dict :223-233)
//   * borda (n-idx+1)/n and reciprocal rank 1/(60+idx+1) in fp64 (:247-252)
//   * min-max / z-score (unbiased std) / arctan / percentile-rank / NCE in fp32 (:254-278), weights in fp32 (:291)
//   * normalization 'none' leaves the fp64 scores untouched and weights them in fp64 (:280, :291)
//   * union-sum in system order, stable descending sort => ties by first insertion (:294-307)
// The table lives in shared memory when it fits (4 systems x 1000 candidates) and in a global workspace otherwise
// (the reference's full-length lists, n = N).
#include "common.cuh"
#include "radix_select.cuh"

#include <limits>

namespace fz {

constexpr int kFuseThreads = 512;
constexpr int kEmptyKey = (int)0x80000000;
constexpr int kSortChunk = 4096;

struct FuseParams {
    const int32_t* ids[FZ_FUSE_MAX_SYSTEMS];
    const void* scores[FZ_FUSE_MAX_SYSTEMS];
    const int32_t* lens[FZ_FUSE_MAX_SYSTEMS];
    const float* distr[FZ_FUSE_MAX_SYSTEMS];
    int32_t distr_len[FZ_FUSE_MAX_SYSTEMS];
    int32_t is_f64[FZ_FUSE_MAX_SYSTEMS];
    int32_t stride[FZ_FUSE_MAX_SYSTEMS];
    double weight[FZ_FUSE_MAX_SYSTEMS];
    int n_sys, n_queries, method, norm;
    int keep_order;    // FZ_FUSE_KEEP_ORDER: one system, output in first-insertion order instead of sorted
    int promote64;     // FZ_FUSE_PROMOTE_F64: fp32 normalised scores are weighted and summed in fp64 (NumPy 1.x promotion)
    int table_slots;   // power of two
    int max_len;       // longest list
    int use_smem;
    size_t ws_per_query;
    unsigned char* ws;
    int32_t* out_ids;
    double* out_scores;
    int32_t* out_len;
    int out_stride;
};

// footprint of one query's scratch: table (key, order, first, last, acc) + per-entry (val, slot)
__host__ __device__ inline size_t fuse_footprint(int slots, int max_len) {
    return (size_t)slots * 24 + (size_t)max_len * 12 + 64;
}

__device__ __forceinline__ int hash_slot(int key, int mask) {
    uint32_t h = (uint32_t)key * 2654435761u;
    return (int)((h >> 7) & (uint32_t)mask);
}

// block-wide reductions over per-thread partials (kFuseThreads threads)
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < kFuseThreads / 32; ++w) t += red[w];
    return t;
}
__device__ float block_minmax(float v, bool is_max, double* red) {
    for (int o = 16; o > 0; o >>= 1) {
        float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : fminf(v, u);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)v;
    __syncthreads();
    float t = (float)red[0];
    for (int w = 1; w < kFuseThreads / 32; ++w) t = is_max ? fmaxf(t, (float)red[w]) : fminf(t, (float)red[w]);
    return t;
}

// index of the first minimum of |d_i - s| over an ascending fp32 distribution, fp32 arithmetic
// (torch.argmin(torch.abs(distribution[:, None] - scores), axis=0), hybrid.py:274-275)
__device__ int percentile_index(const float* __restrict__ d, int L, float s) {
    int lo = 0, hi = L;                       // first j with d[j] >= s
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (d[mid] < s) lo = mid + 1; else hi = mid;
    }
    const int j = lo;
    float below = j > 0 ? fabsf(__fsub_rn(d[j - 1], s)) : INFINITY;
    float above = j < L ? fabsf(__fsub_rn(d[j], s)) : INFINITY;
    if (below <= above) {
        // |d_i - s| is non-increasing for i < j: first index whose distance equals `below`
        int a = 0, b = j - 1;
        while (a < b) {
            int mid = (a + b) >> 1;
            if (fabsf(__fsub_rn(d[mid], s)) <= below) b = mid; else a = mid + 1;
        }
        return a;
    }
    return j;
}

__global__ void __launch_bounds__(kFuseThreads) fuse_kernel(const FuseParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[kFuseThreads / 32];
    __shared__ int s_scan[kFuseThreads / 32];
    __shared__ int s_base, s_union;

    const int q = blockIdx.x;
    const int H = P.table_slots, mask = H - 1;
    unsigned char* mem = P.use_smem ? smem_raw : P.ws + (size_t)q * P.ws_per_query;
    // layout: acc[H] f64 | first[H] u32 | last[H] u32 | key[H] i32 | order[H] u32 | val[max_len] f64 | slot[max_len] i32
    double* h_acc = reinterpret_cast<double*>(mem);
    uint32_t* h_first = reinterpret_cast<uint32_t*>(mem + (size_t)H * 8);
    uint32_t* h_last = h_first + H;
    int* h_key = reinterpret_cast<int*>(h_last + H);
    uint32_t* h_order = reinterpret_cast<uint32_t*>(h_key + H);
    double* e_val = reinterpret_cast<double*>(h_order + H);
    int* e_slot = reinterpret_cast<int*>(e_val + P.max_len);
    Entry* sort_buf = reinterpret_cast<Entry*>(h_first);   // reuses first/last once the union is complete

    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        h_key[i] = kEmptyKey;
        h_order[i] = 0xffffffffu;
        h_first[i] = 0xffffffffu;
        h_last[i] = 0u;
        h_acc[i] = 0.0;
    }
    if (threadIdx.x == 0) { s_base = 0; s_union = 0; }
    __syncthreads();

    const bool f32_path = P.method == FZ_FUSE_NSF && P.norm != FZ_NORM_NONE;

    for (int s = 0; s < P.n_sys; ++s) {
        const int n_raw = P.lens[s] ? min(P.lens[s][q], P.stride[s]) : P.stride[s];
        const int32_t* ids = P.ids[s] + (size_t)q * P.stride[s];
        const float* sc32 = P.is_f64[s] ? nullptr : reinterpret_cast<const float*>(P.scores[s]) + (size_t)q * P.stride[s];
        const double* sc64 = P.is_f64[s] ? reinterpret_cast<const double*>(P.scores[s]) + (size_t)q * P.stride[s] : nullptr;

        // pass 1: find-or-claim the slot of every entry; remember first / last position of each id in this list
        for (int p = threadIdx.x; p < n_raw; p += blockDim.x) {
            const int key = ids[p];
            int slot = hash_slot(key, mask);
            while (true) {
                int cur = h_key[slot];
                if (cur == key) break;
                if (cur == kEmptyKey) {
                    int old = atomicCAS(&h_key[slot], kEmptyKey, key);
                    if (old == kEmptyKey || old == key) break;
                }
                slot = (slot + 1) & mask;
            }
            e_slot[p] = slot;
            atomicMin(&h_first[slot], (uint32_t)p);
            atomicMax(&h_last[slot], (uint32_t)p);
        }
        __syncthreads();

        // pass 2: primaries (first occurrence) get their rank among primaries; their score is the last occurrence's
        int base_rank = 0;    // running count of primaries before the current chunk
        double sum = 0.0;
        float vmin = INFINITY, vmax = -INFINITY;
        const int n_chunks = (n_raw + blockDim.x - 1) / blockDim.x;
        for (int c = 0; c < n_chunks; ++c) {
            const int p = c * blockDim.x + threadIdx.x;
            int prim = 0;
            double eff = 0.0;
            if (p < n_raw) {
                const int slot = e_slot[p];
                prim = h_first[slot] == (uint32_t)p;
                if (prim) {
                    const uint32_t lp = h_last[slot];
                    eff = sc64 ? sc64[lp] : (double)sc32[lp];
                }
            }
            // exclusive scan of prim over the chunk
            unsigned bal = __ballot_sync(0xffffffffu, prim);
            int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            int in_warp = __popc(bal & ((1u << lane) - 1));
            __syncthreads();
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < kFuseThreads / 32; ++w) {
                int v = s_scan[w];
                if (w < warp) before += v;
                total += v;
            }
            if (p < n_raw) {
                if (prim) {
                    const int rank = base_rank + before + in_warp;
                    e_val[p] = eff;
                    h_last[e_slot[p]] = (uint32_t)rank;              // reuse `last` as the rank of this id in list s
                    if (f32_path) {
                        float f = (float)eff;
                        sum += (double)f;
                        vmin = fminf(vmin, f);
                        vmax = fmaxf(vmax, f);
                    }
                } else {
                    e_slot[p] = -1;
                }
            }
            base_rank += total;
        }
        const int n = base_rank;     // len(dict) of this list
        __syncthreads();

        float mean_f = 0.f, std_f = 0.f, min_f = 0.f, max_f = 0.f;
        if (f32_path && (P.norm == FZ_NORM_MINMAX || P.norm == FZ_NORM_ZSCORE)) {
            if (P.norm == FZ_NORM_MINMAX) {
                min_f = block_minmax(vmin, false, red);
                max_f = block_minmax(vmax, true, red);
            } else {
                const double tot = block_sum(sum, red);
                const double mean_d = n > 0 ? tot / n : 0.0;
                mean_f = (float)mean_d;
                double ss = 0.0;
                for (int p = threadIdx.x; p < n_raw; p += blockDim.x)
                    if (e_slot[p] >= 0) {
                        double dlt = (double)(float)e_val[p] - mean_d;
                        ss += dlt * dlt;
                    }
                ss = block_sum(ss, red);
                std_f = n > 1 ? (float)sqrt(ss / (n - 1)) : __int_as_float(0x7fc00000);   // torch.std(n=1) = nan
            }
        }

        // pass 3: transformed + weighted value of every primary, accumulated into the union in system order
        const double w64 = P.weight[s];
        const float w32 = (float)w64;
        const int base_order = s_base;
        for (int p = threadIdx.x; p < n_raw; p += blockDim.x) {
            const int slot = e_slot[p];
            if (slot < 0) continue;
            const int rank = (int)h_last[slot];
            double v;
            if (P.method == FZ_FUSE_BCF) {
                v = __ddiv_rn((double)(n - rank + 1), (double)n);
            } else if (P.method == FZ_FUSE_RRF) {
                v = __ddiv_rn(1.0, (double)(60 + rank + 1));
            } else if (P.norm == FZ_NORM_NONE) {
                v = __dmul_rn(e_val[p], w64);
            } else {
                float x = (float)e_val[p], t;
                switch (P.norm) {
                    case FZ_NORM_MINMAX:
                        t = (min_f != max_f) ? __fdiv_rn(__fsub_rn(x, min_f), __fsub_rn(max_f, min_f)) : 1.0f;
                        break;
                    case FZ_NORM_ZSCORE:
                        t = (std_f != 0.0f) ? __fdiv_rn(__fsub_rn(x, mean_f), std_f) : 0.0f;
                        break;
                    case FZ_NORM_ARCTAN:
                        t = __fmul_rn((float)(2.0 / 3.14159265358979323846), atanf(__fmul_rn(0.1f, x)));
                        break;
                    case FZ_NORM_IDENTITY_F32:
                        t = x;
                        break;
                    default: {
                        const int L = P.distr_len[s];
                        const int idx = percentile_index(P.distr[s], L, x);
                        t = __fdiv_rn((float)idx, (float)L);
                        if (P.norm == FZ_NORM_NCE) {
                            float u = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(t, 100.0f)), 1.0f);
                            float z = __fmul_rn(erfinvf(u), (float)1.4142135623730951);
                            t = __fadd_rn(__fmul_rn(z, 21.06f), 50.0f);
                        }
                    }
                }
                v = P.promote64 ? __dmul_rn((double)t, w64) : (double)__fmul_rn(t, w32);
            }
            if (h_order[slot] == 0xffffffffu) {
                h_order[slot] = (uint32_t)(base_order + rank);
                atomicAdd(&s_union, 1);
            }
            if (f32_path && !P.promote64)
                h_acc[slot] = (double)__fadd_rn((float)h_acc[slot], (float)v);
            else
                h_acc[slot] = __dadd_rn(h_acc[slot], v);
            h_first[slot] = 0xffffffffu;      // reset the per-list temporaries
            h_last[slot] = 0u;
        }
        __syncthreads();
        // non-primary duplicates also touched first/last of their slot: already reset through the primary
        if (threadIdx.x == 0) s_base = base_order + n;
        __syncthreads();
    }

    const int U = s_union;
    if (P.keep_order) {
        // one system: first-insertion order is dense 0 .. U-1, so the (deduplicated, transformed) list is written in place
        for (int i = threadIdx.x; i < P.out_stride; i += blockDim.x) {
            P.out_ids[(size_t)q * P.out_stride + i] = -1;
            P.out_scores[(size_t)q * P.out_stride + i] = -std::numeric_limits<double>::infinity();
        }
        __syncthreads();
        for (int i = threadIdx.x; i < H; i += blockDim.x) {
            if (h_key[i] != kEmptyKey && h_order[i] != 0xffffffffu && (int)h_order[i] < P.out_stride) {
                P.out_ids[(size_t)q * P.out_stride + h_order[i]] = h_key[i];
                P.out_scores[(size_t)q * P.out_stride + h_order[i]] = h_acc[i];
            }
        }
        if (threadIdx.x == 0) P.out_len[q] = min(U, P.out_stride);
        return;
    }
    // gather the union into sortable records (over the dead first/last arrays), sort, write out
    int U2 = 1;
    while (U2 < U) U2 <<= 1;
    __shared__ int s_fill;
    if (threadIdx.x == 0) s_fill = 0;
    __syncthreads();
    // two-step: read slot fields into registers first (sort_buf aliases first/last, not key/order/acc)
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        if (h_key[i] != kEmptyKey && h_order[i] != 0xffffffffu) {
            int pos = atomicAdd(&s_fill, 1);
            Entry e;
            e.skey = ord64(h_acc[i] + 0.0);      // -0.0 ties with +0.0 like Python's sorted()
            e.tie = ~h_order[i];
            e.payload = (uint32_t)h_key[i];
            sort_buf[pos] = e;
        }
    }
    __syncthreads();
    // Only the best out_stride entries are written: when that is less than the union, pick them with a radix select over
    // (fused score, first insertion) and sort just those - the sort network is the shared-memory-bound part of the kernel.
    if (P.use_smem && P.out_stride < U && (size_t)P.out_stride * sizeof(Entry) <= (size_t)H * 8) {
        __shared__ int s_hist[256];
        __shared__ int s_bc[4];
        __shared__ int s_nwin;
        uint64_t kth_hi = 0;
        uint32_t kth_lo = 0;
        const Entry* src = sort_buf;
        cta_radix_select_kth_by<uint64_t>([src](int i) { return src[i].skey; }, [src](int i) { return src[i].tie; }, U,
                                          P.out_stride, s_hist, s_bc, kth_hi, kth_lo);
        Entry* win = reinterpret_cast<Entry*>(h_acc);     // the accumulators are dead once the records are gathered
        if (threadIdx.x == 0) s_nwin = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < U; i += blockDim.x) {
            const Entry e = sort_buf[i];
            if (key_ge<uint64_t>(e.skey, e.tie, kth_hi, kth_lo)) win[atomicAdd(&s_nwin, 1)] = e;
        }
        __syncthreads();
        const int m = s_nwin;                             // == out_stride (keys are unique)
        int m2 = 1;
        while (m2 < m) m2 <<= 1;
        for (int i = m + threadIdx.x; i < m2; i += blockDim.x) {
            Entry e;
            e.skey = 0; e.tie = 0; e.payload = 0;
            win[i] = e;
        }
        __syncthreads();
        bitonic_sort_cta(win, m2);
        for (int i = threadIdx.x; i < P.out_stride; i += blockDim.x) {
            P.out_ids[(size_t)q * P.out_stride + i] = i < m ? (int32_t)win[i].payload : -1;
            P.out_scores[(size_t)q * P.out_stride + i] = i < m ? unord64(win[i].skey) : -std::numeric_limits<double>::infinity();
        }
        if (threadIdx.x == 0) P.out_len[q] = m;
        return;
    }
    for (int i = U + threadIdx.x; i < U2; i += blockDim.x) {
        Entry e;
        e.skey = 0; e.tie = 0; e.payload = 0;
        sort_buf[i] = e;
    }
    __syncthreads();
    if (P.use_smem || U2 <= 1) {
        bitonic_sort_cta(sort_buf, U2);
    } else {
        // global scratch: stage chunks through the dynamic shared memory
        Entry* sm = reinterpret_cast<Entry*>(smem_raw);
        if (U2 <= kSortChunk) {
            for (int i = threadIdx.x; i < U2; i += blockDim.x) sm[i] = sort_buf[i];
            __syncthreads();
            bitonic_sort_cta(sm, U2);
            for (int i = threadIdx.x; i < U2; i += blockDim.x) sort_buf[i] = sm[i];
            __syncthreads();
        } else {
            // same schedule as select.cu::bitonic_sort_large
            for (int base = 0; base < U2; base += kSortChunk) {
                for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sm[i] = sort_buf[base + i];
                __syncthreads();
                for (int k = 2; k <= kSortChunk; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) {
                            int p = i ^ j;
                            if (p > i) {
                                Entry x = sm[i], y = sm[p];
                                bool up = ((base + i) & k) == 0;
                                if (entry_before(y, x) == up) { sm[i] = y; sm[p] = x; }
                            }
                        }
                        __syncthreads();
                    }
                for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sort_buf[base + i] = sm[i];
                __syncthreads();
            }
            for (int k = 2 * kSortChunk; k <= U2; k <<= 1) {
                for (int j = k >> 1; j >= kSortChunk; j >>= 1) {
                    for (int i = threadIdx.x; i < U2; i += blockDim.x) {
                        int p = i ^ j;
                        if (p > i) {
                            Entry x = sort_buf[i], y = sort_buf[p];
                            bool up = (i & k) == 0;
                            if (entry_before(y, x) == up) { sort_buf[i] = y; sort_buf[p] = x; }
                        }
                    }
                    __syncthreads();
                }
                for (int base = 0; base < U2; base += kSortChunk) {
                    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sm[i] = sort_buf[base + i];
                    __syncthreads();
                    for (int j = kSortChunk >> 1; j > 0; j >>= 1) {
                        for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) {
                            int p = i ^ j;
                            if (p > i) {
                                Entry x = sm[i], y = sm[p];
                                bool up = ((base + i) & k) == 0;
                                if (entry_before(y, x) == up) { sm[i] = y; sm[p] = x; }
                            }
                        }
                        __syncthreads();
                    }
                    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sort_buf[base + i] = sm[i];
                    __syncthreads();
                }
            }
        }
    }
    const int n_out = min(U, P.out_stride);
    for (int i = threadIdx.x; i < P.out_stride; i += blockDim.x) {
        int32_t id = -1;
        double sc = -std::numeric_limits<double>::infinity();
        if (i < n_out) {
            id = (int32_t)sort_buf[i].payload;
            sc = unord64(sort_buf[i].skey);
        }
        P.out_ids[(size_t)q * P.out_stride + i] = id;
        P.out_scores[(size_t)q * P.out_stride + i] = sc;
    }
    if (threadIdx.x == 0) P.out_len[q] = n_out;
}

static void fuse_plan(int n_sys, const int32_t* list_stride_h, int* slots, int* max_len, size_t* per_query) {
    long long total = 0;
    int ml = 1;
    for (int s = 0; s < n_sys; ++s) {
        total += list_stride_h[s];
        ml = ml > list_stride_h[s] ? ml : list_stride_h[s];
    }
    int h = 64;
    while (h < 2 * total) h <<= 1;
    *slots = h;
    *max_len = ml;
    *per_query = align_up(fuse_footprint(h, ml), 256);
}

constexpr size_t kFuseSmemLimit = 216 * 1024;

}  // namespace fz

using namespace fz;

extern "C" {

size_t fz_fuse_workspace_bytes(int n_sys, int n_queries, const int32_t* list_stride_h) {
    if (n_sys < 1 || n_sys > FZ_FUSE_MAX_SYSTEMS || !list_stride_h) return 0;
    int slots, max_len;
    size_t per_query;
    fuse_plan(n_sys, list_stride_h, &slots, &max_len, &per_query);
    if (per_query <= kFuseSmemLimit) return 0;
    return per_query * (size_t)n_queries;
}

int fz_fuse(const int32_t* const* ids_h, const void* const* scores_h, const int32_t* const* lens_h,
            const int32_t* score_is_f64_h, const int32_t* list_stride_h, int n_sys, int n_queries, int method,
            int normalization, const double* weights_h, const float* const* distr_h, const int32_t* distr_len_h,
            int32_t* out_ids, double* out_scores, int32_t* out_len, int out_stride, void* ws, size_t ws_bytes,
            fz_stream_t stream) {
    FZ_REQUIRE(n_sys >= 1 && n_sys <= FZ_FUSE_MAX_SYSTEMS, "n_sys=%d must be in [1,%d]", n_sys, FZ_FUSE_MAX_SYSTEMS);
    FZ_REQUIRE(ids_h && scores_h && score_is_f64_h && list_stride_h && out_ids && out_scores && out_len, "null pointer");
    const int keep_order = (method & FZ_FUSE_KEEP_ORDER) != 0;
    const int promote64 = (method & FZ_FUSE_PROMOTE_F64) != 0;
    method &= ~(FZ_FUSE_KEEP_ORDER | FZ_FUSE_PROMOTE_F64);
    FZ_REQUIRE(method == FZ_FUSE_BCF || method == FZ_FUSE_RRF || method == FZ_FUSE_NSF, "unknown fusion method %d", method);
    FZ_REQUIRE(!keep_order || n_sys == 1, "FZ_FUSE_KEEP_ORDER needs exactly one system");
    FZ_REQUIRE(out_stride >= 1, "out_stride must be positive");
    if (method == FZ_FUSE_NSF) {
        FZ_REQUIRE(normalization >= FZ_NORM_NONE && normalization <= FZ_NORM_IDENTITY_F32, "unknown normalization %d", normalization);
        FZ_REQUIRE(weights_h, "nsf needs linear weights");
        if (normalization == FZ_NORM_PERCENTILE || normalization == FZ_NORM_NCE) {
            FZ_REQUIRE(distr_h && distr_len_h, "percentile normalisation needs distributions");
            for (int s = 0; s < n_sys; ++s)
                FZ_REQUIRE(distr_h[s] && distr_len_h[s] >= 1, "missing percentile distribution for system %d", s);
        }
    }
    if (n_queries == 0) return FZ_OK;
    FuseParams P;
    memset(&P, 0, sizeof(P));
    for (int s = 0; s < n_sys; ++s) {
        FZ_REQUIRE(ids_h[s] && scores_h[s] && list_stride_h[s] >= 1, "system %d: null list or empty stride", s);
        P.ids[s] = ids_h[s];
        P.scores[s] = scores_h[s];
        P.lens[s] = lens_h ? lens_h[s] : nullptr;
        P.is_f64[s] = score_is_f64_h[s];
        P.stride[s] = list_stride_h[s];
        P.weight[s] = weights_h ? weights_h[s] : 1.0;
        P.distr[s] = distr_h ? distr_h[s] : nullptr;
        P.distr_len[s] = distr_len_h ? distr_len_h[s] : 0;
    }
    P.n_sys = n_sys;
    P.n_queries = n_queries;
    P.method = method;
    P.keep_order = keep_order;
    P.promote64 = promote64;
    P.norm = normalization;
    fuse_plan(n_sys, list_stride_h, &P.table_slots, &P.max_len, &P.ws_per_query);
    P.use_smem = P.ws_per_query <= kFuseSmemLimit;
    P.ws = (unsigned char*)ws;
    P.out_ids = out_ids;
    P.out_scores = out_scores;
    P.out_len = out_len;
    P.out_stride = out_stride;
    size_t smem = P.use_smem ? P.ws_per_query : (size_t)kSortChunk * sizeof(Entry);
    if (!P.use_smem) {
        size_t need = P.ws_per_query * (size_t)n_queries;
        FZ_REQUIRE(ws && ws_bytes >= need, "fuse workspace too small: %zu < %zu", ws_bytes, need);
    }
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFuseSmemLimit));
        attr = true;
    }
    ProfScope prof("fuse", (cudaStream_t)stream);
    fuse_kernel<<<n_queries, kFuseThreads, smem, (cudaStream_t)stream>>>(P);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
