// Segmented sort / top-k selection / k-way merge (K5) and the candidate-buffer maintenance used by K1/K2.
//
// One CTA per query.  Records are 16-byte `Entry`s (order-preserving score key, tie key, payload) sorted
// "best first" with a bitonic network: in shared memory when the segment fits (<= 8192 records), otherwise in
// a global workspace with the short-stride stages done on 4096-record chunks staged through shared memory.
#include "common.cuh"
#include "topk_state.cuh"
#include "radix_select.cuh"

#include <cstdarg>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace fz {

// ------------------------------------------------------------------------------------------- error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------- profiler
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

void prof_begin(const char* name, cudaStream_t stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream);
    g_prof.push_back(r);
}
void prof_end(cudaStream_t stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, stream);
}

constexpr int kSortThreads = 1024;
// cand_select CTA size: fp32 buffers (64 KB of keys) fit two 512-thread CTAs per SM, which overlap each other's barriers;
// fp64 buffers (96 KB) fit one, which then wants all 1024 threads
template <typename ST> struct SelectCfg { static constexpr int threads = sizeof(ST) == 4 ? 512 : 1024, blocks = sizeof(ST) == 4 ? 2 : 1; };
constexpr int kSmemEntries = 8192;   // 128 KB
constexpr int kChunk = 4096;         // chunk staged through smem by the large sort

static inline int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// compare-exchange stages j = j_hi .. 1 of merge size k on `cnt` records in smem whose global index starts at base
__device__ __forceinline__ void smem_stages(Entry* sm, int cnt, int base, int k, int j_hi) {
    for (int j = j_hi; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            int p = i ^ j;
            if (p > i) {
                Entry x = sm[i], y = sm[p];
                bool up = ((base + i) & k) == 0;
                if (entry_before(y, x) == up) {
                    sm[i] = y;
                    sm[p] = x;
                }
            }
        }
        __syncthreads();
    }
}

// Sort n_pow2 (> kChunk) records in global memory `a`; `sm` holds kChunk records.
__device__ void bitonic_sort_large(Entry* a, int n_pow2, Entry* sm) {
    for (int base = 0; base < n_pow2; base += kChunk) {
        for (int i = threadIdx.x; i < kChunk; i += blockDim.x) sm[i] = a[base + i];
        __syncthreads();
        for (int k = 2; k <= kChunk; k <<= 1) smem_stages(sm, kChunk, base, k, k >> 1);
        for (int i = threadIdx.x; i < kChunk; i += blockDim.x) a[base + i] = sm[i];
        __syncthreads();
    }
    for (int k = 2 * kChunk; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j >= kChunk; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    Entry x = a[i], y = a[p];
                    bool up = (i & k) == 0;
                    if (entry_before(y, x) == up) {
                        a[i] = y;
                        a[p] = x;
                    }
                }
            }
            __syncthreads();
        }
        for (int base = 0; base < n_pow2; base += kChunk) {
            for (int i = threadIdx.x; i < kChunk; i += blockDim.x) sm[i] = a[base + i];
            __syncthreads();
            smem_stages(sm, kChunk, base, k, kChunk >> 1);
            for (int i = threadIdx.x; i < kChunk; i += blockDim.x) a[base + i] = sm[i];
            __syncthreads();
        }
    }
}

__device__ __forceinline__ size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ Entry pad_entry() {
    Entry e;
    e.skey = 0;
    e.tie = 0;
    e.payload = 0xffffffffu;
    return e;
}

// --------------------------------------------------------------------------------- candidate buffers
template <typename ST>
__global__ void cand_init_kernel(CandState<ST> st, int n_queries) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_queries) {
        st.cnt[q] = 0;
        st.npos[q] = 0;
        st.nneg[q] = 0;
        st.tau[q] = -std::numeric_limits<ST>::infinity();
        st.status[q] = 0;
    }
}

// One CTA per query.  The buffer's records are staged in shared memory as composite keys (order-preserving
// score bits : ~doc id); a radix select finds the k-th largest key, everything that survives the new threshold is
// written back compacted (unsorted - only the final round sorts, and then just the k winners).
template <typename ST> struct HiOf;
template <> struct HiOf<float> { using type = uint32_t; };
template <> struct HiOf<double> { using type = uint64_t; };
__device__ __forceinline__ void hi_score(uint32_t h, float& s) { s = unord32(h); }
__device__ __forceinline__ void hi_score(uint64_t h, double& s) { s = unord64(h); }

template <typename ST>
__global__ void __launch_bounds__(SelectCfg<ST>::threads, SelectCfg<ST>::blocks) cand_select_kernel(CandState<ST> st, int k, ST margin, int final_out,
                                                                   long long doc_base, ST* out_scores,
                                                                   int32_t* out_ids, int32_t* out_n, int k_pow2,
                                                                   const ST* __restrict__ floor) {
    using HiT = typename HiOf<ST>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HiT* s_hi = reinterpret_cast<HiT*>(smem_raw);                               // [cap]
    uint32_t* s_lo = reinterpret_cast<uint32_t*>(s_hi + st.cap);                // [cap]
    Entry* s_sort = reinterpret_cast<Entry*>(smem_raw + align_up_dev((size_t)st.cap * (sizeof(HiT) + 4), 16));   // [k_pow2]
    __shared__ int s_hist[256];
    __shared__ int s_bcast[4];
    __shared__ int s_keep, s_win;

    const int q = blockIdx.x;
    const int raw = st.cnt[q];
    const int n = min(raw, st.cap);
    if (threadIdx.x == 0) {
        s_keep = 0;
        s_win = 0;
        if (raw > st.cap) st.status[q] |= FZ_STATUS_OVERFLOW;
    }
    const size_t off = (size_t)q * st.cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_hi[i] = (HiT)score_key(st.score[off + i]);
        s_lo[i] = ~(uint32_t)st.id[off + i];
    }
    __syncthreads();

    ST tau = st.tau[q];
    HiT kth_hi = 0;
    uint32_t kth_lo = 0;
    const bool full = n >= k;
    if (full) {
        cta_radix_select_kth<HiT>(s_hi, s_lo, n, k, s_hist, s_bcast, kth_hi, kth_lo);
        ST kth;
        hi_score(kth_hi, kth);
        const ST t = kth - margin;
        if (t > tau) tau = t;
    }
    // Cross-shard floor (sharded corpus): a score f such that at least k documents over ALL shards reach it.  Records
    // below f can never enter the global top-k; records that TIE f must stay (the doc that ties may have a lower id than
    // the k counted ones, which live on other shards), so the stored threshold is the next value below f: the scoring
    // kernels emit score > tau.
    HiT floor_hi = 0;
    if (floor) {
        const ST f = floor[q];
        if (f > -std::numeric_limits<ST>::infinity()) {
            floor_hi = (HiT)score_key(f);
            ST below;
            hi_score((HiT)(floor_hi - 1), below);
            if (below > tau) tau = below;
        }
    }
    // survivors: exactly the k best when margin == 0 (a later doc that only ties the k-th score loses the tie:
    // rounds visit docs in ascending id order), everything with score >= tau otherwise; all records if n < k
    const HiT tau_hi = (HiT)score_key(tau);
    // compaction: one shared-memory atomic per warp (ballot + prefix), not one per record on a single counter
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        HiT h = 0;
        uint32_t l = 0;
        bool win = false, keep = false;
        if (i < n) {
            h = s_hi[i];
            l = s_lo[i];
            win = !full || key_ge<HiT>(h, l, kth_hi, kth_lo);
            keep = (!full || (margin == (ST)0 ? win : h >= tau_hi)) && h >= floor_hi;
        }
        const unsigned bk = __ballot_sync(0xffffffffu, keep);
        const unsigned bw = __ballot_sync(0xffffffffu, final_out && win);
        int pk = 0, pw = 0;
        if (lane == 0) {
            if (bk) pk = atomicAdd(&s_keep, __popc(bk));
            if (bw) pw = atomicAdd(&s_win, __popc(bw));
        }
        pk = __shfl_sync(0xffffffffu, pk, 0);
        pw = __shfl_sync(0xffffffffu, pw, 0);
        const unsigned below = (1u << lane) - 1;
        if (keep) {
            const int p = pk + __popc(bk & below);
            ST s;
            hi_score(h, s);
            st.score[off + p] = s;
            st.id[off + p] = (int32_t)~l;
        }
        if (final_out && win) {
            const int p = pw + __popc(bw & below);
            Entry e;
            e.skey = (uint64_t)h;
            e.tie = l;
            e.payload = ~l;
            s_sort[p] = e;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st.cnt[q] = s_keep;
        st.tau[q] = tau;
    }
    if (final_out) {
        const int m = s_win;        // min(n, k)
        int m2 = 1;
        while (m2 < m) m2 <<= 1;
        for (int i = m + threadIdx.x; i < m2; i += blockDim.x) s_sort[i] = pad_entry();
        __syncthreads();
        bitonic_sort_cta(s_sort, m2);
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            ST s = -std::numeric_limits<ST>::infinity();
            int32_t d = -1;
            if (i < m) {
                hi_score((HiT)s_sort[i].skey, s);
                d = (int32_t)(doc_base + (long long)s_sort[i].payload);
            }
            out_scores[(size_t)q * k + i] = s;
            out_ids[(size_t)q * k + i] = d;
        }
        if (out_n && threadIdx.x == 0) out_n[q] = m;
    }
}

// out[q] = (rank-th best score in query q's candidate buffer) - margin, or -inf when it holds fewer than `rank` records;
// the buffer is left untouched (cross-shard thresholds: every shard holds >= rank records at or above its value, so the
// minimum over G shards with rank = ceil(k / G) is reached by at least k documents overall)
template <typename ST>
__global__ void __launch_bounds__(512) cand_kth_kernel(CandState<ST> st, int rank, ST margin, ST* __restrict__ out) {
    using HiT = typename HiOf<ST>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HiT* s_hi = reinterpret_cast<HiT*>(smem_raw);
    uint32_t* s_lo = reinterpret_cast<uint32_t*>(s_hi + st.cap);
    __shared__ int s_hist[256];
    __shared__ int s_bcast[4];
    const int q = blockIdx.x;
    const int n = min(st.cnt[q], st.cap);
    if (n < rank) {
        if (threadIdx.x == 0) out[q] = -std::numeric_limits<ST>::infinity();
        return;
    }
    const size_t off = (size_t)q * st.cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_hi[i] = (HiT)score_key(st.score[off + i]);
        s_lo[i] = ~(uint32_t)st.id[off + i];
    }
    __syncthreads();
    HiT kth_hi = 0;
    uint32_t kth_lo = 0;
    cta_radix_select_kth<HiT>(s_hi, s_lo, n, rank, s_hist, s_bcast, kth_hi, kth_lo);
    if (threadIdx.x == 0) {
        ST v;
        hi_score(kth_hi, v);
        out[q] = v - margin;
    }
}

template <typename ST>
int cand_kth_score(const CandState<ST>& st, int n_queries, int rank, ST margin, ST* out, cudaStream_t stream) {
    using HiT = typename HiOf<ST>::type;
    static bool attr_f = false, attr_d = false;
    bool& done = std::is_same<ST, float>::value ? attr_f : attr_d;
    if (!done) {
        FZ_CUDA(cudaFuncSetAttribute(cand_kth_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        done = true;
    }
    ProfScope prof("cand_kth", stream);
    cand_kth_kernel<ST><<<n_queries, 512, (size_t)st.cap * (sizeof(HiT) + 4), stream>>>(st, rank, margin, out);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}
template int cand_kth_score<float>(const CandState<float>&, int, int, float, float*, cudaStream_t);
template int cand_kth_score<double>(const CandState<double>&, int, int, double, double*, cudaStream_t);

template <typename ST>
int cand_init(const CandState<ST>& st, int n_queries, cudaStream_t stream) {
    cand_init_kernel<ST><<<ceil_div(n_queries, 256), 256, 0, stream>>>(st, n_queries);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

template <typename ST>
int cand_select(const CandState<ST>& st, int n_queries, int k, ST margin, bool final_out, int64_t doc_base,
                ST* out_scores, int32_t* out_ids, int32_t* out_n, cudaStream_t stream, const ST* floor) {
    FZ_REQUIRE(st.cap <= kSmemEntries, "candidate capacity %d exceeds %d", st.cap, kSmemEntries);
    FZ_REQUIRE(k >= 1 && k <= st.cap, "k=%d must be in [1, cap=%d]", k, st.cap);
    using HiT = typename HiOf<ST>::type;
    const int k_pow2 = next_pow2(k);
    size_t smem = align_up((size_t)st.cap * (sizeof(HiT) + 4), 16) + (size_t)k_pow2 * sizeof(Entry);
    FZ_REQUIRE(smem <= 200 * 1024, "cap=%d / k=%d need %zu bytes of shared memory", st.cap, k, smem);
    static bool attr_f = false, attr_d = false;
    bool& done = std::is_same<ST, float>::value ? attr_f : attr_d;
    if (!done) {
        FZ_CUDA(cudaFuncSetAttribute(cand_select_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(cand_select_kernel<ST>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        done = true;
    }
    ProfScope prof(std::is_same<ST, float>::value ? "cand_select_f32" : "cand_select_f64", stream);
    cand_select_kernel<ST><<<n_queries, SelectCfg<ST>::threads, smem, stream>>>(st, k, margin, final_out ? 1 : 0,
                                                                      (long long)doc_base, out_scores, out_ids, out_n,
                                                                      k_pow2, floor);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

template int cand_init<float>(const CandState<float>&, int, cudaStream_t);
template int cand_init<double>(const CandState<double>&, int, cudaStream_t);
template int cand_select<float>(const CandState<float>&, int, int, float, bool, int64_t, float*, int32_t*, int32_t*,
                                cudaStream_t, const float*);
template int cand_select<double>(const CandState<double>&, int, int, double, bool, int64_t, double*, int32_t*,
                                 int32_t*, cudaStream_t, const double*);

// --------------------------------------------------------------------------------- merge / rank rows
// Segment q gathers `n_src` runs of `run_len` records spaced `src_stride` apart (merge), or one run of
// n records (rank rows: ids are implicit 0..n-1), sorts them and writes the best k_out.
template <typename ST>
__global__ void __launch_bounds__(kSortThreads)
    segsort_kernel(const ST* __restrict__ scores, const int32_t* __restrict__ ids, int n_src, long long src_stride,
                   int run_len, long long seg_stride, int k_out, long long id_base, ST* out_scores, int32_t* out_ids,
                   Entry* gws, int n_pow2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Entry* sm = reinterpret_cast<Entry*>(smem_raw);
    const int q = blockIdx.x;
    const int n = n_src * run_len;
    const bool in_smem = n_pow2 <= kSmemEntries;
    Entry* a = in_smem ? sm : gws + (size_t)q * n_pow2;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        Entry e = pad_entry();
        if (i < n) {
            int g = i / run_len, r = i - g * run_len;
            size_t src = (size_t)g * src_stride + (size_t)q * seg_stride + r;
            int32_t d = ids ? ids[src] : (int32_t)r;
            if (d >= 0) {
                e.skey = score_key(scores[src]);
                e.tie = ~(uint32_t)d;
                e.payload = (uint32_t)d;
            }
        }
        a[i] = e;
    }
    __syncthreads();
    if (in_smem && k_out * 4 <= n && k_out <= 2048) {
        // Only the best k_out of n are wanted (k-way merge of G top-k lists): radix-select them, then sort just those.
        __shared__ int s_hist[256];
        __shared__ int s_bc[4];
        __shared__ int s_nwin;
        __shared__ Entry s_win[2048];
        uint64_t kth_hi = 0;
        uint32_t kth_lo = 0;
        // padding / invalid records carry key (0, 0): they are unique only through `payload`, so rank them by position
        for (int i = threadIdx.x; i < n_pow2; i += blockDim.x)
            if (a[i].payload == 0xffffffffu) a[i].tie = ~(uint32_t)(0x7f000000u + i);
        __syncthreads();
        const Entry* src = a;
        const int kk = min(k_out, n);
        cta_radix_select_kth_by<uint64_t>([src](int i) { return src[i].skey; }, [src](int i) { return src[i].tie; }, n, kk, s_hist,
                                          s_bc, kth_hi, kth_lo);
        if (threadIdx.x == 0) s_nwin = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const Entry e = a[i];
            if (key_ge<uint64_t>(e.skey, e.tie, kth_hi, kth_lo)) {
                const int p = atomicAdd(&s_nwin, 1);
                if (p < 2048) s_win[p] = e;
            }
        }
        __syncthreads();
        const int m = min(s_nwin, 2048);
        int m2 = 1;
        while (m2 < m) m2 <<= 1;
        for (int i = m + threadIdx.x; i < m2; i += blockDim.x) s_win[i] = pad_entry();
        __syncthreads();
        bitonic_sort_cta(s_win, m2);
        for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
            ST s = -std::numeric_limits<ST>::infinity();
            int32_t d = -1;
            if (i < m && s_win[i].payload != 0xffffffffu) {
                key_score(s_win[i].skey, s);
                d = (int32_t)(id_base + (long long)s_win[i].payload);
            }
            out_scores[(size_t)q * k_out + i] = s;
            out_ids[(size_t)q * k_out + i] = d;
        }
        return;
    }
    if (in_smem)
        bitonic_sort_cta(a, n_pow2);
    else
        bitonic_sort_large(a, n_pow2, sm);
    for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
        ST s = -std::numeric_limits<ST>::infinity();
        int32_t d = -1;
        if (i < n_pow2) {
            Entry e = a[i];
            if (e.payload != 0xffffffffu) {
                key_score(e.skey, s);
                d = (int32_t)(id_base + (long long)e.payload);
            }
        }
        out_scores[(size_t)q * k_out + i] = s;
        out_ids[(size_t)q * k_out + i] = d;
    }
}

template <typename ST>
static int launch_segsort(const ST* scores, const int32_t* ids, int n_src, long long src_stride, int run_len,
                          long long seg_stride, int n_queries, int k_out, long long id_base, ST* out_scores,
                          int32_t* out_ids, void* ws, size_t ws_bytes, cudaStream_t stream) {
    long long n = (long long)n_src * run_len;
    FZ_REQUIRE(n >= 1 && n <= (1ll << 24), "segment length %lld out of range", n);
    int n_pow2 = next_pow2((int)n);
    size_t need = n_pow2 > kSmemEntries ? (size_t)n_queries * n_pow2 * sizeof(Entry) : 0;
    FZ_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
    size_t smem = (size_t)(n_pow2 > kSmemEntries ? kChunk : n_pow2) * sizeof(Entry);
    static bool attr_f = false, attr_d = false;
    bool& done = std::is_same<ST, float>::value ? attr_f : attr_d;
    if (!done) {
        FZ_CUDA(cudaFuncSetAttribute(segsort_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemEntries * (int)sizeof(Entry)));
        done = true;
    }
    ProfScope prof("segsort", stream);
    segsort_kernel<ST><<<n_queries, kSortThreads, smem, stream>>>(scores, ids, n_src, src_stride, run_len, seg_stride,
                                                                  k_out, id_base, out_scores, out_ids, (Entry*)ws,
                                                                  n_pow2);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz

using namespace fz;

extern "C" {

const char* fz_last_error(void) { return fz::last_error(); }
int fz_abi_version(void) { return FZ_ABI_VERSION; }

int fz_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(fz::g_prof_mu);
    fz::g_prof_on = on != 0;
    return FZ_OK;
}

int fz_profile_summary(char* out, size_t cap) {
    std::lock_guard<std::mutex> lk(fz::g_prof_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : fz::g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = agg[r.name];
        e.first += 1;
        e.second += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    fz::g_prof.clear();
    std::string s;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        s += line;
    }
    if (out && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return FZ_OK;
}

size_t fz_merge_topk_workspace_bytes(int n_src, int n_queries, int k_in) {
    long long n = (long long)n_src * k_in;
    int p = next_pow2((int)n);
    return p > kSmemEntries ? (size_t)n_queries * p * sizeof(Entry) : 0;
}

int fz_merge_topk_f32(const float* scores, const int32_t* ids, int n_src, int n_queries, int k_in, int k_out,
                      float* out_scores, int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream) {
    FZ_REQUIRE(scores && ids && out_scores && out_ids, "null pointer");
    FZ_REQUIRE(n_src >= 1 && n_queries >= 0 && k_in >= 1 && k_out >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    return launch_segsort<float>(scores, ids, n_src, (long long)n_queries * k_in, k_in, k_in, n_queries, k_out, 0,
                                 out_scores, out_ids, ws, ws_bytes, (cudaStream_t)stream);
}

int fz_merge_topk_f64(const double* scores, const int32_t* ids, int n_src, int n_queries, int k_in, int k_out,
                      double* out_scores, int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream) {
    FZ_REQUIRE(scores && ids && out_scores && out_ids, "null pointer");
    FZ_REQUIRE(n_src >= 1 && n_queries >= 0 && k_in >= 1 && k_out >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    return launch_segsort<double>(scores, ids, n_src, (long long)n_queries * k_in, k_in, k_in, n_queries, k_out, 0,
                                  out_scores, out_ids, ws, ws_bytes, (cudaStream_t)stream);
}

size_t fz_rank_rows_workspace_bytes(int n_queries, int64_t n_docs) {
    if (n_docs > (1ll << 24)) return 0;
    int p = next_pow2((int)n_docs);
    return p > kSmemEntries ? (size_t)n_queries * p * sizeof(Entry) : 0;
}

int fz_rank_rows_f32(const float* scores, int n_queries, int64_t n_docs, int k, int64_t doc_base, float* out_scores,
                     int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream) {
    FZ_REQUIRE(scores && out_scores && out_ids, "null pointer");
    FZ_REQUIRE(n_docs >= 1 && n_docs <= (1ll << 24) && k >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    return launch_segsort<float>(scores, nullptr, 1, 0, (int)n_docs, n_docs, n_queries, k, doc_base, out_scores,
                                 out_ids, ws, ws_bytes, (cudaStream_t)stream);
}

int fz_rank_rows_f64(const double* scores, int n_queries, int64_t n_docs, int k, int64_t doc_base, double* out_scores,
                     int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream) {
    FZ_REQUIRE(scores && out_scores && out_ids, "null pointer");
    FZ_REQUIRE(n_docs >= 1 && n_docs <= (1ll << 24) && k >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    return launch_segsort<double>(scores, nullptr, 1, 0, (int)n_docs, n_docs, n_queries, k, doc_base, out_scores,
                                  out_ids, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
