// Index build behind the C-ABI (SURVEY 8b "fz_build_*"): the CSR assembly of the reference's index-time code
// (src/retrievers/bm25.py:53-83 term statistics + postings, :141-143 BM25 statistics; the [N, V] activation matrix of
// splade/base.py:186-197 turned into an inverted index) and the layout work the scoring kernels depend on (the three
// storage forms of fz_postings_t with the bank-ordered tile segments, the SPLADE head matrix).  Off the query path: the
// sorts and scans are CUB (toolkit library, plumbing), the layout kernels are plain CUDA.
//
// Every builder is "plan, allocate, fill": the plan call returns the array sizes to the host (it synchronises the stream
// once), the caller allocates, the fill call writes.  Nothing here allocates device memory.
#include "common.cuh"

#include <cub/cub.cuh>
#include <cuda_bf16.h>

namespace fz {
namespace {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ int64_t lower_bound_dev(const T* a, int64_t lo, int64_t hi, T key) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// row r such that ptr[r] <= i < ptr[r + 1] (rows may be empty)
__device__ __forceinline__ int64_t row_of(const int64_t* ptr, int64_t n_rows, int64_t i) {
    int64_t lo = 0, hi = n_rows;          // invariant: ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

struct Carver {
    char* p;
    size_t left;
    bool ok = true;
    Carver(void* ws, size_t bytes) : p((char*)ws), left(bytes) {}
    template <typename T>
    T* take(size_t n) {
        const size_t b = align_up(n * sizeof(T), 256);
        if (b > left) { ok = false; return nullptr; }
        T* r = (T*)p;
        p += b;
        left -= b;
        return r;
    }
};

// ------------------------------------------------------------------------------------------------------------
// (row, term) entries -> term-major order
// ------------------------------------------------------------------------------------------------------------
// (an id outside its range is clamped and flagged: the call fails, nothing is written out of bounds)
__global__ void entry_key_kernel(const int32_t* row, const int32_t* term, int64_t n, int64_t n_rows, int32_t n_terms,
                                 uint64_t* key, int64_t* idx, int* bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t t = term[i], r = row[i];
    if (t < 0 || t >= n_terms || r < 0 || r >= n_rows) { *bad = 1; t = 0; r = 0; }
    key[i] = (uint64_t)t * (uint64_t)n_rows + (uint64_t)r;
    idx[i] = i;
}

__global__ void token_key_kernel(const int64_t* doc_ptr, const int32_t* tok, int64_t n, int64_t n_docs, int32_t vocab,
                                 uint64_t* key, int* bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t t = tok[i];
    if (t < 0 || t >= vocab) { *bad = 1; t = 0; }
    key[i] = (uint64_t)t * (uint64_t)n_docs + (uint64_t)row_of(doc_ptr, n_docs, i);
}

__global__ void term_hist_kernel(const uint64_t* key, int64_t n, uint64_t n_rows, unsigned long long* hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&hist[key[i] / n_rows], 1ull);
}

__global__ void run_flag_kernel(const uint64_t* key, int64_t n, int64_t* flag /* [n + 1] */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flag[i] = i < n && (i == 0 || key[i] != key[i - 1]);
}

// run r of equal keys starts at token start[r]; start[n_runs] = n
__global__ void run_start_kernel(const uint64_t* key, const int64_t* run_idx, int64_t n, uint64_t* ukey, int64_t* start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { start[run_idx[n]] = n; return; }
    if (i == 0 || key[i] != key[i - 1]) {
        ukey[run_idx[i]] = key[i];
        start[run_idx[i]] = i;
    }
}

__global__ void lexical_fill_kernel(const uint64_t* ukey, const int64_t* start, int64_t nnz, uint64_t n_docs,
                                    int32_t* post_doc, int32_t* post_tf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    post_doc[i] = (int32_t)(ukey[i] % n_docs);
    post_tf[i] = (int32_t)(start[i + 1] - start[i]);
}

__global__ void doc_len_kernel(const int64_t* doc_ptr, int64_t n_docs, int32_t* doc_len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_docs) doc_len[i] = (int32_t)(doc_ptr[i + 1] - doc_ptr[i]);
}

int key_bits(uint64_t max_key) {
    int b = 1;
    while (b < 64 && (max_key >> b)) ++b;
    return b;
}

// ------------------------------------------------------------------------------------------------------------
// postings layout
// ------------------------------------------------------------------------------------------------------------
__global__ void classify_kernel(const int64_t* term_ptr, int32_t n_terms, int64_t tiled_min, int64_t dense_min,
                                int32_t* is_tiled, int32_t* is_dense, int64_t* short_len) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_terms) return;
    if (t == n_terms) { is_tiled[t] = 0; is_dense[t] = 0; short_len[t] = 0; return; }     // trailing 0: the scans' totals
    const int64_t df = term_ptr[t + 1] - term_ptr[t];
    const bool dense = df >= dense_min, tiled = df >= tiled_min && !dense;
    is_tiled[t] = tiled;
    is_dense[t] = dense;
    short_len[t] = (tiled || dense) ? 0 : df;
}

__global__ void slot_kernel(const int32_t* is_tiled, const int32_t* is_dense, const int32_t* slot_tiled, const int32_t* slot_dense,
                            int32_t n_terms, int32_t* term_slot, int32_t* tiled_term) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    int32_t s = -1;
    if (is_tiled[t]) { s = slot_tiled[t]; tiled_term[s] = (int32_t)t; }
    else if (is_dense[t]) s = -2 - slot_dense[t];
    term_slot[t] = s;
}

// one CTA per tiled term: the padded segment starts of its row (relative to the row start) and the row total
__global__ void __launch_bounds__(kThreads)
tiled_rows_kernel(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* tiled_term, int32_t n_tiles, int32_t tile_docs,
                  uint32_t* tile_off /* [n_tiled, n_tiles + 1] or null */, int64_t* row_total) {
    using Scan = cub::BlockScan<unsigned long long, kThreads>;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ unsigned long long carry_s;
    const int r = blockIdx.x;
    const int32_t t = tiled_term[r];
    const int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n_tiles; j0 += kThreads) {
        const int j = j0 + threadIdx.x;
        unsigned long long pad = 0;
        if (j < n_tiles) {
            const int64_t a = lower_bound_dev<int32_t>(post_doc, lo, hi, (int32_t)min((int64_t)j * tile_docs, (int64_t)INT32_MAX));
            const int64_t b = j + 1 < n_tiles
                                  ? lower_bound_dev<int32_t>(post_doc, a, hi, (int32_t)min((int64_t)(j + 1) * tile_docs, (int64_t)INT32_MAX))
                                  : hi;
            pad = (unsigned long long)((b - a + 3) / 4 * 4);
        }
        unsigned long long excl, total;
        Scan(tmp).ExclusiveSum(pad, excl, total);
        const unsigned long long carry = carry_s;
        if (tile_off && j < n_tiles) tile_off[(size_t)r * (n_tiles + 1) + j] = (uint32_t)(carry + excl);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (tile_off) tile_off[(size_t)r * (n_tiles + 1) + n_tiles] = (uint32_t)carry_s;
        row_total[r] = (int64_t)carry_s;
    }
}

// short lists and dense rows: one thread per posting
template <typename V>
__global__ void short_dense_kernel(const int64_t* term_ptr, const int32_t* post_doc, const V* post_val, int64_t nnz, int32_t n_terms,
                                   const int64_t* short_ptr, const int32_t* term_slot, int64_t dense_stride, int32_t* short_doc,
                                   V* short_val, V* dense_val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t t = row_of(term_ptr, n_terms, i);
    const int32_t s = term_slot[t];
    if (s == -1) {
        const int64_t dst = short_ptr[t] + (i - term_ptr[t]);
        short_doc[dst] = post_doc[i];
        short_val[dst] = post_val[i];
    } else if (s <= -2) {
        dense_val[(size_t)(-2 - s) * dense_stride + post_doc[i]] = post_val[i];
    }
}

// coarse marks of the short lists: postings of the term below tile FZ_COARSE_TILES * c (0 for the other forms)
__global__ void coarse_kernel(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* term_slot, int32_t n_terms,
                              int32_t n_coarse, int64_t span_docs, uint16_t* coarse) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = n_coarse + 1;
    if (i >= (int64_t)n_terms * per) return;
    const int64_t t = i / per, c = i - t * per;
    uint16_t v = 0;
    if (term_slot[t] == -1) {
        const int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
        const int64_t key = c * span_docs;
        v = (uint16_t)((key > INT32_MAX ? hi : lower_bound_dev<int32_t>(post_doc, lo, hi, (int32_t)key)) - lo);
    }
    coarse[i] = v;
}

// Tiled segments.  One CTA per tiled term, its warps take runs of 32 tiles.  Inside a segment the postings (doc-ascending
// on input) are dealt round-robin over the 32 shared-memory banks of the scoring kernel's accumulator tile: posting with
// bank b = offset % 32 and rank r among the segment's postings of that bank gets sequence number
//     e = sum_b' min(cnt[b'], r) + #{b' < b : cnt[b'] > r}
// and a thread of the scoring kernel owns the postings of one 16-byte value vector (VEC = 4 fp32 / 2 fp64), so element e
// is stored at slot VEC * (e mod P/VEC) + e div (P/VEC) of the segment padded to P (a multiple of 4).
template <typename V>
__global__ void __launch_bounds__(kThreads)
tiled_fill_kernel(const int64_t* term_ptr, const int32_t* post_doc, const V* post_val, const int32_t* tiled_term,
                  const int64_t* tiled_base, const uint32_t* tile_off, int32_t n_tiles, int32_t tile_docs, uint16_t* out_off,
                  V* out_val) {
    constexpr int VEC = 16 / (int)sizeof(V);
    constexpr int kWarps = kThreads / 32;
    __shared__ int cnt_s[kWarps][32];
    __shared__ int run_s[kWarps][32];
    const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t t = tiled_term[r];
    const int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
    const int64_t base = tiled_base[r];
    const uint32_t* toff = tile_off + (size_t)r * (n_tiles + 1);
    int* cnt = cnt_s[warp];
    int* run = run_s[warp];
    for (int j0 = warp * 32; j0 < n_tiles; j0 += kWarps * 32) {
        const int j = j0 + lane;
        int64_t a = hi, b = hi;
        uint32_t start = 0, end = 0;
        if (j < n_tiles) {
            a = lower_bound_dev<int32_t>(post_doc, lo, hi, (int32_t)min((int64_t)j * tile_docs, (int64_t)INT32_MAX));
            b = j + 1 < n_tiles ? lower_bound_dev<int32_t>(post_doc, a, hi, (int32_t)min((int64_t)(j + 1) * tile_docs, (int64_t)INT32_MAX)) : hi;
            start = toff[j];
            end = toff[j + 1];
        }
        const int n_here = min(32, n_tiles - j0);
        for (int i = 0; i < n_here; ++i) {
            const int64_t s_lo = __shfl_sync(0xffffffffu, a, i), s_hi = __shfl_sync(0xffffffffu, b, i);
            const uint32_t s_start = __shfl_sync(0xffffffffu, start, i), s_end = __shfl_sync(0xffffffffu, end, i);
            const int len = (int)(s_hi - s_lo), pad = (int)(s_end - s_start);
            if (len == 0) continue;
            const int32_t doc0 = (j0 + i) * tile_docs;
            const int part = pad / VEC;
            // pass 1: postings per bank
            cnt[lane] = 0;
            run[lane] = 0;
            __syncwarp();
            for (int p = lane; p < len; p += 32) atomicAdd(&cnt[(post_doc[s_lo + p] - doc0) & 31], 1);
            __syncwarp();
            // pass 2: rank inside the bank (input order), sequence number, destination
            for (int p0 = 0; p0 < len; p0 += 32) {
                const int p = p0 + lane;
                const bool act = p < len;
                const int off = act ? post_doc[s_lo + p] - doc0 : 0;
                const int bank = act ? (off & 31) : 32 + lane;            // inactive lanes match nobody
                const unsigned same = __match_any_sync(0xffffffffu, bank);
                if (act) {
                    const int rk = run[bank] + __popc(same & ((1u << lane) - 1u));
                    int e = 0;
#pragma unroll 8
                    for (int bb = 0; bb < 32; ++bb) {
                        const int c = cnt[bb];
                        e += min(c, rk) + ((bb < bank && c > rk) ? 1 : 0);
                    }
                    const int64_t dst = base + s_start + (int64_t)VEC * (e % part) + e / part;
                    out_off[dst] = (uint16_t)off;
                    out_val[dst] = post_val[s_lo + p];
                }
                __syncwarp();
                if (act && (same & ((1u << lane) - 1u)) == 0) run[bank] += __popc(same);       // the bank's first lane
                __syncwarp();
            }
            // padding: sequence numbers len .. pad - 1
            for (int e = len + lane; e < pad; e += 32) {
                const int64_t dst = base + s_start + (int64_t)VEC * (e % part) + e / part;
                out_off[dst] = (uint16_t)tile_docs;
                out_val[dst] = V(0);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// SPLADE: row normalisation, term statistics, head matrix
// ------------------------------------------------------------------------------------------------------------
__global__ void csr_normalize_kernel(const int64_t* doc_ptr, const float* w, int64_t n_docs, float* out) {
    const int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= n_docs) return;
    const int64_t lo = doc_ptr[d], hi = doc_ptr[d + 1];
    double s = 0.0;
    for (int64_t i = lo + lane; i < hi; i += 32) s += (double)w[i] * (double)w[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = fmaxf(sqrtf((float)s), 1e-12f);
    for (int64_t i = lo + lane; i < hi; i += 32) out[i] = w[i] / nrm;
}

__global__ void term_stats_kernel(const int32_t* term, const float* w, int64_t nnz, int32_t n_terms, unsigned long long* df,
                                  float* term_max, int* negative) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int32_t t = term[i];
    if (t < 0 || t >= n_terms) { atomicExch(negative, 2); return; }
    atomicAdd(&df[t], 1ull);
    const float v = w[i];
    if (v < 0.0f) atomicExch(negative, 1);
    else atomicMax((int*)&term_max[t], __float_as_int(v));       // non-negative floats order like their bit patterns
}

__global__ void splade_head_kernel(const int64_t* doc_ptr, const int32_t* term, const float* w, int64_t nnz, int64_t n_docs,
                                   const int32_t* term_head, int head_dim, __nv_bfloat16* head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int32_t h = term_head[term[i]];
    if (h < 0) return;
    head[(size_t)row_of(doc_ptr, n_docs, i) * head_dim + h] = __float2bfloat16(w[i]);
}

size_t sort_temp_bytes(int64_t n, bool pairs) {
    size_t b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    if (pairs) {
        cub::DoubleBuffer<int64_t> v(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, b, k, v, n);
    } else {
        cub::DeviceRadixSort::SortKeys(nullptr, b, k, n);
    }
    return b;
}

template <typename T>
size_t scan_temp_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const T*)nullptr, (T*)nullptr, n);
    return b;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)ceil_div<int64_t>(n > 0 ? n : 1, kThreads); }

}  // namespace
}  // namespace fz

using namespace fz;

// ---------------------------------------------------------------------------------------------------------------
// term-major order of (row, term) entries
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t fz_build_term_major_workspace_bytes(int64_t nnz, int32_t n_terms) {
    if (nnz < 0 || n_terms < 0) return 0;
    const size_t n = (size_t)(nnz > 0 ? nnz : 1);
    return 2 * align_up(n * 8, 256) + 2 * align_up(n * 8, 256) + align_up(sort_temp_bytes(nnz, true), 256) +
           align_up(((size_t)n_terms + 1) * 8, 256) + align_up(scan_temp_bytes<unsigned long long>(n_terms + 1), 256) + 256 + 1024;
}

extern "C" int fz_build_term_major(const int32_t* row, const int32_t* term, int64_t nnz, int64_t n_rows, int32_t n_terms,
                                   int64_t* out_term_ptr, int64_t* out_order, void* ws, size_t ws_bytes, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(nnz >= 0 && n_rows >= 0 && n_terms >= 0 && out_term_ptr, "fz_build_term_major: bad sizes");
    FZ_REQUIRE(nnz == 0 || (row && term && out_order && n_rows > 0 && n_terms > 0), "fz_build_term_major: null input");
    FZ_REQUIRE((double)n_terms * (double)(n_rows > 0 ? n_rows : 1) < 1.8e19, "fz_build_term_major: term * rows overflows 64 bits");
    FZ_REQUIRE(ws && ws_bytes >= fz_build_term_major_workspace_bytes(nnz, n_terms), "fz_build_term_major: workspace too small");
    Carver c(ws, ws_bytes);
    const size_t n = (size_t)(nnz > 0 ? nnz : 1);
    uint64_t* k0 = c.take<uint64_t>(n);
    uint64_t* k1 = c.take<uint64_t>(n);
    int64_t* v0 = c.take<int64_t>(n);
    int64_t* v1 = c.take<int64_t>(n);
    size_t sort_b = sort_temp_bytes(nnz, true);
    void* sort_t = c.take<char>(sort_b);
    unsigned long long* hist = c.take<unsigned long long>((size_t)n_terms + 1);
    size_t scan_b = scan_temp_bytes<unsigned long long>(n_terms + 1);
    void* scan_t = c.take<char>(scan_b);
    int* bad = c.take<int>(1);
    FZ_REQUIRE(c.ok, "fz_build_term_major: workspace too small");
    FZ_CUDA(cudaMemsetAsync(hist, 0, ((size_t)n_terms + 1) * 8, stream));
    FZ_CUDA(cudaMemsetAsync(bad, 0, 4, stream));
    if (nnz > 0) {
        entry_key_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(row, term, nnz, n_rows, n_terms, k0, v0, bad);
        FZ_LAUNCH_CHECK();
        cub::DoubleBuffer<uint64_t> kb(k0, k1);
        cub::DoubleBuffer<int64_t> vb(v0, v1);
        const int bits = key_bits((uint64_t)n_terms * (uint64_t)n_rows);
        FZ_CUDA(cub::DeviceRadixSort::SortPairs(sort_t, sort_b, kb, vb, nnz, 0, bits, stream));
        term_hist_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(kb.Current(), nnz, (uint64_t)n_rows, hist);
        FZ_LAUNCH_CHECK();
        FZ_CUDA(cudaMemcpyAsync(out_order, vb.Current(), (size_t)nnz * 8, cudaMemcpyDeviceToDevice, stream));
    }
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(scan_t, scan_b, hist, (unsigned long long*)out_term_ptr, n_terms + 1, stream));
    int bad_h = 0;
    FZ_CUDA(cudaMemcpyAsync(&bad_h, bad, 4, cudaMemcpyDeviceToHost, stream));
    FZ_CUDA(cudaStreamSynchronize(stream));
    FZ_REQUIRE(!bad_h, "fz_build_term_major: a row or term id lies outside its range");
    return FZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lexical postings from token ids: (term, doc, tf), term-major / doc-ascending, + document lengths
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct LexWs {
    uint64_t *k0, *k1, *ukey;
    int64_t *run_idx, *start;
    int* bad;
    void* temp;
    size_t temp_b;
    unsigned long long* hist;
    void* scan_t;
    size_t scan_b;
    bool ok;
};
size_t lex_temp_bytes(int64_t n) {
    const size_t a = sort_temp_bytes(n, false), b = scan_temp_bytes<int64_t>(n + 1);
    return a > b ? a : b;
}
LexWs lex_carve(void* ws, size_t ws_bytes, int64_t n_tokens, int32_t vocab) {
    Carver c(ws, ws_bytes);
    const size_t n = (size_t)n_tokens + 1;
    LexWs w;
    w.k0 = c.take<uint64_t>(n);
    w.k1 = c.take<uint64_t>(n);
    w.ukey = c.take<uint64_t>(n);
    w.run_idx = c.take<int64_t>(n);
    w.start = c.take<int64_t>(n);
    w.bad = c.take<int>(1);
    w.temp_b = lex_temp_bytes(n_tokens);
    w.temp = c.take<char>(w.temp_b);
    w.hist = c.take<unsigned long long>((size_t)vocab + 1);
    w.scan_b = scan_temp_bytes<unsigned long long>(vocab + 1);
    w.scan_t = c.take<char>(w.scan_b);
    w.ok = c.ok;
    return w;
}
}  // namespace

extern "C" size_t fz_build_lexical_workspace_bytes(int64_t n_tokens, int32_t vocab) {
    if (n_tokens < 0 || vocab < 0) return 0;
    const size_t n = (size_t)n_tokens + 1;
    return 5 * align_up(n * 8, 256) + 256 + align_up(lex_temp_bytes(n_tokens), 256) + align_up(((size_t)vocab + 1) * 8, 256) +
           align_up(scan_temp_bytes<unsigned long long>(vocab + 1), 256) + 1024;
}

extern "C" int fz_build_lexical_plan(const int64_t* doc_ptr, const int32_t* doc_tok, int64_t n_docs, int64_t n_tokens,
                                     int32_t vocab, int64_t* out_nnz_h, void* ws, size_t ws_bytes, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(n_docs >= 0 && n_tokens >= 0 && vocab >= 0 && out_nnz_h && doc_ptr, "fz_build_lexical_plan: bad arguments");
    FZ_REQUIRE(n_tokens == 0 || (doc_tok && n_docs > 0 && vocab > 0), "fz_build_lexical_plan: null input");
    FZ_REQUIRE((double)vocab * (double)(n_docs > 0 ? n_docs : 1) < 1.8e19, "fz_build_lexical_plan: term * docs overflows 64 bits");
    FZ_REQUIRE(ws && ws_bytes >= fz_build_lexical_workspace_bytes(n_tokens, vocab), "fz_build_lexical_plan: workspace too small");
    LexWs w = lex_carve(ws, ws_bytes, n_tokens, vocab);
    FZ_REQUIRE(w.ok, "fz_build_lexical_plan: workspace too small");
    *out_nnz_h = 0;
    if (n_tokens == 0) return FZ_OK;
    FZ_CUDA(cudaMemsetAsync(w.bad, 0, 4, stream));
    token_key_kernel<<<blocks_for(n_tokens), kThreads, 0, stream>>>(doc_ptr, doc_tok, n_tokens, n_docs, vocab, w.k0, w.bad);
    FZ_LAUNCH_CHECK();
    cub::DoubleBuffer<uint64_t> kb(w.k0, w.k1);
    const int bits = key_bits((uint64_t)vocab * (uint64_t)n_docs);
    size_t tb = w.temp_b;
    FZ_CUDA(cub::DeviceRadixSort::SortKeys(w.temp, tb, kb, n_tokens, 0, bits, stream));
    // runs of equal (term, doc) keys: one posting each, tf = run length
    run_flag_kernel<<<blocks_for(n_tokens + 1), kThreads, 0, stream>>>(kb.Current(), n_tokens, w.start);
    FZ_LAUNCH_CHECK();
    tb = w.temp_b;
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, w.start, w.run_idx, n_tokens + 1, stream));
    run_start_kernel<<<blocks_for(n_tokens + 1), kThreads, 0, stream>>>(kb.Current(), w.run_idx, n_tokens, w.ukey, w.start);
    FZ_LAUNCH_CHECK();
    FZ_CUDA(cudaMemcpyAsync(out_nnz_h, w.run_idx + n_tokens, 8, cudaMemcpyDeviceToHost, stream));
    int bad_h = 0;
    FZ_CUDA(cudaMemcpyAsync(&bad_h, w.bad, 4, cudaMemcpyDeviceToHost, stream));
    FZ_CUDA(cudaStreamSynchronize(stream));
    FZ_REQUIRE(!bad_h, "fz_build_lexical_plan: a token id lies outside [0, vocab)");
    return FZ_OK;
}

extern "C" int fz_build_lexical_fill(const int64_t* doc_ptr, int64_t n_docs, int64_t n_tokens, int32_t vocab, int64_t nnz,
                                     int64_t* out_term_ptr, int32_t* out_post_doc, int32_t* out_post_tf, int32_t* out_doc_len,
                                     void* ws, size_t ws_bytes, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(n_docs >= 0 && n_tokens >= 0 && vocab >= 0 && nnz >= 0 && nnz <= n_tokens && out_term_ptr && doc_ptr,
               "fz_build_lexical_fill: bad arguments");
    FZ_REQUIRE(nnz == 0 || (out_post_doc && out_post_tf), "fz_build_lexical_fill: null output");
    FZ_REQUIRE(ws && ws_bytes >= fz_build_lexical_workspace_bytes(n_tokens, vocab), "fz_build_lexical_fill: workspace too small");
    LexWs w = lex_carve(ws, ws_bytes, n_tokens, vocab);      // same carving as the plan call: the runs are still there
    FZ_REQUIRE(w.ok, "fz_build_lexical_fill: workspace too small");
    FZ_CUDA(cudaMemsetAsync(w.hist, 0, ((size_t)vocab + 1) * 8, stream));
    if (nnz > 0) {
        lexical_fill_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(w.ukey, w.start, nnz, (uint64_t)n_docs, out_post_doc, out_post_tf);
        FZ_LAUNCH_CHECK();
        term_hist_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(w.ukey, nnz, (uint64_t)n_docs, w.hist);
        FZ_LAUNCH_CHECK();
    }
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, w.scan_b, w.hist, (unsigned long long*)out_term_ptr, vocab + 1, stream));
    if (out_doc_len && n_docs > 0) {
        doc_len_kernel<<<blocks_for(n_docs), kThreads, 0, stream>>>(doc_ptr, n_docs, out_doc_len);
        FZ_LAUNCH_CHECK();
    }
    return FZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the three storage forms of fz_postings_t
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct PostWs {
    int32_t *is_tiled, *is_dense, *slot_tiled, *slot_dense, *tiled_term;
    int64_t *short_len, *row_total, *row_base;
    void* scan_t;
    size_t scan_b;
    bool ok;
};
size_t post_scan_bytes(int32_t n_terms) {
    const size_t a = scan_temp_bytes<int32_t>(n_terms + 1), b = scan_temp_bytes<int64_t>(n_terms + 1);
    return a > b ? a : b;
}
PostWs post_carve(void* ws, size_t ws_bytes, int32_t n_terms) {
    Carver c(ws, ws_bytes);
    const size_t n = (size_t)n_terms + 1;
    PostWs w;
    w.is_tiled = c.take<int32_t>(n);
    w.is_dense = c.take<int32_t>(n);
    w.slot_tiled = c.take<int32_t>(n);
    w.slot_dense = c.take<int32_t>(n);
    w.tiled_term = c.take<int32_t>(n);
    w.short_len = c.take<int64_t>(n);
    w.row_total = c.take<int64_t>(n);
    w.row_base = c.take<int64_t>(n);
    w.scan_b = post_scan_bytes(n_terms);
    w.scan_t = c.take<char>(w.scan_b);
    w.ok = c.ok;
    return w;
}
}  // namespace

extern "C" size_t fz_build_postings_workspace_bytes(int32_t n_terms) {
    if (n_terms < 0) return 0;
    const size_t n = (size_t)n_terms + 1;
    return 5 * align_up(n * 4, 256) + 3 * align_up(n * 8, 256) + align_up(post_scan_bytes(n_terms), 256) + 1024;
}

extern "C" int fz_build_postings_plan(const int64_t* term_ptr, const int32_t* post_doc, int32_t n_terms, int64_t n_docs,
                                      int32_t tile_docs, int64_t tiled_min, int64_t dense_min, int64_t* out_short_ptr,
                                      int32_t* out_term_slot, fz_build_plan_t* out_plan_h, void* ws, size_t ws_bytes,
                                      fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(term_ptr && out_short_ptr && out_plan_h && n_terms >= 0 && n_docs >= 0, "fz_build_postings_plan: bad arguments");
    FZ_REQUIRE(n_terms == 0 || out_term_slot, "fz_build_postings_plan: null term_slot");
    FZ_REQUIRE(tile_docs >= 4 && tile_docs <= 32768 && tile_docs % 4 == 0, "tile_docs=%d must be a multiple of 4 in [4, 32768]", tile_docs);
    FZ_REQUIRE(tiled_min >= 0 && tiled_min <= 65535, "tiled_min must be <= 65535 (short-list offsets are 16 bits)");
    FZ_REQUIRE(dense_min >= tiled_min, "dense_min must be >= tiled_min");
    FZ_REQUIRE(n_docs <= INT32_MAX, "n_docs exceeds the 31-bit local doc rows of a shard");
    FZ_REQUIRE(ws && ws_bytes >= fz_build_postings_workspace_bytes(n_terms), "fz_build_postings_plan: workspace too small");
    PostWs w = post_carve(ws, ws_bytes, n_terms);
    FZ_REQUIRE(w.ok, "fz_build_postings_plan: workspace too small");
    fz_build_plan_t P;
    memset(&P, 0, sizeof(P));
    P.n_tiles = (int32_t)ceil_div<int64_t>(n_docs, tile_docs);
    P.n_coarse = (P.n_tiles + FZ_COARSE_TILES - 1) / FZ_COARSE_TILES;
    P.dense_stride = (int64_t)P.n_tiles * tile_docs;
    const int64_t n1 = (int64_t)n_terms + 1;
    classify_kernel<<<blocks_for(n1), kThreads, 0, stream>>>(term_ptr, n_terms, tiled_min, dense_min, w.is_tiled, w.is_dense, w.short_len);
    FZ_LAUNCH_CHECK();
    size_t sb = w.scan_b;
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, sb, w.is_tiled, w.slot_tiled, n1, stream));
    sb = w.scan_b;
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, sb, w.is_dense, w.slot_dense, n1, stream));
    sb = w.scan_b;
    FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, sb, w.short_len, out_short_ptr, n1, stream));
    if (n_terms > 0) {
        slot_kernel<<<blocks_for(n_terms), kThreads, 0, stream>>>(w.is_tiled, w.is_dense, w.slot_tiled, w.slot_dense, n_terms,
                                                                  out_term_slot, w.tiled_term);
        FZ_LAUNCH_CHECK();
    }
    int32_t counts[2];
    FZ_CUDA(cudaMemcpyAsync(&counts[0], w.slot_tiled + n_terms, 4, cudaMemcpyDeviceToHost, stream));
    FZ_CUDA(cudaMemcpyAsync(&counts[1], w.slot_dense + n_terms, 4, cudaMemcpyDeviceToHost, stream));
    FZ_CUDA(cudaMemcpyAsync(&P.n_short, out_short_ptr + n_terms, 8, cudaMemcpyDeviceToHost, stream));
    FZ_CUDA(cudaStreamSynchronize(stream));
    P.n_tiled = counts[0];
    P.n_dense = counts[1];
    if (P.n_tiled > 0) {
        FZ_REQUIRE(post_doc, "fz_build_postings_plan: null post_doc");
        tiled_rows_kernel<<<P.n_tiled, kThreads, 0, stream>>>(term_ptr, post_doc, w.tiled_term, P.n_tiles, tile_docs, nullptr, w.row_total);
        FZ_LAUNCH_CHECK();
        FZ_CUDA(cudaMemsetAsync(w.row_total + P.n_tiled, 0, 8, stream));
        sb = w.scan_b;
        FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, sb, w.row_total, w.row_base, (int64_t)P.n_tiled + 1, stream));
        FZ_CUDA(cudaMemcpyAsync(&P.n_tiled_entries, w.row_base + P.n_tiled, 8, cudaMemcpyDeviceToHost, stream));
        FZ_CUDA(cudaStreamSynchronize(stream));
    }
    *out_plan_h = P;
    return FZ_OK;
}

template <typename V>
static int postings_fill(const int64_t* term_ptr, const int32_t* post_doc, const V* post_val, int32_t n_terms, int64_t n_docs,
                         int32_t tile_docs, const int64_t* short_ptr, const int32_t* term_slot, const fz_build_plan_t& P,
                         int32_t* short_doc, V* short_val, uint16_t* short_coarse, int64_t* tiled_base, uint32_t* tile_off,
                         uint16_t* tiled_off, V* tiled_val, V* dense_val, PostWs& w, int64_t nnz, cudaStream_t stream) {
    if (P.n_dense > 0) FZ_CUDA(cudaMemsetAsync(dense_val, 0, (size_t)P.n_dense * (size_t)P.dense_stride * sizeof(V), stream));
    if (nnz > 0 && (P.n_short > 0 || P.n_dense > 0)) {
        short_dense_kernel<V><<<blocks_for(nnz), kThreads, 0, stream>>>(term_ptr, post_doc, post_val, nnz, n_terms, short_ptr, term_slot,
                                                                        P.dense_stride, short_doc, short_val, dense_val);
        FZ_LAUNCH_CHECK();
    }
    if (n_terms > 0) {
        const int64_t n = (int64_t)n_terms * (P.n_coarse + 1);
        coarse_kernel<<<blocks_for(n), kThreads, 0, stream>>>(term_ptr, post_doc, term_slot, n_terms, P.n_coarse,
                                                              (int64_t)FZ_COARSE_TILES * tile_docs, short_coarse);
        FZ_LAUNCH_CHECK();
    }
    if (P.n_tiled > 0) {
        tiled_rows_kernel<<<P.n_tiled, kThreads, 0, stream>>>(term_ptr, post_doc, w.tiled_term, P.n_tiles, tile_docs, tile_off, w.row_total);
        FZ_LAUNCH_CHECK();
        FZ_CUDA(cudaMemsetAsync(w.row_total + P.n_tiled, 0, 8, stream));
        size_t sb = w.scan_b;
        FZ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_t, sb, w.row_total, w.row_base, (int64_t)P.n_tiled + 1, stream));
        FZ_CUDA(cudaMemcpyAsync(tiled_base, w.row_base, (size_t)P.n_tiled * 8, cudaMemcpyDeviceToDevice, stream));
        tiled_fill_kernel<V><<<P.n_tiled, kThreads, 0, stream>>>(term_ptr, post_doc, post_val, w.tiled_term, tiled_base, tile_off,
                                                                 P.n_tiles, tile_docs, tiled_off, tiled_val);
        FZ_LAUNCH_CHECK();
    }
    return FZ_OK;
}

extern "C" int fz_build_postings_fill(const int64_t* term_ptr, const int32_t* post_doc, const void* post_val, int value_bytes,
                                      int32_t n_terms, int64_t n_docs, int32_t tile_docs, const int64_t* short_ptr,
                                      const int32_t* term_slot, const fz_build_plan_t* plan, int64_t nnz, int32_t* out_short_doc,
                                      void* out_short_val, uint16_t* out_short_coarse, int64_t* out_tiled_base,
                                      uint32_t* out_tiled_tile_off, uint16_t* out_tiled_off, void* out_tiled_val,
                                      void* out_dense_val, void* ws, size_t ws_bytes, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(term_ptr && short_ptr && plan && n_terms >= 0 && n_docs >= 0 && nnz >= 0, "fz_build_postings_fill: bad arguments");
    FZ_REQUIRE(value_bytes == 4 || value_bytes == 8, "value_bytes must be 4 (float weights) or 8 (double impacts)");
    FZ_REQUIRE(tile_docs >= 4 && tile_docs <= 32768 && tile_docs % 4 == 0, "tile_docs=%d must be a multiple of 4 in [4, 32768]", tile_docs);
    FZ_REQUIRE(nnz == 0 || (post_doc && post_val && term_slot), "fz_build_postings_fill: null input");
    FZ_REQUIRE(plan->n_tiles == (int32_t)ceil_div<int64_t>(n_docs, tile_docs), "fz_build_postings_fill: plan does not match n_docs / tile_docs");
    FZ_REQUIRE(plan->n_short == 0 || (out_short_doc && out_short_val), "fz_build_postings_fill: null short-list output");
    FZ_REQUIRE(n_terms == 0 || out_short_coarse, "fz_build_postings_fill: null short_coarse");
    FZ_REQUIRE(plan->n_tiled == 0 || (out_tiled_base && out_tiled_tile_off), "fz_build_postings_fill: null tiled output");
    FZ_REQUIRE(plan->n_tiled_entries == 0 || (out_tiled_off && out_tiled_val), "fz_build_postings_fill: null tiled output");
    FZ_REQUIRE(plan->n_dense == 0 || out_dense_val, "fz_build_postings_fill: null dense output");
    FZ_REQUIRE(ws && ws_bytes >= fz_build_postings_workspace_bytes(n_terms), "fz_build_postings_fill: workspace too small");
    PostWs w = post_carve(ws, ws_bytes, n_terms);       // tiled_term of the plan call is still there
    FZ_REQUIRE(w.ok, "fz_build_postings_fill: workspace too small");
    if (value_bytes == 4)
        return postings_fill<float>(term_ptr, post_doc, (const float*)post_val, n_terms, n_docs, tile_docs, short_ptr, term_slot, *plan,
                                    out_short_doc, (float*)out_short_val, out_short_coarse, out_tiled_base, out_tiled_tile_off,
                                    out_tiled_off, (float*)out_tiled_val, (float*)out_dense_val, w, nnz, stream);
    return postings_fill<double>(term_ptr, post_doc, (const double*)post_val, n_terms, n_docs, tile_docs, short_ptr, term_slot, *plan,
                                 out_short_doc, (double*)out_short_val, out_short_coarse, out_tiled_base, out_tiled_tile_off,
                                 out_tiled_off, (double*)out_tiled_val, (double*)out_dense_val, w, nnz, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// SPLADE helpers
// ---------------------------------------------------------------------------------------------------------------
extern "C" int fz_build_csr_normalize(const int64_t* doc_ptr, const float* weight, int64_t n_docs, float* out_weight,
                                      fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(doc_ptr && n_docs >= 0, "fz_build_csr_normalize: bad arguments");
    if (n_docs == 0) return FZ_OK;
    FZ_REQUIRE(weight && out_weight, "fz_build_csr_normalize: null weights");
    csr_normalize_kernel<<<blocks_for(n_docs * 32), kThreads, 0, stream>>>(doc_ptr, weight, n_docs, out_weight);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

extern "C" int fz_build_term_stats(const int32_t* term, const float* weight, int64_t nnz, int32_t n_terms, int64_t* out_df,
                                   float* out_term_max, int32_t* out_flags, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(nnz >= 0 && n_terms >= 0 && out_flags && (n_terms == 0 || (out_df && out_term_max)), "fz_build_term_stats: bad arguments");
    if (n_terms > 0) {
        FZ_CUDA(cudaMemsetAsync(out_df, 0, (size_t)n_terms * 8, stream));
        FZ_CUDA(cudaMemsetAsync(out_term_max, 0, (size_t)n_terms * 4, stream));
    }
    FZ_CUDA(cudaMemsetAsync(out_flags, 0, 4, stream));
    if (nnz == 0) return FZ_OK;
    FZ_REQUIRE(term && weight, "fz_build_term_stats: null input");
    term_stats_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(term, weight, nnz, n_terms, (unsigned long long*)out_df, out_term_max, out_flags);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

extern "C" int fz_build_splade_head(const int64_t* doc_ptr, const int32_t* term, const float* weight, int64_t n_docs,
                                    int64_t nnz, const int32_t* term_head, int32_t head_dim, void* out_head_bf16,
                                    fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(doc_ptr && n_docs >= 0 && nnz >= 0 && out_head_bf16, "fz_build_splade_head: bad arguments");
    FZ_REQUIRE(head_dim % 64 == 0 && head_dim >= 64 && head_dim <= 256, "head_dim=%d must be 64, 128, 192 or 256", head_dim);
    FZ_CUDA(cudaMemsetAsync(out_head_bf16, 0, (size_t)n_docs * head_dim * 2, stream));
    if (nnz == 0) return FZ_OK;
    FZ_REQUIRE(term && weight && term_head, "fz_build_splade_head: null input");
    splade_head_kernel<<<blocks_for(nnz), kThreads, 0, stream>>>(doc_ptr, term, weight, nnz, n_docs, term_head, head_dim,
                                                                 (__nv_bfloat16*)out_head_bf16);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}
