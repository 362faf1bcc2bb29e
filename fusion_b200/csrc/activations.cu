// SPLADE activation head: what sits between the MLM logits of the (stock PyTorch) encoder and the
// sparse index / query CSR that K2 consumes.
//
//   fz_splade_pool     logits [B, L, V] + attention mask [B, L] -> activations [B, V]:
//                      sum_l / amax_l log1p(relu(logits * mask))          (src/retrievers/splade/splade.py:88-94)
//   fz_prune_topk      keep the keep_topk largest activations of every row, zero the rest
//                      (SPLADE._prune_activations, splade.py:295-306: torch.topk + scatter into zeros)
//   fz_csr_count/fill  dense [B, V] activations -> CSR (term ids ascending, zeros dropped): the reference keeps SPLADE
//                      vectors dense ([N, V] fp32 = 1.1 TB at 8.8M passages); the inverted index is built from this CSR.
//
// All three are single-pass streaming kernels bound by HBM: the pool reads every unmasked logit exactly once
// (B * sum_l(mask) * V elements), padded positions are never loaded.
#include <cuda_bf16.h>

#include "common.cuh"
#include "radix_select.cuh"

namespace fz {

constexpr int kPoolThreads = 256;
constexpr int kPoolUnroll = 16;      // independent loads in flight per thread

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }

// relu as torch computes it: max(x, 0) with NaN passed through
__device__ __forceinline__ float relu_f32(float x) { return x > 0.0f ? x : (x != x ? x : 0.0f); }

// One thread per (row b, vocabulary entry v); consecutive threads read consecutive v of one sequence position, so every
// warp request is one contiguous run of the logits.  Positions with mask == 0 contribute log1p(relu(0)) = 0 and are never
// loaded: the CTA first compacts the unmasked positions of its row into shared memory, then every thread streams over
// that list with kPoolUnroll independent loads in flight (a load that waits for its mask value first is latency-bound).
//   POOL 0 (max): log1p is monotone, so amax_l log1p(relu(x_l)) = log1p(relu(max_l x_l)): one log1p per output.
//   POOL 1 (sum): fp32 sum in sequence order (torch.sum's order over a strided dim is unspecified: parity to ~1e-6).
template <typename T, int POOL>
__global__ void __launch_bounds__(kPoolThreads) splade_pool_kernel(const T* __restrict__ logits, const int32_t* __restrict__ mask,
                                                                  int L, int V, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* s_pos = reinterpret_cast<int*>(smem_raw);              // [L] unmasked positions, ascending
    float* s_mk = reinterpret_cast<float*>(s_pos + L);          // [L] their mask values (1.0 for a 0/1 mask)
    __shared__ int s_n;
    const int b = blockIdx.y;
    const int32_t* __restrict__ m = mask + (size_t)b * L;
    if (threadIdx.x < 32) {
        int n = 0;
        for (int l0 = 0; l0 < L; l0 += 32) {
            const int l = l0 + threadIdx.x;
            const int mk = l < L ? m[l] : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, mk != 0);
            if (mk != 0) {
                const int p = n + __popc(bal & ((1u << threadIdx.x) - 1));
                s_pos[p] = l;
                s_mk[p] = (float)mk;
            }
            n += __popc(bal);
        }
        if (threadIdx.x == 0) s_n = n;
    }
    __syncthreads();
    const int v = blockIdx.x * kPoolThreads + threadIdx.x;
    if (v >= V) return;
    const int n = s_n;
    const T* __restrict__ p = logits + (size_t)b * L * V + v;
    float acc = 0.0f;
    auto fold = [&](float x) {
        const float r = relu_f32(x);
        if (POOL == 0) acc = (r != r) ? r : (acc != acc ? acc : fmaxf(acc, r));
        else acc += r > 0.0f ? log1pf(r) : r;        // relu zeros (most real logits) skip the log1p
    };
    int i = 0;
    for (; i + kPoolUnroll <= n; i += kPoolUnroll) {
        float x[kPoolUnroll];
#pragma unroll
        for (int u = 0; u < kPoolUnroll; ++u) x[u] = to_f32(p[(size_t)s_pos[i + u] * V]);
#pragma unroll
        for (int u = 0; u < kPoolUnroll; ++u) fold(x[u] * s_mk[i + u]);
    }
    for (; i < n; ++i) fold(to_f32(p[(size_t)s_pos[i] * V]) * s_mk[i]);
    out[(size_t)b * V + v] = POOL == 0 ? log1pf(acc) : acc;
}

// Keep the k largest values of the row (ties at the cutoff: the lower term id stays), zero the others.
__global__ void __launch_bounds__(512) prune_topk_kernel(const float* act, int V, int k, float* out) {
    __shared__ int s_hist[256];
    __shared__ int s_bcast[4];
    const float* row = act + (size_t)blockIdx.x * V;      // out may alias act (in-place pruning)
    float* dst = out + (size_t)blockIdx.x * V;
    uint32_t kth_hi = 0, kth_lo = 0;
    cta_radix_select_kth_by<uint32_t>([row](int i) { return ord32(row[i] + 0.0f); }, [](int i) { return ~(uint32_t)i; }, V, k,
                                      s_hist, s_bcast, kth_hi, kth_lo);
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float x = row[i];
        dst[i] = key_ge<uint32_t>(ord32(x + 0.0f), ~(uint32_t)i, kth_hi, kth_lo) ? x : 0.0f;
    }
}

__global__ void __launch_bounds__(256) csr_count_kernel(const float* __restrict__ act, int V, int32_t* __restrict__ nnz) {
    __shared__ int s_w[8];
    const float* __restrict__ row = act + (size_t)blockIdx.x * V;
    int c = 0;
#pragma unroll 8
    for (int i = threadIdx.x; i < V; i += 256) c += row[i] != 0.0f;
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        nnz[blockIdx.x] = t;
    }
}

// Ordered compaction of one row: term ids ascending, as a CSR row must be.  A step covers kCsrSub * 256 entries with
// all of a thread's loads in flight and one barrier pair.
constexpr int kCsrSub = 8;
__global__ void __launch_bounds__(256) csr_fill_kernel(const float* __restrict__ act, int V, const int64_t* __restrict__ row_ptr,
                                                       int32_t* __restrict__ out_term, float* __restrict__ out_w) {
    __shared__ int s_w[kCsrSub][8];
    const float* __restrict__ row = act + (size_t)blockIdx.x * V;
    const int64_t base = row_ptr[blockIdx.x];
    if (row_ptr[blockIdx.x + 1] == base) return;                 // empty row: nothing to write
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int done = 0;
    for (int i0 = 0; i0 < V; i0 += kCsrSub * 256) {
        float x[kCsrSub];
        unsigned bal[kCsrSub];
#pragma unroll
        for (int u = 0; u < kCsrSub; ++u) {
            const int i = i0 + u * 256 + threadIdx.x;
            x[u] = i < V ? row[i] : 0.0f;
        }
        __syncthreads();            // s_w of the previous step has been read
#pragma unroll
        for (int u = 0; u < kCsrSub; ++u) {
            bal[u] = __ballot_sync(0xffffffffu, x[u] != 0.0f);
            if (lane == 0) s_w[u][warp] = __popc(bal[u]);
        }
        __syncthreads();
        int run = done;
#pragma unroll
        for (int u = 0; u < kCsrSub; ++u) {
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int c = s_w[u][w];
                if (w < warp) before += c;
                total += c;
            }
            if (x[u] != 0.0f) {
                const int64_t p = base + run + before + __popc(bal[u] & ((1u << lane) - 1));
                out_term[p] = i0 + u * 256 + threadIdx.x;
                out_w[p] = x[u];
            }
            run += total;
        }
        done = run;
    }
}

}  // namespace fz

using namespace fz;

extern "C" {

int fz_splade_pool(const void* logits, int logits_are_bf16, const int32_t* mask, int n_rows, int seq_len, int vocab, int pooling,
                   float* out_act, fz_stream_t stream) {
    FZ_REQUIRE(n_rows >= 0 && seq_len >= 1 && vocab >= 1, "bad shape [%d, %d, %d]", n_rows, seq_len, vocab);
    FZ_REQUIRE(pooling == FZ_POOL_MAX || pooling == FZ_POOL_SUM, "pooling must be FZ_POOL_MAX or FZ_POOL_SUM");
    if (n_rows == 0) return FZ_OK;
    FZ_REQUIRE(logits && mask && out_act, "null pointer");
    FZ_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    const dim3 grid(ceil_div(vocab, kPoolThreads), n_rows);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)seq_len * 8;
    FZ_REQUIRE(smem <= 48 * 1024, "seq_len=%d too long (max 6144)", seq_len);
    ProfScope prof("splade_pool", s);
    if (logits_are_bf16) {
        const __nv_bfloat16* p = (const __nv_bfloat16*)logits;
        if (pooling == FZ_POOL_MAX) splade_pool_kernel<__nv_bfloat16, 0><<<grid, kPoolThreads, smem, s>>>(p, mask, seq_len, vocab, out_act);
        else splade_pool_kernel<__nv_bfloat16, 1><<<grid, kPoolThreads, smem, s>>>(p, mask, seq_len, vocab, out_act);
    } else {
        const float* p = (const float*)logits;
        if (pooling == FZ_POOL_MAX) splade_pool_kernel<float, 0><<<grid, kPoolThreads, smem, s>>>(p, mask, seq_len, vocab, out_act);
        else splade_pool_kernel<float, 1><<<grid, kPoolThreads, smem, s>>>(p, mask, seq_len, vocab, out_act);
    }
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_prune_topk(const float* act, int n_rows, int vocab, int keep_topk, float* out_act, fz_stream_t stream) {
    FZ_REQUIRE(n_rows >= 0 && vocab >= 1, "bad shape [%d, %d]", n_rows, vocab);
    FZ_REQUIRE(keep_topk >= 1 && keep_topk <= vocab, "keep_topk=%d out of range [1, %d]", keep_topk, vocab);
    if (n_rows == 0) return FZ_OK;
    FZ_REQUIRE(act && out_act, "null pointer");
    prune_topk_kernel<<<n_rows, 512, 0, (cudaStream_t)stream>>>(act, vocab, keep_topk, out_act);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_csr_count(const float* act, int n_rows, int vocab, int32_t* out_nnz, fz_stream_t stream) {
    FZ_REQUIRE(n_rows >= 0 && vocab >= 1, "bad shape [%d, %d]", n_rows, vocab);
    if (n_rows == 0) return FZ_OK;
    FZ_REQUIRE(act && out_nnz, "null pointer");
    csr_count_kernel<<<n_rows, 256, 0, (cudaStream_t)stream>>>(act, vocab, out_nnz);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_csr_fill(const float* act, int n_rows, int vocab, const int64_t* row_ptr, int32_t* out_term, float* out_weight,
                fz_stream_t stream) {
    FZ_REQUIRE(n_rows >= 0 && vocab >= 1, "bad shape [%d, %d]", n_rows, vocab);
    if (n_rows == 0) return FZ_OK;
    FZ_REQUIRE(act && row_ptr && out_term && out_weight, "null pointer");
    csr_fill_kernel<<<n_rows, 256, 0, (cudaStream_t)stream>>>(act, vocab, row_ptr, out_term, out_weight);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
