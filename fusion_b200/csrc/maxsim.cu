// K3: ColBERT late interaction (MaxSim) over candidate token tiles on the tcgen05 tensor cores.
//
//   S(q, d) = sum_{i < Lq} max_{j < Ld} <Q_i, D_j>        (colbert-ai `colbert_score`; reference call sites
//   src/utils/colbert_ir.py:245-255, src/retrievers/hybrid.py:109-137)
//
// Layout: the UMMA A operand is the query's token matrix (rows = query tokens -> TMEM lanes), the B operand is a
// candidate document's token block (rows = doc tokens -> TMEM columns), K = 128 embedding dims.  The max over doc
// tokens is then a per-thread running max over the columns each epilogue thread reads back with tcgen05.ld, and the
// sum over query tokens one warp reduction per candidate.  A persistent CTA owns whole queries: warp 0 streams the
// candidates' token rows with TMA (one 2-D box per 64-dim half, box height = the doc length rounded up to 16, so a
// 70-token passage moves 80 rows), warp 1 issues the MMAs, warps 4-7 reduce.  The kernel is HBM-bound: 2*Lq = 128
// FLOP per bf16 element read.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdlib>
#include <limits>

namespace fz {

int make_bf16_tile_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

constexpr int kMsDim = 128;                 // embedding dim (two 64-element swizzle rows)
constexpr int kMsRows = 128;                // UMMA M: query-token rows per tile
constexpr int kMsChunk = 96;                // doc tokens per MMA (UMMA N <= 96 here), longer docs are chunked
constexpr int kMsStages = 6;                // deep ring: ~20 KB per candidate must cover the HBM latency
constexpr int kMsABytes = kMsRows * kMsDim * 2;      // 32 KB per query buffer
constexpr int kMsBBytes = kMsChunk * kMsDim * 2;     // 24 KB per stage
constexpr int kMsTBufs = 4;                           // 4 x 128 TMEM columns
constexpr int kMsTCols = 128;                         // TMEM columns per accumulator buffer
constexpr int kMsThreads = 384;                       // warps 0-3 control, warps 4-7 and 8-11 two epilogue teams
constexpr int kMsBoxes = kMsChunk / 16;               // tensor maps with box heights 16, 32, ..., 96
constexpr size_t kMsSmem = 2 * (size_t)kMsABytes + (size_t)kMsStages * kMsBBytes + 1024 + 256;

struct alignas(64) MsMaps {
    CUtensorMap q;
    CUtensorMap d[kMsBoxes];
};

struct MsArgs {
    const int32_t* cand;        // [n_queries, n_cand] global doc ids
    const int64_t* tok_ptr;     // [n_docs + 1]
    long long n_docs, doc_base;
    int n_queries, n_cand, lq;
    int debug;                  // FZ_MAXSIM_DEBUG: 1 = skip the TMEM read-back, 2 = skip the MMAs (pipeline experiments)
    float* out;                 // [n_queries, n_cand], zero-initialised
};

// (start row, token count) of candidate c0 + lane of query q; len -1 = not in this shard
__device__ __forceinline__ void ms_cand_info(const MsArgs& M, int q, int c0, int lane, int& start, int& len) {
    start = 0;
    len = -1;
    const int c = c0 + lane;
    if (c < M.n_cand) {
        const long long d = (long long)M.cand[(size_t)q * M.n_cand + c] - M.doc_base;
        if (d >= 0 && d < M.n_docs) {
            const long long s = M.tok_ptr[d], e = M.tok_ptr[d + 1];
            start = (int)s;
            len = (int)(e - s);
        }
    }
}

__global__ void __launch_bounds__(kMsThreads, 1) maxsim_kernel(const __grid_constant__ MsMaps maps, const MsArgs M) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_a = smem;                                  // 2 query buffers
    unsigned char* smem_b = smem + 2 * (size_t)kMsABytes;          // kMsStages doc-chunk stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)kMsStages * kMsBBytes);
    uint64_t* full_bar = bars;                          // [kMsStages]
    uint64_t* empty_bar = full_bar + kMsStages;         // [kMsStages]
    uint64_t* afull_bar = empty_bar + kMsStages;        // [2]
    uint64_t* aempty_bar = afull_bar + 2;               // [2]
    uint64_t* tfull_bar = aempty_bar + 2;               // [kMsTBufs]
    uint64_t* tempty_bar = tfull_bar + kMsTBufs;        // [kMsTBufs]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kMsTBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&maps.q);
        for (int i = 0; i < kMsBoxes; ++i) ptx::prefetch_tensormap(&maps.d[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMsStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&afull_bar[i], 1); ptx::mbar_init(&aempty_bar[i], 1); }
        for (int i = 0; i < kMsTBufs; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        int stage = 0;
        uint32_t phase = 0;
        int qi = 0;
        for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x, ++qi) {
            const int abuf = qi & 1;
            if (lane == 0) {
                ptx::mbar_wait(&aempty_bar[abuf], ((qi >> 1) & 1) ^ 1);
                unsigned char* sa = smem_a + (size_t)abuf * kMsABytes;
                ptx::mbar_arrive_expect_tx(&afull_bar[abuf], kMsABytes);
                ptx::tma_load_2d(sa, &maps.q, &afull_bar[abuf], 0, q * M.lq);
                ptx::tma_load_2d(sa + kMsABytes / 2, &maps.q, &afull_bar[abuf], 64, q * M.lq);
            }
            __syncwarp();
            int ns, nl;
            ms_cand_info(M, q, 0, lane, ns, nl);
            for (int c0 = 0; c0 < M.n_cand; c0 += 32) {
                const int cs = ns, cl = nl;
                if (c0 + 32 < M.n_cand) ms_cand_info(M, q, c0 + 32, lane, ns, nl);
                const int nb = min(32, M.n_cand - c0);
                for (int l = 0; l < nb; ++l) {
                    const int start = __shfl_sync(0xffffffffu, cs, l);
                    const int len = __shfl_sync(0xffffffffu, cl, l);
                    for (int off = 0; off < len; off += kMsChunk) {
                        const int n = min(kMsChunk, len - off);
                        const int R = (n + 15) & ~15;
                        if (lane == 0) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                            unsigned char* sb = smem_b + (size_t)stage * kMsBBytes;
                            const CUtensorMap* mp = &maps.d[R / 16 - 1];
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)R * kMsDim * 2);
                            ptx::tma_load_2d(sb, mp, &full_bar[stage], 0, start + off);
                            ptx::tma_load_2d(sb + (size_t)R * 128, mp, &full_bar[stage], 64, start + off);
                        }
                        __syncwarp();
                        if (++stage == kMsStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        int stage = 0;
        uint32_t phase = 0;
        int qi = 0;
        uint32_t t = 0;     // chunk counter -> TMEM buffer
        for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x, ++qi) {
            const int abuf = qi & 1;
            if (lane == 0) {
                ptx::mbar_wait(&afull_bar[abuf], (qi >> 1) & 1);
                ptx::tc_fence_after();
            }
            __syncwarp();
            const uint32_t sa = ptx::smem_u32(smem_a + (size_t)abuf * kMsABytes);
            int ns, nl;
            ms_cand_info(M, q, 0, lane, ns, nl);
            for (int c0 = 0; c0 < M.n_cand; c0 += 32) {
                const int cl = nl;
                if (c0 + 32 < M.n_cand) ms_cand_info(M, q, c0 + 32, lane, ns, nl);
                const int nb = min(32, M.n_cand - c0);
                for (int l = 0; l < nb; ++l) {
                    const int len = __shfl_sync(0xffffffffu, cl, l);
                    for (int off = 0; off < len; off += kMsChunk, ++t) {
                        const int n = min(kMsChunk, len - off);
                        const int R = (n + 15) & ~15;
                        if (lane == 0) {
                            const uint32_t tb = t % kMsTBufs;
                            ptx::mbar_wait(&tempty_bar[tb], ((t / kMsTBufs) & 1) ^ 1);
                            ptx::mbar_wait(&full_bar[stage], phase);
                            ptx::tc_fence_after();
                            const uint32_t sb = ptx::smem_u32(smem_b + (size_t)stage * kMsBBytes);
                            const uint32_t idesc = ptx::make_idesc_bf16(kMsRows, (uint32_t)R);
                            const uint32_t d_tmem = tmem_base + tb * kMsTCols;
#pragma unroll
                            for (int k = 0; k < kMsDim / 16; ++k) {
                                if (M.debug & 2) break;
                                const int half = k >> 2, kk = k & 3;
                                const uint64_t da = ptx::make_smem_desc_sw128(sa + half * (kMsABytes / 2) + kk * 32);
                                const uint64_t db = ptx::make_smem_desc_sw128(sb + half * (R * 128) + kk * 32);
                                ptx::mma_bf16_ss(d_tmem, da, db, idesc, k != 0 ? 1u : 0u);
                            }
                            ptx::mma_commit(&empty_bar[stage]);
                            ptx::mma_commit(&tfull_bar[tb]);
                        }
                        __syncwarp();
                        if (++stage == kMsStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (lane == 0) ptx::mma_commit(&aempty_bar[abuf]);   // query buffer free once its last MMA retires
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================================== epilogue ==========================================
        // Two teams of four warps (TMEM lane quarter = warp % 4) take alternate candidates, so the TMEM read-back
        // and the max/sum reduction of one candidate overlap the next candidate's.
        const int team = (warp - 4) >> 2;
        const int ew = (warp - 4) & 3;
        const int row = ew * 32 + lane;                 // query token handled by this thread
        const bool row_ok = row < M.lq;
        const bool warp_ok = ew * 32 < M.lq;            // warp has at least one live row
        uint32_t t = 0;
        uint32_t cand_no = 0;
        for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x) {
            int ns, nl;
            ms_cand_info(M, q, 0, lane, ns, nl);
            for (int c0 = 0; c0 < M.n_cand; c0 += 32) {
                const int cl = nl;
                if (c0 + 32 < M.n_cand) ms_cand_info(M, q, c0 + 32, lane, ns, nl);
                const int nb = min(32, M.n_cand - c0);
                for (int l = 0; l < nb; ++l) {
                    const int len = __shfl_sync(0xffffffffu, cl, l);
                    if (len < 0) continue;
                    const bool mine = (cand_no++ & 1u) == (uint32_t)team;
                    const int n_chunks = (len + kMsChunk - 1) / kMsChunk;
                    if (!mine) {
                        // stay in phase with every accumulator barrier: a parity wait that skipped a phase would alias
                        for (int c = 0; c < n_chunks; ++c, ++t) ptx::mbar_wait(&tfull_bar[t % kMsTBufs], (t / kMsTBufs) & 1);
                        continue;
                    }
                    float m = -std::numeric_limits<float>::infinity();
                    if (len == 0) m = -9999.f;           // every (padded) doc token is masked to -9999
                    for (int off = 0; off < len; off += kMsChunk, ++t) {
                        const int n = min(kMsChunk, len - off);
                        const uint32_t tb = t % kMsTBufs;
                        ptx::mbar_wait(&tfull_bar[tb], (t / kMsTBufs) & 1);
                        ptx::tc_fence_after();
                        if (warp_ok && !(M.debug & 1)) {
                            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + tb * kMsTCols;
                            uint32_t r[kMsChunk / 16][16];
#pragma unroll
                            for (int g = 0; g < kMsChunk / 16; ++g)
                                if (g * 16 < n) ptx::tmem_ld_32x16(t_row + g * 16, r[g]);     // all loads in flight
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < kMsChunk / 16; ++g)
                                if (g * 16 < n) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (g * 16 + j < n) m = fmaxf(m, __uint_as_float(r[g][j]));
                                }
                        }
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tempty_bar[tb]);
                    }
                    if (warp_ok) {
                        const float s = warp_sum(row_ok ? m : 0.f);
                        if (lane == 0) atomicAdd(&M.out[(size_t)q * M.n_cand + c0 + l], s);
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace fz

using namespace fz;

extern "C" int fz_maxsim_bf16(const void* q_tok, int lq, const int32_t* cand_ids, const int64_t* tok_ptr,
                              const void* tok_emb, int64_t n_tokens, int64_t n_docs, int64_t doc_base, int n_queries,
                              int n_cand, float* out_scores, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(q_tok && cand_ids && tok_ptr && tok_emb && out_scores, "null pointer");
    FZ_REQUIRE(lq >= 1 && lq <= kMsRows, "lq=%d must be in [1,%d]", lq, kMsRows);
    FZ_REQUIRE(n_tokens >= 1 && n_tokens < (1ll << 31), "n_tokens out of range");
    FZ_REQUIRE(n_docs >= 1 && n_cand >= 1, "bad sizes");
    if (n_queries == 0) return FZ_OK;
    FZ_REQUIRE((long long)n_queries * lq < (1ll << 31), "too many query tokens");

    MsMaps maps;
    int rc = make_bf16_tile_map(&maps.q, q_tok, (uint64_t)n_queries * lq, kMsDim, kMsRows);
    if (rc) return rc;
    for (int i = 0; i < kMsBoxes; ++i) {
        rc = make_bf16_tile_map(&maps.d[i], tok_emb, (uint64_t)n_tokens, kMsDim, 16 * (i + 1));
        if (rc) return rc;
    }
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsSmem));
        attr = true;
    }
    FZ_CUDA(cudaMemsetAsync(out_scores, 0, (size_t)n_queries * n_cand * sizeof(float), stream));
    MsArgs M;
    M.cand = cand_ids;
    M.tok_ptr = tok_ptr;
    M.n_docs = n_docs;
    M.doc_base = doc_base;
    M.n_queries = n_queries;
    M.n_cand = n_cand;
    M.lq = lq;
    {
        const char* e = getenv("FZ_MAXSIM_DEBUG");
        M.debug = e ? atoi(e) : 0;
    }
    M.out = out_scores;
    const int grid = n_queries < num_sms() ? n_queries : num_sms();
    ProfScope prof("maxsim", stream);
    maxsim_kernel<<<grid, kMsThreads, kMsSmem, stream>>>(maps, M);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}
