// The tcgen05 threshold-filter GEMM shared by K1 (DPR, dense.cu) and the SPLADE head (splade.cu).
//
// scores[q, n] = <Q[q,:], D[n,:]> for a round of documents, never written to HBM: a warp-specialised persistent
// kernel (TMA producer warp, single-thread tcgen05.mma issuer, two teams of four epilogue warps reading the fp32
// accumulators back from TMEM) filters every score against the query's running threshold and appends the few
// survivors to per-query candidate buffers (topk_state.cuh).
//
// kCodes = false: emit score > tau                                     (DPR: bf16 filter, optional fp32 rescoring)
// kCodes = true : SPLADE head + tail bound.  The accumulator holds only the HEAD part of the sparse dot product (the
//                 terms stored as a dense [N, head_dim] bf16 matrix); the TAIL part arrives as a 4-bit code per
//                 (query, doc) written by tail_codes_kernel (sparse.cu): an upper bound of the tail sum in a tiny
//                 floating-point format whose decoding is one shift and one LOP3.  Emit iff
//                      head * gh + decode(code) > tau * g + B0
//                 i.e. iff the doc's score upper bound can still beat the query's exact running k-th score; the
//                 survivors are rescored exactly in fp32 (sparse_rescore_kernel) before the next select.
#pragma once

#include "common.cuh"
#include "ptx.cuh"
#include "topk_state.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <limits>

namespace fz {

extern void* g_debug_stats;

// 2D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols] (128 bytes wide), 128-byte swizzle (dense.cu)
int make_bf16_tile_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

constexpr int kBM = 128;          // queries per tile  (UMMA M)
constexpr int kBN = 256;          // docs per tile     (UMMA N)
constexpr int kBK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;   // 16 KB
constexpr int kBBytes = kBN * kBK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 384;        // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-7 / 8-11 two epilogue teams
constexpr int kTmemCols = 512;           // two 256-column fp32 accumulators
constexpr int kScratchBytes = 256 * 32 * 4;   // plain mode: one 32-column chunk per epilogue thread (pass 2)
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kScratchBytes;

// ---- 4-bit tail codes (kCodes): decode(c) = as_float(kCodeBase | c << 22) = B0 * 2^(c >> 1) * (1 + (c & 1) / 2),
// B0 = 2^-7.  The bound a code stands for is decode(c) - B0 (in units of 1 / g of the query): 0, .5, 1, 2, 3, 5, 7, 11,
// 15, 23, 31, 47, 63, 95, 127, 191 times 2^-7; the writer rounds up, the largest tail a query can reach maps below 191.
constexpr uint32_t kCodeBase = 0x3C000000u;            // 2^-7: exponent field 120 = 0b01111000, bits 22..25 clear
constexpr float kCodeB0 = 0.0078125f;
constexpr float kCodeTop = 1.45f;                      // a query's largest possible (gained) tail, < decode(15) - B0 = 1.4844
constexpr uint32_t kPendingBit = 0x80000000u;          // candidate id flag: appended by the filter, not yet rescored

struct GemmArgs {
    int n_queries;
    long long r_lo, r_hi;    // doc rows of this round
    int num_k_blocks;
    int m_tiles, n_tiles;
    CandState<float> st;
    unsigned long long* stats;   // optional [gridDim.x, 8] cycle counters (fz_debug_set_stats)
    int debug;                   // FZ_DEBUG_GEMM timing probes (results are WRONG): 1 = nothing passes, 2 = no appends; codes mode: 3 = no filter arithmetic, 4 = no TMEM reads either
    // kCodes only
    const uint4* codes;          // [(256-doc tile of the round) * q_pad + q][8]: the 32 codes of (query q, 32-doc chunk); the 8
                                 // chunks of a tile are one 128-byte line: the tail kernel stores whole lines, the epilogue
                                 // warp reads its 32 queries' lines as one contiguous 4 KB
    int q_pad;                   // n_queries rounded up to kBM
    const float2* qparam;        // [n_queries] (gh, g): head gain incl. the bf16 error allowance, tail gain
};

// FZ_GEMM_2CTA=1 selects the CTA-pair MMA (cta_group::2, M = 256) instead of the default cta_group::1 + multicast
// variant.  Both are parity-green; the pair MMA halves the doc-tile bytes entering each SM but measured no faster
// (DPR 95.9 vs 95.1 ms, SPLADE head 41.7 vs 37.6 ms): the kernel waits on the epilogue, not on operands.
inline bool gemm_two_cta() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("FZ_GEMM_2CTA");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

__device__ __forceinline__ float code_decode(uint32_t word, int nib) {
    const uint32_t sh = nib <= 5 ? (word << (22 - 4 * nib)) : (word >> (4 * nib - 22));
    return __uint_as_float((sh & (15u << 22)) | kCodeBase);
}

// CTA pairs (cluster of 2, adjacent query tiles, same doc tile): each CTA fetches HALF of the doc tile and multicasts it
// to both, so the pair moves 16 + 16 KB per k-block and CTA instead of 16 + 32 KB.  The kernel is bound by L2 -> SM
// operand traffic (profiles/), and 55 query tiles asking L2 for the same doc tile at once also miss together.
constexpr int kPair = 2;
// k2Cta: the pair executes ONE tcgen05.mma.cta_group::2 of M = 256 per k-step instead of two M = 128 MMAs on a multicast B
// tile: each CTA then receives 16 KB (its queries) + 16 KB (its HALF of the doc tile) per k-block instead of 16 + 32 KB.
// The kernel is bound by operand bytes entering the SM (48 KB per 128x256x64 k-block = 126 GB/s per SM at full tensor rate).
template <bool kCodes, bool k2Cta = false>
__global__ void __cluster_dims__(kPair, 1, 1) __launch_bounds__(kGemmThreads, 1)
filter_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                   const GemmArgs G) {
    constexpr int kStages = k2Cta ? 6 : fz::kStages;
    constexpr int kBBytes = k2Cta ? fz::kBBytes / kPair : fz::kBBytes;
    constexpr int kStageBytes = kABytes + kBBytes;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment: required by the 128-byte swizzle atoms the UMMA descriptors assume
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    uint64_t* full_bar = bars;                     // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + kStages;          // [kStages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kStages;      // [2]        MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * kStages + 2; // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float* scratch = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile schedule: the cluster walks (doc tile, pair of query tiles); this CTA takes query tile 2 * pair + rank
    const int crank = (int)ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / kPair, n_clusters = gridDim.x / kPair;
    const int m_pairs = (G.m_tiles + kPair - 1) / kPair;
    const int total_tiles = m_pairs * G.n_tiles;     // per cluster

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_d);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            // multicast: both CTAs' MMAs must have consumed a stage; 2-CTA: the leader's commit is the one arrival
            ptx::mbar_init(&empty_bar[i], k2Cta ? 1 : kPair);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], k2Cta ? 8 : 4);      // 2-CTA: the epilogue warps of BOTH CTAs arrive on the leader's
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (k2Cta) {
            ptx::tmem_alloc_2cta(tmem_slot, kTmemCols);
            ptx::tmem_relinquish_2cta();
        } else {
            ptx::tmem_alloc(tmem_slot, kTmemCols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                 // the partner's barriers exist before anything is multicast to them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer (one elected lane) ================================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            long long st_wait_empty = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int n_t = tile / m_pairs, m_t = (tile - n_t * m_pairs) * kPair + crank;
                const int q0 = m_t * kBM;       // may lie past the last query (odd tile count): TMA zero-fills
                const long long d0 = G.r_lo + (long long)n_t * kBN;
                for (int kb = 0; kb < G.num_k_blocks; ++kb) {
                    const long long t0 = FZ_CLOCK();
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    st_wait_empty += FZ_CLOCK() - t0;
                    unsigned char* sa = smem + (size_t)stage * kStageBytes;
                    if constexpr (k2Cta) {
                        // both CTAs' bytes are credited to the LEADER's barrier: it expects the pair's 2 x (A + half B)
                        if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], kPair * kStageBytes);
                        ptx::tma_load_2d_2cta(sa, &tmap_q, &full_bar[stage], kb * kBK, q0);
                        ptx::tma_load_2d_2cta(sa + kABytes, &tmap_d, &full_bar[stage], kb * kBK, (int32_t)(d0 + crank * (kBN / kPair)));
                    } else {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);     // own queries + both halves of the docs
                        ptx::tma_load_2d(sa, &tmap_q, &full_bar[stage], kb * kBK, q0);
                        ptx::tma_load_2d_multicast(sa + kABytes + crank * (kBBytes / kPair), &tmap_d, &full_bar[stage], kb * kBK,
                                                   (int32_t)(d0 + crank * (kBN / kPair)), (uint16_t)((1u << kPair) - 1));
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (G.stats) G.stats[blockIdx.x * 8 + 0] += (unsigned long long)st_wait_empty;
        }
    } else if (warp == 1 && (!k2Cta || crank == 0)) {
        // ================================ MMA issuer (one elected lane; 2-CTA: of the leader CTA only) =====
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_bf16(k2Cta ? kPair * kBM : kBM, kBN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            long long st_wait_tempty = 0, st_wait_full = 0, st_issue = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
                const int buf = it & 1;
                const long long t0 = FZ_CLOCK();
                ptx::mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
                st_wait_tempty += FZ_CLOCK() - t0;
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * kBN;
                for (int kb = 0; kb < G.num_k_blocks; ++kb) {
                    const long long t1 = FZ_CLOCK();
                    ptx::mbar_wait(&full_bar[stage], phase);
                    const long long t2 = FZ_CLOCK();
                    st_wait_full += t2 - t1;
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * kStageBytes);
                    const uint32_t sb = sa + kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 32);
                        const uint64_t db = ptx::make_smem_desc_sw128(sb + k * 32);
                        if constexpr (k2Cta) ptx::mma_bf16_ss_2cta(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        else ptx::mma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // the stage is reusable once BOTH CTAs' MMAs on it have retired (either producer refills both copies)
                    if constexpr (k2Cta) ptx::mma_commit_2cta(&empty_bar[stage], (uint16_t)((1u << kPair) - 1));
                    else ptx::mma_commit_multicast(&empty_bar[stage], (uint16_t)((1u << kPair) - 1));
                    st_issue += FZ_CLOCK() - t2;
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if constexpr (k2Cta) ptx::mma_commit_2cta(&tfull_bar[buf], (uint16_t)((1u << kPair) - 1));   // in both CTAs
                else ptx::mma_commit(&tfull_bar[buf]);           // accumulator complete
            }
            if (G.stats) {
                G.stats[blockIdx.x * 8 + 1] += (unsigned long long)st_wait_tempty;
                G.stats[blockIdx.x * 8 + 2] += (unsigned long long)st_wait_full;
                G.stats[blockIdx.x * 8 + 3] += (unsigned long long)st_issue;
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue: TMEM -> threshold filter -> candidate append ==========
        // Two teams of four warps: team t owns accumulator buffer t, i.e. every other tile of this CTA, so the candidate
        // appends of one tile (an atomic round trip per thread) overlap the read-back of the next.
        const int ew = (warp - 4) & 3;                      // == warp % 4: the TMEM lane quarter this warp may read
        const int team = (warp - 4) >> 2;
        int it = team;
        long long st_wait_tfull = 0, st_pass2 = 0, st_pass2_n = 0;
        const long long st_begin = FZ_CLOCK();
        // Every global load on this path is issued one tile ahead: under a saturated memory system a demand load
        // takes thousands of cycles and would otherwise sit between "accumulator ready" and "accumulator released".
        // -> (threshold, head gain): plain mode (tau, 1); codes mode (tau * g + B0 nudged down, gh).  A query without k
        // positive-score docs yet (tau <= 0) takes every doc that matches anything: threshold -inf.
        // par_load only LOADS (tau, gh, g) of the next tile's query; par_finish turns them into the threshold at the top of
        // the next iteration - with the arithmetic (and its branch) next to the load, the warp waited out the load's whole
        // latency every tile (9 % of the head GEMM's warp samples, profiles/r02_stall_sampling.md).
        auto par_load = [&](int tile, float& tau_raw, float2& p) {
            tau_raw = std::numeric_limits<float>::infinity();      // marks "no query": nothing passes
            p = make_float2(1.0f, 1.0f);
            if (tile >= total_tiles) return;
            const int n_t = tile / m_pairs, m_t = (tile - n_t * m_pairs) * kPair + crank;
            const int q = m_t * kBM + ew * 32 + lane;
            if (q >= G.n_queries) return;
            tau_raw = G.st.tau[q];
            if constexpr (kCodes) p = G.qparam[q];
        };
        auto par_finish = [&](float tau_raw, float2 p, float& thr, float& gh) {
            gh = 1.0f;
            if constexpr (kCodes) {
                gh = p.x;
                if (tau_raw == std::numeric_limits<float>::infinity()) {
                    thr = tau_raw;
                } else if (tau_raw > 0.0f) {
                    const float t = fmaf(tau_raw, p.y, kCodeB0);
                    // the fma below rounds (2^-24 relative): never lose a borderline doc.  Never below B0: a doc that
                    // shares no term with the query (head 0, code 0; also the zero-filled rows past the last doc)
                    // evaluates to exactly B0 and must not pass.
                    thr = fmaxf(t - t * 3.8e-6f, kCodeB0);
                } else {
                    thr = -std::numeric_limits<float>::infinity();
                }
            } else {
                thr = G.debug == 1 ? std::numeric_limits<float>::infinity() : tau_raw;
            }
        };
        // codes mode: the 8 chunks of (query, 256-doc tile) are one 128-byte line and the lines of a tile's queries are
        // adjacent, so the warp's 32 lines are one contiguous 4 KB.  They are copied global -> shared with cp.async ONE TILE
        // AHEAD (copy c of a lane: row 4 c + lane / 8, 16-byte column lane % 8, parked at column ^ (row & 7): conflict-free
        // for the copy and for the transposed read) - the epilogue is the bottleneck of this kernel, so the accumulator is
        // always ready and a load issued at the top of a tile had its whole latency (thousands of cycles under load) exposed.
        auto issue_codes = [&](int tile) {
            if constexpr (kCodes) {
                if (tile < total_tiles) {
                    const int n_t = tile / m_pairs, m_t = (tile - n_t * m_pairs) * kPair + crank;
                    const uint4* cp = G.codes + ((size_t)n_t * G.q_pad + (size_t)(m_t * kBM + ew * 32)) * (kBN / 32) + lane;
                    uint4* sw = reinterpret_cast<uint4*>(scratch) + (warp - 4) * 256;
#pragma unroll
                    for (int c = 0; c < kBN / 32; ++c) {
                        const int r = 4 * c + (lane >> 3);
                        ptx::cp_async_16(sw + r * 8 + ((lane & 7) ^ (r & 7)), cp + 32 * c);
                    }
                }
                ptx::cp_async_commit();
            }
        };
        issue_codes(cluster_id + team * n_clusters);
        float tau_next;
        float2 p_next;
        par_load(cluster_id + team * n_clusters, tau_next, p_next);
        for (int tile = cluster_id + team * n_clusters; tile < total_tiles; tile += 2 * n_clusters, it += 2) {
            const int n_t = tile / m_pairs, m_t = (tile - n_t * m_pairs) * kPair + crank;
            const int buf = it & 1;
            const int q = m_t * kBM + ew * 32 + lane;
            const long long d0 = G.r_lo + (long long)n_t * kBN;
            const int limit = (int)min((long long)kBN, G.r_hi - d0);
            float tau, gh;
            par_finish(tau_next, p_next, tau, gh);
            par_load(tile + 2 * n_clusters, tau_next, p_next);
            uint4 cw[kCodes ? kBN / 32 : 1];
            const long long t0 = FZ_CLOCK();
            ptx::mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
            st_wait_tfull += FZ_CLOCK() - t0;
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * kBN;
            uint32_t ra[32], rb[32];
            if constexpr (kCodes) {
                {   // this tile's codes were copied into the warp's scratch a whole tile ago (issue_codes); fetch the next
                    // tile's as soon as this one's are in registers
                    ptx::cp_async_wait_all();
                    __syncwarp();
                    const uint4* sw = reinterpret_cast<const uint4*>(scratch) + (warp - 4) * 256;
#pragma unroll
                    for (int c = 0; c < kBN / 32; ++c) cw[c] = sw[lane * 8 + (c ^ (lane & 7))];
                    __syncwarp();
                    issue_codes(tile + 2 * n_clusters);
                }
                // Branch-free: lanes are different queries, so "some lane of the warp has a survivor in this chunk" is the
                // normal case and a warp-level slow path would run almost always.  Every element costs a shift, ONE lop3
                // ((shifted word & field) | exponent; the constants are held in registers - as immediates they cost a second
                // LOP3 on the half-rate INT pipe), one FFMA, and the survivor bit accumulated on the FMA pipe: (x > tau) as
                // 0.0 / 1.0 times 2^j into two fp32 sums of 16 exact bits each.  No max tree, no second pass over TMEM.
                // The chunk loop is ROLLED (two chunks per iteration, register arrays rotated instead of indexed): fully
                // unrolled, the epilogue was 70 KB of code that every warp streamed through once per tile, and a third of
                // all warp samples were instruction-fetch stalls (ncu `stall_no_inst`, profiles/r02d_splade_head_gemm.md).
                uint32_t masks[kBN / 32];
                uint32_t k_field, k_base;
                asm volatile("mov.u32 %0, 0x03C00000;" : "=r"(k_field));
                asm volatile("mov.u32 %0, 0x3C000000;" : "=r"(k_base));
                const bool collect_all = !(tau > -std::numeric_limits<float>::infinity());   // no threshold yet (first rounds)
                auto chunk_mask = [&](const uint32_t(&cur)[32], const uint4 cwc, int c) -> uint32_t {
                    const uint32_t w4[4] = {cwc.x, cwc.y, cwc.z, cwc.w};
                    uint32_t mask = 0;
                    if (G.debug >= 3) return 0u;
                    if (!collect_all) {
                        float lo16 = 0.0f, hi16 = 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int nib = j & 7;
                            const uint32_t word = w4[j >> 3];
                            const uint32_t sh = nib <= 5 ? (word << (22 - 4 * nib)) : (word >> (4 * nib - 22));
                            uint32_t dec;
                            asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(dec) : "r"(sh), "r"(k_field), "r"(k_base));
                            const float x = fmaf(__uint_as_float(cur[j]), gh, __uint_as_float(dec));
                            const float hit = x > tau ? 1.0f : 0.0f;
                            if (j < 16) lo16 = fmaf(hit, (float)(1u << j), lo16);
                            else hi16 = fmaf(hit, (float)(1u << (j - 16)), hi16);
                        }
                        mask = (uint32_t)lo16 | ((uint32_t)hi16 << 16);
                    } else {
                        // every doc of the round that shares a term with the query (head != 0 or code != 0)
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if ((cur[j] != 0u || ((w4[j >> 3] >> (4 * (j & 7))) & 15u) != 0u) && c * 32 + j < limit) mask |= 1u << j;
                    }
                    return mask;
                };
                if (G.debug != 4) ptx::tmem_ld_32x32(t_row, ra);
#pragma unroll 1
                for (int c = 0; c < kBN / 32; c += 2) {
                    if (G.debug != 4) {
                        ptx::tmem_ld_wait(ra);
                        ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
                    }
                    const uint32_t me = chunk_mask(ra, cw[0], c);
                    if (G.debug != 4) {
                        ptx::tmem_ld_wait(rb);
                        if (c + 2 < kBN / 32) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, ra);
                    }
                    const uint32_t mo = chunk_mask(rb, cw[1], c + 1);
                    // rotate: the next iteration finds its codes in cw[0..1]; after the last one masks[] is in chunk order
#pragma unroll
                    for (int i = 0; i + 2 < kBN / 32; ++i) {
                        cw[i] = cw[i + 2];
                        masks[i] = masks[i + 2];
                    }
                    masks[kBN / 32 - 2] = me;
                    masks[kBN / 32 - 1] = mo;
                }
                // The survivors are rescored exactly afterwards, so only their doc ids are recorded: the accumulator is
                // released BEFORE the slot-reserving atomic (its round trip is off the tensor pipe's critical path) and the
                // append walks the set bits of the masks - work proportional to the survivors, not to the tile.
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (k2Cta) ptx::mbar_arrive_cluster(&tempty_bar[buf], 0);
                    else ptx::mbar_arrive(&tempty_bar[buf]);
                }
                int total = 0;
#pragma unroll
                for (int c = 0; c < kBN / 32; ++c) total += __popc(masks[c]);
                if (total > 0 && G.debug != 2) {
                    const long long tp2 = FZ_CLOCK();
                    int base = atomicAdd(&G.st.cnt[q], total);
                    int32_t* ids = G.st.id + (size_t)q * G.st.cap;
#pragma unroll
                    for (int c = 0; c < kBN / 32; ++c) {
                        uint32_t m = masks[c];
                        while (m) {
                            const int j = __ffs(m) - 1;
                            m &= m - 1;
                            if (base < G.st.cap) ids[base] = (int32_t)((uint32_t)(d0 + c * 32 + j) | kPendingBit);
                            ++base;
                        }
                    }
                    st_pass2 += FZ_CLOCK() - tp2;
                    ++st_pass2_n;
                }
                continue;
            }
            // Pass 1: each 32-column chunk is reduced with a max TREE (no dependent chain) while the next chunk's
            // tcgen05.ld is already in flight; only a chunk whose maximum beats tau pays for the compare mask.
            // Both passes are ROLLED loops over chunk pairs (mask registers rotated instead of indexed): see the codes path.
            uint32_t masks[kBN / 32];
            uint32_t flags = 0;
            int total = 0;
            auto chunk_mask = [&](const uint32_t(&cur)[32], int c) -> uint32_t {
                float m8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    m8[j] = fmaxf(fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 8])),
                                  fmaxf(__uint_as_float(cur[j + 16]), __uint_as_float(cur[j + 24])));
                const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                       fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                uint32_t mask = 0;
                if (mx > tau) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (__uint_as_float(cur[j]) > tau && c * 32 + j < limit) mask |= 1u << j;
                }
                return mask;
            };
            ptx::tmem_ld_32x32(t_row, ra);
#pragma unroll 1
            for (int c = 0; c < kBN / 32; c += 2) {
                ptx::tmem_ld_wait(ra);
                ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
                const uint32_t me = chunk_mask(ra, c);
                ptx::tmem_ld_wait(rb);
                if (c + 2 < kBN / 32) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, ra);
                const uint32_t mo = chunk_mask(rb, c + 1);
#pragma unroll
                for (int i = 0; i + 2 < kBN / 32; ++i) masks[i] = masks[i + 2];
                masks[kBN / 32 - 2] = me;
                masks[kBN / 32 - 1] = mo;
                flags |= ((me != 0u ? 1u : 0u) | (mo != 0u ? 2u : 0u)) << c;
                total += __popc(me) + __popc(mo);
            }
            const long long tp2 = FZ_CLOCK();
            // Pass 2: one atomic per thread and tile reserves the slots; the flagged chunks (warp-wide union: tcgen05.ld is
            // warp-collective) are read again, double-buffered.  A lane with survivors in the chunk drops its 32 values
            // into its private shared-memory scratch and walks the SET BITS of its mask, so the work is proportional to the
            // survivors.  (The previous form tested all 32 columns of every flagged chunk behind 32 divergent branches and
            // serialised the re-reads: with the exactness margin most chunks are flagged, and pass 2 took 9k of the 12k
            // cycles a tile's accumulator was held - FZ_KERNEL_STATS, DESIGN.md K1.  Per-chunk atomics were 3x slower.)
            const uint32_t wflags = G.debug == 2 ? 0u : __reduce_or_sync(0xffffffffu, flags);
            if (wflags) {
                int base = total > 0 ? atomicAdd(&G.st.cnt[q], total) : 0;
                const size_t off = (size_t)q * G.st.cap;
                float* scr = scratch + ((int)threadIdx.x - 128) * 4;          // [8][256 threads] float4: conflict-free STS.128
                auto emit = [&](const uint32_t(&cur)[32], uint32_t m, int c) {
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4)
                        *reinterpret_cast<uint4*>(scr + k4 * 1024) = make_uint4(cur[4 * k4], cur[4 * k4 + 1], cur[4 * k4 + 2], cur[4 * k4 + 3]);
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        if (base < G.st.cap) {
                            G.st.score[off + base] = scr[(j >> 2) * 1024 + (j & 3)];
                            G.st.id[off + base] = (int32_t)(d0 + c * 32 + j);
                        }
                        ++base;
                    }
                };
                if (wflags & 1u) ptx::tmem_ld_32x32(t_row, ra);
#pragma unroll 1
                for (int c = 0; c < kBN / 32; c += 2) {
                    const bool h0 = (wflags >> c) & 1u, h1 = (wflags >> (c + 1)) & 1u;
                    if (h0) ptx::tmem_ld_wait(ra);
                    if (h1) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, rb);
                    if (h0 && masks[0] != 0u) emit(ra, masks[0], c);
                    if (h1) ptx::tmem_ld_wait(rb);
                    if (c + 2 < kBN / 32 && ((wflags >> (c + 2)) & 1u)) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, ra);
                    if (h1 && masks[1] != 0u) emit(rb, masks[1], c + 1);
#pragma unroll
                    for (int i = 0; i + 2 < kBN / 32; ++i) masks[i] = masks[i + 2];
                }
                st_pass2 += FZ_CLOCK() - tp2;
                ++st_pass2_n;
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (k2Cta) ptx::mbar_arrive_cluster(&tempty_bar[buf], 0);
                else ptx::mbar_arrive(&tempty_bar[buf]);
            }
        }
        if (G.stats && ew == 0 && lane == 0 && team == 0) {
            G.stats[blockIdx.x * 8 + 4] += (unsigned long long)st_wait_tfull;
            G.stats[blockIdx.x * 8 + 5] += (unsigned long long)(FZ_CLOCK() - st_begin);
            G.stats[blockIdx.x * 8 + 6] += (unsigned long long)st_pass2;
            G.stats[blockIdx.x * 8 + 7] += (unsigned long long)st_pass2_n;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                 // no CTA leaves while its partner may still multicast into it
    if (warp == 2) {
        if constexpr (k2Cta) ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
        else ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace fz
