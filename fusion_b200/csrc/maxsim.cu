// K3: ColBERT late interaction (MaxSim) over candidate token tiles on the tcgen05 tensor cores.
//
//   S(q, d) = sum_{i < Lq} max_{j < Ld} <Q_i, D_j>        (colbert-ai `colbert_score`; reference call sites
//   src/utils/colbert_ir.py:245-255, src/retrievers/hybrid.py:109-137)
//
// Layout: the UMMA A operand is the query's token matrix (rows = query tokens -> TMEM lanes), the B operand is a
// GROUP of candidate passages' token blocks packed back to back (rows = doc tokens -> TMEM columns, up to 256 per
// MMA), K = 128 embedding dims.  The max over doc tokens is then a per-thread running max over the columns each
// epilogue thread reads back with tcgen05.ld, and the sum over query tokens one warp reduction per candidate.
//
// The token store is kept in HBM in the exact image the UMMA wants in shared memory (fz_maxsim_pack): per passage the
// two 64-dim halves as 128-byte rows, rows padded to a multiple of 8, 16-byte chunks pre-swizzled (chunk ^ row % 8).
// A passage therefore arrives with two plain bulk copies of exactly its bytes - no tensor map per box height (switching
// between them serialised the TMA unit), no rounding of the row count to the box, and the bytes of a passage are
// contiguous in HBM.  A persistent CTA owns whole queries: warp 0 streams the passages into a 2-stage ring of 64 KB
// groups, warp 1 issues 8 MMAs per group into one of two 256-column TMEM accumulators, the eight epilogue warps reduce
// the passages round-robin in 2 * rep teams (rep = replicas of the query rows in the 128-row A tile, see the kernel).
// Packing ~3 passages per group amortises the barrier round trips, the >= 95-cycle issue cost of a tcgen05.mma and the
// shared-memory reads of the query operand; the packing itself is decided once per query by the pre-pass.  The kernel
// is HBM-bound: 2*Lq = 128 FLOP per bf16 element read.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <limits>

namespace fz {

int make_bf16_tile_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
extern void* g_debug_stats;

constexpr int kMsDim = 128;                 // embedding dim (two 64-element swizzle rows)
constexpr int kMsRows = 128;                // UMMA M: query-token rows per tile
constexpr int kMsGroupMax = 256;            // doc-token rows per MMA group (UMMA N <= 256)
constexpr int kMsStagesDefault = 2;     // doc-group ring depth (FZ_MS_STAGES overrides, 2..8: tuning aid)
constexpr int kMsStagesMax = 8;
constexpr int kMsTBufs = 2;                 // 2 x 256 TMEM columns
constexpr int kMsThreads = 384;             // warps 0-3 control, warps 4-7 and 8-11 two epilogue teams
constexpr size_t kMsSmemMax = 227 * 1024;

struct alignas(64) MsMaps {
    CUtensorMap q;
};

struct MsArgs {
    const int4* info;           // [n_queries, n_cand] owned candidates: (first packed row, token count, candidate, 0)
    const int32_t* count;       // [n_queries] how many of them
    const unsigned char* packed;   // packed token store (fz_maxsim_pack): 256 bytes per packed row
    int n_queries, n_cand, lq;
    int a_rows;                 // query rows per replica (lq rounded up to 8) = TMA box height
    int rep;                    // replicas of the query rows in the 128-row A tile (1, 2 or 4)
    int group_rows;             // doc-token rows per stage (multiple of 16, <= 256)
    int stages;                 // ring depth
    float* out;                 // [n_queries, n_cand], zero-initialised
    unsigned long long* stats;  // optional [gridDim.x, 8] cycle counters (fz_debug_set_stats), NULL = off
};

// Pre-pass, one warp per query.  Under a saturated memory system a demand load takes thousands of cycles, so the
// persistent kernel must not chase cand -> tok_ptr -> tokens pointers on its critical path, and its four roles (one warp
// each, all walking the same candidate list) must not spend their issue slots on bookkeeping either.  Phase 1 gathers
// the candidates this shard owns, compacted in candidate order (a shard of a G-way sharded store owns 1/G of them).
// Phase 2 packs them greedily into MMA groups of <= cap token rows and records the result per candidate:
//     x = first packed row, y = token count, z = candidate index | flags << 24, w = first TMEM column of the passage
// so that the persistent kernel only reads decisions.  A passage longer than one group always starts a group and is cut
// into pieces that each fill a group of their own; the passage after it starts a new group.
constexpr int kMsNewGroup = 1;      // the passage opens a new group (the previous one, if any, ends before it)
constexpr int kMsTeamShift = 1;     // flags bits 1..3: epilogue team that reduces the passage (passages go round-robin)
__global__ void maxsim_info_kernel(const int32_t* __restrict__ cand, const int64_t* __restrict__ tok_ptr,
                                   const int64_t* __restrict__ pk_ptr, long long n_docs, long long doc_base, int n_queries,
                                   int n_cand, int cap, int n_teams, int4* __restrict__ info, int32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n_queries) return;
    int4* __restrict__ mine = info + (size_t)q * n_cand;
    int n_own = 0;
    for (int c0 = 0; c0 < n_cand; c0 += 32) {
        const int c = c0 + lane;
        int4 v = make_int4(0, -1, c, 0);
        if (c < n_cand) {
            const long long d = (long long)cand[(size_t)q * n_cand + c] - doc_base;
            if (d >= 0 && d < n_docs) v = make_int4((int)pk_ptr[d], (int)(tok_ptr[d + 1] - tok_ptr[d]), c, 0);
        }
        const unsigned own = __ballot_sync(0xffffffffu, v.y >= 0);
        if (v.y >= 0) mine[n_own + __popc(own & ((1u << lane) - 1))] = v;
        n_own += __popc(own);
    }
    if (lane == 0) count[q] = n_own;
    __syncwarp();
    int rows = 0, team = 0;
    bool force_new = true;
    for (int c0 = 0; c0 < n_own; c0 += 32) {
        int4 v = c0 + lane < n_own ? mine[c0 + lane] : make_int4(0, -1, 0, 0);
        const int nb = min(32, n_own - c0);
        for (int l = 0; l < nb; ++l) {
            const int len = __shfl_sync(0xffffffffu, v.y, l);
            if (len <= 0) continue;
            int flags = team << kMsTeamShift, col = 0;
            if (++team == n_teams) team = 0;
            if (len > cap) {
                flags |= kMsNewGroup;
                rows = 0;
                force_new = true;
            } else {
                const int R = (len + 7) & ~7;
                if (force_new || rows + R > cap) {
                    flags |= kMsNewGroup;
                    rows = 0;
                    force_new = false;
                }
                col = rows;
                rows += R;
            }
            if (lane == l) {
                v.z |= flags << 24;
                v.w = col;
            }
        }
        if (c0 + lane < n_own) mine[c0 + lane] = v;
    }
}

// Packed image of one passage: [half 0: R rows x 128 B][half 1: R rows x 128 B], R = rows rounded up to 8 (zero rows),
// the 16-byte chunk c of row r stored at chunk c ^ (r % 8) - the 128-byte swizzle of the UMMA K-major layout.
__global__ void maxsim_pack_kernel(const int64_t* __restrict__ tok_ptr, const uint4* __restrict__ emb,
                                   const int64_t* __restrict__ pk_ptr, long long n_docs, uint4* __restrict__ packed) {
    for (long long d = blockIdx.x; d < n_docs; d += gridDim.x) {
        const long long t0 = tok_ptr[d];
        const int len = (int)(tok_ptr[d + 1] - t0);
        const long long p0 = pk_ptr[d];
        const int R = (int)(pk_ptr[d + 1] - p0);
        uint4* dst = packed + (size_t)p0 * 16;                  // 16 uint4 per packed row (256 B)
        for (int i = threadIdx.x; i < R * 16; i += blockDim.x) {
            const int r = i >> 4, c = i & 15, half = c >> 3, cc = c & 7;
            const uint4 v = r < len ? emb[(size_t)(t0 + r) * 16 + c] : make_uint4(0, 0, 0, 0);
            dst[(size_t)half * R * 8 + (size_t)r * 8 + (cc ^ (r & 7))] = v;
        }
    }
}

__device__ __forceinline__ int4 ms_cand_info(const MsArgs& M, int q, int c0, int lane, int n_own) {
    const int c = c0 + lane;
    return c < n_own ? __ldg(&M.info[(size_t)q * M.n_cand + c]) : make_int4(0, -1, 0, 0);
}

// Walk query q's candidates in the packing the pre-pass decided.  Every role (producer, MMA issuer, both epilogue
// teams) runs this same warp-uniform walk; the only state it carries is the row count of the open group.
//   on_piece(c, row0, off, Rdoc, n, R, col, first_of_cand, last_of_cand, first_of_group, team1)
//                                row0 = first packed row of the passage, off = first token of the piece, Rdoc / R =
//                                rows of the passage / piece rounded up to 8, col = first TMEM column of the piece,
//                                team1 = epilogue team the passage belongs to
//   on_group_end(rows)           rows = columns of the group (multiple of 8; the MMA rounds up to 16)
//   on_empty(c)                  candidate with zero tokens
template <class FP, class FG, class FE>
__device__ __forceinline__ void ms_walk(const MsArgs& M, int q, int lane, FP on_piece, FG on_group_end, FE on_empty) {
    int rows = 0;
    const int n_own = __ldg(&M.count[q]);
    // 8 columns stay free at the end of a group: the epilogue's 16 / 32-column tcgen05.ld may read up to 8 columns
    // past a piece
    const int cap = M.group_rows - 8;
    int4 i1 = ms_cand_info(M, q, 0, lane, n_own), i2 = ms_cand_info(M, q, 32, lane, n_own);   // two blocks of 32 in flight
    for (int c0 = 0; c0 < n_own; c0 += 32) {
        const int4 cur = i1;
        i1 = i2;
        i2 = ms_cand_info(M, q, c0 + 64, lane, n_own);
        const int nb = min(32, n_own - c0);
        for (int l = 0; l < nb; ++l) {
            const int start = __shfl_sync(0xffffffffu, cur.x, l);
            const int len = __shfl_sync(0xffffffffu, cur.y, l);
            const int zf = __shfl_sync(0xffffffffu, cur.z, l);
            const int col = __shfl_sync(0xffffffffu, cur.w, l);
            const int cidx = zf & 0xffffff;
            if (len <= 0) {
                if (len == 0) on_empty(cidx);
                continue;
            }
            const int team1 = (zf >> (24 + kMsTeamShift)) & 7;
            const bool new_group = (zf >> 24) & kMsNewGroup;
            const int Rdoc = (len + 7) & ~7;
            const bool is_long = len > cap;             // cut into pieces that are each a group of their own
            for (int off = 0; off < len; off += cap) {  // one trip unless the passage is longer than a group
                const int n = min(cap, len - off);
                const int R = is_long ? (n + 7) & ~7 : Rdoc;
                const int pc = is_long ? 0 : col;
                if ((new_group || off > 0) && rows > 0) {
                    on_group_end(rows);
                    rows = 0;
                }
                on_piece(cidx, start, off, Rdoc, n, R, pc, off == 0, off + n >= len, rows == 0, team1);
                rows = pc + R;
            }
        }
    }
    if (rows > 0) on_group_end(rows);
}

__global__ void __launch_bounds__(kMsThreads, 1) maxsim_kernel(const __grid_constant__ MsMaps maps, const MsArgs M) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int a_half = kMsRows * 128;                               // bytes of one 64-dim half of the query operand
    const int a_bytes = 2 * a_half;
    const int b_half = M.group_rows * 128;
    const int b_bytes = 2 * b_half;                                 // multiple of 4096
    unsigned char* smem_a = smem;                                   // 2 query buffers
    unsigned char* smem_b = smem + 2 * (size_t)a_bytes;             // M.stages doc-group stages
    const int kMsStages = M.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)kMsStages * b_bytes);
    uint64_t* full_bar = bars;                          // [kMsStagesMax]
    uint64_t* empty_bar = full_bar + kMsStagesMax;      // [kMsStagesMax]
    uint64_t* afull_bar = empty_bar + kMsStagesMax;     // [2]
    uint64_t* aempty_bar = afull_bar + 2;               // [2]
    uint64_t* tfull_bar = aempty_bar + 2;               // [kMsTBufs]
    uint64_t* tempty_bar = tfull_bar + kMsTBufs;        // [kMsTBufs]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kMsTBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The UMMA computes 128 query rows whatever lq is.  A query of <= 64 (<= 32) tokens is therefore loaded 2 (4) times
    // into the A tile: the replicas' accumulator rows hold the same scores in different TMEM lane quarters, and every
    // replica gets epilogue teams of its own - all eight epilogue warps reduce passages, not just those of the first
    // lane quarters.
    const int rep = M.rep;                              // 1, 2 or 4 replicas of the query rows
    const int wpt = 4 / rep;                            // epilogue warps (TMEM lane quarters) per team
    const int live_per_team = min(wpt, (M.lq + 31) >> 5);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&maps.q);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMsStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&afull_bar[i], 1); ptx::mbar_init(&aempty_bar[i], 1); }
        // every team passes every accumulator (the owner after reading it), so a waiter can never fall a phase behind
        for (int i = 0; i < kMsTBufs; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 2 * rep * live_per_team); }
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
        long long st_wait_empty = 0;
        int stage = 0;
        uint32_t phase = 0;
        int qi = 0;
        for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x, ++qi) {
            const int abuf = qi & 1;
            if (lane == 0) {
                ptx::mbar_wait(&aempty_bar[abuf], ((qi >> 1) & 1) ^ 1);
                unsigned char* sa = smem_a + (size_t)abuf * a_bytes;
                ptx::mbar_arrive_expect_tx(&afull_bar[abuf], 2 * rep * M.a_rows * 128);
                for (int r = 0; r < rep; ++r) {
                    unsigned char* dst = sa + (size_t)r * (kMsRows / rep) * 128;        // replica r: rows r * 128 / rep ...
                    ptx::tma_load_2d(dst, &maps.q, &afull_bar[abuf], 0, q * M.lq);
                    ptx::tma_load_2d(dst + a_half, &maps.q, &afull_bar[abuf], 64, q * M.lq);
                }
            }
            __syncwarp();
            ms_walk(M, q, lane,
                [&](int, int row0, int off, int Rdoc, int, int R, int col, bool, bool, bool first_of_group, int) {
                    if (lane == 0) {
                        if (first_of_group) {
                            const long long t0 = FZ_CLOCK();
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                            st_wait_empty += FZ_CLOCK() - t0;
                        }
                        unsigned char* sb = smem_b + (size_t)stage * b_bytes + (size_t)col * 128;
                        const unsigned char* src = M.packed + (size_t)row0 * 256 + (size_t)off * 128;
                        ptx::mbar_expect_tx(&full_bar[stage], (uint32_t)R * 256);
                        ptx::bulk_load(sb, src, (uint32_t)R * 128, &full_bar[stage]);                               // dims 0..63
                        ptx::bulk_load(sb + b_half, src + (size_t)Rdoc * 128, (uint32_t)R * 128, &full_bar[stage]);   // 64..127
                    }
                    __syncwarp();
                },
                [&](int) {
                    if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
                    __syncwarp();
                    if (++stage == kMsStages) { stage = 0; phase ^= 1; }
                },
                [&](int) {});
        }
        if (M.stats && lane == 0) M.stats[blockIdx.x * 8 + 0] = (unsigned long long)st_wait_empty;
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        long long st_wait_tempty = 0, st_wait_full = 0, st_issue = 0;
        int stage = 0;
        uint32_t phase = 0;
        int qi = 0;
        uint32_t g = 0;     // group counter -> TMEM buffer
        for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x, ++qi) {
            const int abuf = qi & 1;
            if (lane == 0) {
                ptx::mbar_wait(&afull_bar[abuf], (qi >> 1) & 1);
                ptx::tc_fence_after();
            }
            __syncwarp();
            const uint32_t sa = ptx::smem_u32(smem_a + (size_t)abuf * a_bytes);
            ms_walk(M, q, lane,
                [&](int, int, int, int, int, int, int, bool, bool, bool, int) {},
                [&](int rows) {
                    if (lane == 0) {
                        const uint32_t tb = g % kMsTBufs;
                        const long long t0 = FZ_CLOCK();
                        ptx::mbar_wait(&tempty_bar[tb], ((g / kMsTBufs) & 1) ^ 1);
                        const long long t1 = FZ_CLOCK();
                        ptx::mbar_wait(&full_bar[stage], phase);
                        const long long t2 = FZ_CLOCK();
                        st_wait_tempty += t1 - t0;
                        st_wait_full += t2 - t1;
                        ptx::tc_fence_after();
                        const uint32_t sb = ptx::smem_u32(smem_b + (size_t)stage * b_bytes);
                        const uint32_t idesc = ptx::make_idesc_bf16(kMsRows, (uint32_t)((rows + 15) & ~15));
                        const uint32_t d_tmem = tmem_base + tb * kMsGroupMax;
#pragma unroll
                        for (int k = 0; k < kMsDim / 16; ++k) {
                            const int half = k >> 2, kk = k & 3;
                            const uint64_t da = ptx::make_smem_desc_sw128(sa + half * a_half + kk * 32);
                            const uint64_t db = ptx::make_smem_desc_sw128(sb + half * b_half + kk * 32);
                            ptx::mma_bf16_ss(d_tmem, da, db, idesc, k != 0 ? 1u : 0u);
                        }
                        ptx::mma_commit(&empty_bar[stage]);
                        ptx::mma_commit(&tfull_bar[tb]);
                        st_issue += FZ_CLOCK() - t2;
                    }
                    __syncwarp();
                    ++g;
                    if (++stage == kMsStages) { stage = 0; phase ^= 1; }
                },
                [&](int) {});
            if (lane == 0) ptx::mma_commit(&aempty_bar[abuf]);   // query buffer free once its last MMA retires
            __syncwarp();
        }
        if (M.stats && lane == 0) {
            M.stats[blockIdx.x * 8 + 1] = (unsigned long long)st_wait_tempty;
            M.stats[blockIdx.x * 8 + 2] = (unsigned long long)st_wait_full;
            M.stats[blockIdx.x * 8 + 3] = (unsigned long long)st_issue;
        }
    } else if (warp >= 4) {
        // ===================================== epilogue ==========================================
        // 2 * rep teams (TMEM lane quarter = warp % 4; a team = the wpt quarters of one replica, in one of the two warp
        // sets 4-7 / 8-11) take the passages round-robin, so the TMEM read-back and the max/sum reduction of one passage
        // overlap the others'.  A passage longer than one group keeps its team.
        const int ew = (warp - 4) & 3;                      // TMEM lane quarter
        const int team = ((warp - 4) >> 2) * rep + ew / wpt;
        if (ew % wpt < live_per_team) {
            const int row = (ew % wpt) * 32 + lane;         // query token handled by this thread
            const bool row_ok = row < M.lq;
            uint32_t g = 0;
            float m = -std::numeric_limits<float>::infinity();
            long long st_wait_tfull = 0, st_tmem = 0, st_sum = 0;
            const long long st_begin = FZ_CLOCK();
            for (int q = blockIdx.x; q < M.n_queries; q += gridDim.x) {
                ms_walk(M, q, lane,
                    [&](int c, int, int, int, int n, int, int col, bool, bool last_of_cand, bool first_of_group, int team1) {
                        const uint32_t tb = g % kMsTBufs;
                        if (first_of_group) {
                            const long long t0 = FZ_CLOCK();
                            ptx::mbar_wait(&tfull_bar[tb], (g / kMsTBufs) & 1);
                            st_wait_tfull += FZ_CLOCK() - t0;
                            ptx::tc_fence_after();
                        }
                        if (team1 != team) return;              // passages go round-robin over the teams (pre-pass)
                        const long long tq0 = FZ_CLOCK();
                        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + tb * kMsGroupMax + (uint32_t)col;
                        // 96 columns per step (a whole passage, usually): every tcgen05.ld in flight before the one wait,
                        // four independent max chains.  A 32-column load is used only where the 16-rounded piece covers
                        // it, so no load runs past the accumulator.
                        float m0 = m, m1 = m, m2 = m, m3 = m;
                        auto load = [&](uint32_t (&r)[32], int c, int rem) {
                            if (rem > 16) ptx::tmem_ld_32x32(t_row + c, r); else if (rem > 0) ptx::tmem_ld_32x16_lo(t_row + c, r);
                        };
                        auto reduce = [&](const uint32_t (&r)[32], int rem) {
                            if (rem >= 32) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    m0 = fmaxf(m0, __uint_as_float(r[j]));
                                    m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
                                    m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
                                    m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
                                }
                            } else if (rem > 0) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    if (j < rem) m0 = fmaxf(m0, __uint_as_float(r[j]));
                                    if (j + 1 < rem) m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
                                    if (j + 2 < rem) m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
                                    if (j + 3 < rem) m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
                                }
                            }
                        };
                        for (int c0 = 0; c0 < n; c0 += 96) {
                            uint32_t ra[32], rb[32], rc[32];
                            const int rem = n - c0;
                            load(ra, c0, rem);
                            load(rb, c0 + 32, rem - 32);
                            load(rc, c0 + 64, rem - 64);
                            ptx::tmem_ld_wait(ra);
                            reduce(ra, rem);
                            if (rem > 32) { ptx::tmem_ld_wait(rb); reduce(rb, rem - 32); }
                            if (rem > 64) { ptx::tmem_ld_wait(rc); reduce(rc, rem - 64); }
                        }
                        m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        const long long tq1 = FZ_CLOCK();
                        st_tmem += tq1 - tq0;
                        if (last_of_cand) {
                            const float s = warp_sum(row_ok ? m : 0.f);
                            if (lane == 0) atomicAdd(&M.out[(size_t)q * M.n_cand + c], s);
                            m = -std::numeric_limits<float>::infinity();
                        }
                        st_sum += FZ_CLOCK() - tq1;
                    },
                    [&](int) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tempty_bar[g % kMsTBufs]);
                        ++g;
                    },
                    [&](int c) {
                        // every (padded) doc token is masked to -9999: max = -9999 for each of the lq query tokens
                        if (warp == 4 && lane == 0) atomicAdd(&M.out[(size_t)q * M.n_cand + c], -9999.f * (float)M.lq);
                    });
            }
            if (M.stats && lane == 0) {
                if (warp == 4) {
                    M.stats[blockIdx.x * 8 + 4] = (unsigned long long)st_wait_tfull;
                    M.stats[blockIdx.x * 8 + 5] = (unsigned long long)(FZ_CLOCK() - st_begin);
                    M.stats[blockIdx.x * 8 + 6] = (unsigned long long)st_tmem;
                    M.stats[blockIdx.x * 8 + 7] = (unsigned long long)st_sum;
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

namespace fz { void* g_debug_stats = nullptr; }
// debugging aid: device buffer of [n_ctas, 8] uint64 cycle counters filled by the persistent tensor-core kernels
extern "C" int fz_debug_set_stats(void* device_buffer) {
    fz::g_debug_stats = device_buffer;
    return FZ_OK;
}

extern "C" size_t fz_maxsim_workspace_bytes(int n_queries, int n_cand) {
    return (size_t)n_queries * (size_t)n_cand * sizeof(int4) + align_up((size_t)n_queries * sizeof(int32_t), 16);
}

extern "C" int fz_maxsim_pack(const int64_t* tok_ptr, const void* tok_emb, const int64_t* pk_ptr, int64_t n_docs,
                              void* packed, fz_stream_t stream) {
    FZ_REQUIRE(tok_ptr && tok_emb && pk_ptr && packed && n_docs >= 1, "bad arguments");
    FZ_REQUIRE((((uintptr_t)tok_emb | (uintptr_t)packed) & 15) == 0, "token store must be 16-byte aligned");
    const int blocks = (int)(n_docs < 148 * 64 ? n_docs : 148 * 64);
    maxsim_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tok_ptr, (const uint4*)tok_emb, pk_ptr, n_docs, (uint4*)packed);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

extern "C" int fz_maxsim_bf16(const void* q_tok, int lq, const int32_t* cand_ids, const int64_t* tok_ptr,
                              const int64_t* pk_ptr, const void* packed, int64_t n_docs, int64_t doc_base, int n_queries,
                              int n_cand, float* out_scores, void* ws, size_t ws_bytes, fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(q_tok && cand_ids && tok_ptr && pk_ptr && packed && out_scores, "null pointer");
    FZ_REQUIRE(((uintptr_t)packed & 1023) == 0, "the packed token store must be 1024-byte aligned");
    FZ_REQUIRE(ws && ws_bytes >= fz_maxsim_workspace_bytes(n_queries, n_cand) && ((uintptr_t)ws & 15) == 0, "workspace too small");
    FZ_REQUIRE(lq >= 1 && lq <= kMsRows, "lq=%d must be in [1,%d]", lq, kMsRows);
    FZ_REQUIRE(n_docs >= 1 && n_cand >= 1 && n_cand < (1 << 24), "bad sizes");
    if (n_queries == 0) return FZ_OK;
    FZ_REQUIRE((long long)n_queries * lq < (1ll << 31), "too many query tokens");

    int4* info = (int4*)ws;
    int32_t* count = (int32_t*)((char*)ws + (size_t)n_queries * n_cand * sizeof(int4));
    MsArgs M;
    M.info = info;
    M.count = count;
    M.packed = (const unsigned char*)packed;
    M.n_queries = n_queries;
    M.n_cand = n_cand;
    M.lq = lq;
    M.a_rows = (lq + 7) & ~7;
    M.out = out_scores;
    M.stats = (unsigned long long*)g_debug_stats;
    // shared memory: 2 query buffers + kMsStages doc groups + barriers; the UMMA reads 128 query rows per half, so the
    // 64 KB it may touch past a short query buffer must still be inside the allocation (the stages follow it)
    M.rep = lq <= 32 ? 4 : (lq <= 64 ? 2 : 1);
    const size_t a_bytes = (size_t)2 * kMsRows * 128;
    const size_t fixed = 1024 /*align*/ + 2 * a_bytes + 256 /*barriers*/;
    static int stages_env = -1;
    if (stages_env < 0) {
        const char* e = getenv("FZ_MS_STAGES");
        stages_env = e ? atoi(e) : kMsStagesDefault;
        if (stages_env < 2 || stages_env > kMsStagesMax) stages_env = kMsStagesDefault;
    }
    const int kMsStages = stages_env;
    M.stages = kMsStages;
    int group_rows = (int)((kMsSmemMax - fixed) / kMsStages / 256) & ~15;
    if (group_rows > kMsGroupMax) group_rows = kMsGroupMax;
    M.group_rows = group_rows;
    const size_t smem = fixed + (size_t)kMsStages * group_rows * 256;
    FZ_REQUIRE(group_rows >= 64 && smem <= kMsSmemMax, "shared memory plan failed (lq=%d)", lq);
    {
        ProfScope prof("maxsim_info", stream);
        maxsim_info_kernel<<<ceil_div(n_queries, 8), 256, 0, stream>>>(cand_ids, tok_ptr, pk_ptr, n_docs, doc_base, n_queries,
                                                                       n_cand, group_rows - 8, 2 * M.rep, info, count);
        FZ_LAUNCH_CHECK();
    }

    MsMaps maps;
    int rc = make_bf16_tile_map(&maps.q, q_tok, (uint64_t)n_queries * lq, kMsDim, (uint32_t)M.a_rows);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsSmemMax));
        attr = true;
    }
    FZ_CUDA(cudaMemsetAsync(out_scores, 0, (size_t)n_queries * n_cand * sizeof(float), stream));
    const int grid = n_queries < num_sms() ? n_queries : num_sms();
    ProfScope prof("maxsim", stream);
    maxsim_kernel<<<grid, kMsThreads, smem, stream>>>(maps, M);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}
