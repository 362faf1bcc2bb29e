// K2c: SPLADE top-k as HEAD GEMM (tcgen05) + TAIL bound (4-bit codes) + EXACT fp32 rescoring.
//
// The reference scores SPLADE as a dense [Q, V] x [V, N] cosine (src/retrievers/hybrid.py:101-103,
// src/retrievers/splade/base.py:186-251).  On a Zipfian vocabulary > 95 % of all (query term, posting) pairs belong to
// the ~200 most frequent terms: walking those posting lists once per query is an L2-bandwidth problem (round 1:
// 1.6 TB through L2 -> SM per pass), while as a [Q x 192] x [192 x N] bf16 GEMM they cost a few milliseconds of tensor
// time.  The pipeline per round of documents:
//   1. tail_codes_kernel                  the docs' tail postings x the batch's inverted QUERY index -> a 4-bit upper bound
//                                         of the tail sum per (query, doc)
//   2. filter_gemm_kernel<true>           head scores on the tensor cores; epilogue: head * gh + decode(code) > threshold?
//                                         survivors appended to the query's candidate buffer, flagged "pending"
//   3. sparse_rescore_kernel              pending candidates: exact fp32 sparse dot product from the doc-major CSR copy
//   4. cand_select (select.cu)            k-th best EXACT score -> the query's threshold for the next round
// The result holds exact fp32 scores only; the bf16 head and the quantised tail decide nothing but which docs get rescored,
// and both err on the side of keeping a doc (the bound is an upper bound: weights are non-negative).
#include "filter_gemm.cuh"
#include "splade_internal.cuh"

namespace fz {

// ----------------------------------------------------------------------------------- query preparation
// One CTA per query: the head part of the query as a bf16 row of the GEMM's A operand, and the two gains:
//   g  = kCodeTop / (sum over tail terms of w_q * max_d w_d): the largest tail the query can reach maps to the top code
//   gh = g * (1 + c): c covers the bf16 rounding of both operands (2^-8 relative on a sum of non-negative products)
//        and the tensor core's fp32 accumulation, so head * gh >= exact head * g.
constexpr int kPrepThreads = 128;
__global__ void __launch_bounds__(kPrepThreads)
splade_query_prep_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term, const float* __restrict__ q_weight,
                         const int32_t* __restrict__ term_head, const float* __restrict__ term_max, int n_terms, int head_dim,
                         __nv_bfloat16* __restrict__ qh, float2* __restrict__ qparam, int32_t* __restrict__ status) {
    __shared__ float hq[256];
    __shared__ float s_tail[kPrepThreads / 32];
    __shared__ int s_neg;
    const int q = blockIdx.x;
    for (int j = threadIdx.x; j < head_dim; j += kPrepThreads) hq[j] = 0.f;
    if (threadIdx.x == 0) s_neg = 0;
    __syncthreads();
    float tail = 0.f;
    for (int i = q_ptr[q] + threadIdx.x; i < q_ptr[q + 1]; i += kPrepThreads) {
        const int term = q_term[i];
        if (term < 0 || term >= n_terms) continue;
        const float w = q_weight ? q_weight[i] : 1.0f;
        if (!(w >= 0.f)) s_neg = 1;
        const int h = term_head[term];
        if (h >= 0) atomicAdd(&hq[h], w);
        else tail = fmaf(w, term_max[term], tail);
    }
    tail = warp_sum(tail);
    if ((threadIdx.x & 31) == 0) s_tail[threadIdx.x >> 5] = tail;
    __syncthreads();
    for (int j = threadIdx.x; j < head_dim; j += kPrepThreads) qh[(size_t)q * head_dim + j] = __float2bfloat16_rn(hq[j]);
    if (threadIdx.x == 0) {
        float ts = 0.f;
        for (int w = 0; w < kPrepThreads / 32; ++w) ts += s_tail[w];
        float g = ts > 0.f ? kCodeTop / (ts * 1.0002f) : 1.0f;
        g = fminf(g, 1048576.0f);
        const float c = 0.00390625f + 0.00004f + (float)head_dim * 2.4e-7f;
        qparam[q] = make_float2(g * (1.0f + c), g);
        if (s_neg) atomicOr(&status[q], FZ_STATUS_FALLBACK);
    }
}

// ----------------------------------------------------------------------------------- tail codes
// The tail terms' contribution is only needed as an upper bound per (query, doc): a 4-bit code (filter_gemm.cuh).
// One CTA owns (a tile of tail_tile docs) x (a block of 128 queries = one GEMM query tile): its codes live in shared memory
// in exactly the layout the GEMM epilogue reads ([32-doc chunk][query] x 16 bytes).  The tail postings are stored
// TILE-major and term-sorted inside a tile, with a directory [tile][term] -> range; the CTA's work items are the ~1300
// (query, tail term) pairs of its query block, each an independent directory lookup + a walk over the term's few postings
// in the tile.  No loop over tiles, no barrier except the one before the copy-out, fully coalesced code writes.
// (Four query-major versions - one CTA or one warp per (query, doc tile) walking the query's posting lists tile after tile -
// ended between 52 and 110 ms per pass, ~60 % of the warp time in barriers or behind the one dependent posting load of
// each of the ~10 tiny list segments per tile; a doc-major stream against an inverted query index drowned in L2 requests.)
//
// The accumulators ARE the codes: a posting's contribution is rounded up to a code level, and a (query, doc) pair that is
// hit again combines the two codes (smallest level >= the sum of the two levels) with a compare-and-swap on the 32-bit
// word that holds the 8 codes.  88 % of the pairs are never hit, 94 % of the others exactly once.
constexpr int kQB = 128;                        // queries per block (== kBM)
constexpr int kTailThreads = 256;

__device__ __forceinline__ uint32_t tail_code_of(float x) {     // smallest code c with decode(c) >= x (x >= B0)
    return (__float_as_uint(x) - kCodeBase + 0x3FFFFFu) >> 22;
}

// per query: number of tail terms -> (after the scan) start of its (term | query % 128 << 24, weight * gain) entries
__global__ void qe_count_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term,
                                const int32_t* __restrict__ term_head, int n_terms, int n_queries, int32_t* __restrict__ cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const int e = min(q_ptr[q + 1], q_ptr[q] + FZ_MAX_QUERY_TERMS);
    int c = 0;
    for (int i = q_ptr[q]; i < e; ++i) {
        const int term = q_term[i];
        c += (term >= 0 && term < n_terms && term_head[term] < 0) ? 1 : 0;
    }
    cnt[q] = c;
}
// exclusive prefix sum of n (+ 1 total) counters in place, one CTA
__global__ void __launch_bounds__(1024) qe_scan_kernel(int32_t* __restrict__ a, int n) {
    __shared__ int s_part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, per * (int)threadIdx.x), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) {
        const int c = a[i];
        a[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) a[n] = s_part[1023];
}
__global__ void qe_fill_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term, const float* __restrict__ q_weight,
                               const int32_t* __restrict__ term_head, int n_terms, int n_queries, const float2* __restrict__ qparam,
                               const int32_t* __restrict__ start, uint2* __restrict__ ent) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const float g = qparam[q].y;
    const int e = min(q_ptr[q + 1], q_ptr[q] + FZ_MAX_QUERY_TERMS);
    int pos = start[q];
    for (int i = q_ptr[q]; i < e; ++i) {
        const int term = q_term[i];
        if (term >= 0 && term < n_terms && term_head[term] < 0)
            ent[pos++] = make_uint2((uint32_t)term | ((uint32_t)(q % kQB) << 24), __float_as_uint((q_weight ? q_weight[i] : 1.0f) * g));
    }
}

struct TailArgs {
    const int64_t* tile_base;    // [n_tail_tiles + 1] first posting of a tile
    const uint32_t* tile_dir;    // [n_tail_tiles, n_terms + 1] the term's range inside the tile, relative to tile_base
    const uint2* tail_post;      // (doc % tail_tile, weight), tile-major, term-sorted inside a tile
    const int32_t* qe_start;     // [n_queries + 1]
    const uint2* qe_ent;
    int n_terms, n_queries, q_pad, n_qblocks, tail_tile;
    long long r_lo, r_hi_pad;    // round (multiples of 256)
    uint4* codes;                // [(doc - r_lo) / 32][q_pad]
    int32_t* status;
};

__global__ void __launch_bounds__(kTailThreads) tail_codes_kernel(const TailArgs T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* cs = reinterpret_cast<uint32_t*>(smem_raw);            // [tail_tile / 32 chunks][kQB queries][4 words of 8 codes]
    const int n_chunks = T.tail_tile / 32;
    const long long tile0 = T.r_lo / T.tail_tile;                    // (a round may start inside a tail tile)
    const int tile = (int)(tile0 + blockIdx.x / T.n_qblocks), qb = blockIdx.x % T.n_qblocks;
    for (int i = threadIdx.x; i < n_chunks * kQB; i += kTailThreads) reinterpret_cast<uint4*>(cs)[i] = make_uint4(0, 0, 0, 0);
    const int q0 = qb * kQB;
    const int e_lo = T.qe_start[min(q0, T.n_queries)], e_hi = T.qe_start[min(q0 + kQB, T.n_queries)];
    const uint32_t* __restrict__ dir = T.tile_dir + (size_t)tile * (T.n_terms + 1);
    const uint2* __restrict__ post = T.tail_post + T.tile_base[tile];
    __syncthreads();
    int bad_q = -1;
    auto hit = [&](uint32_t dl, uint32_t ql, float x) {     // x = posting weight * query weight * gain, doc dl, query ql
        const uint32_t c = tail_code_of(__fadd_ru(x, kCodeB0));
        if (c == 0u) return;
        if (c > 15u) bad_q = (int)ql;
        uint32_t* w = cs + ((((dl >> 5) * kQB) + ql) << 2) + ((dl >> 3) & 3);
        const int sh = 4 * (dl & 7);
        uint32_t old = *w;
        while (true) {
            const uint32_t nib = (old >> sh) & 15u;
            uint32_t nc = min(c, 15u);
            if (nib != 0u) {                    // second hit of this pair: level(nib) + level(c), rounded up again
                const float a = __fsub_ru(__uint_as_float(kCodeBase | (nib << 22)), kCodeB0);
                nc = tail_code_of(__fadd_ru(a, __uint_as_float(kCodeBase | (nc << 22))));
                if (nc > 15u) { bad_q = (int)ql; nc = 15u; }
            }
            const uint32_t seen = atomicCAS(w, old, (old & ~(15u << sh)) | (nc << sh));
            if (seen == old) break;
            old = seen;
        }
    };
    // Work items: the block's (query, tail term) entries, 32 per warp and step.  The number of postings behind an entry is
    // anything from 0 to a few dozen, so a lane that walked its own entry's postings would idle most of the time (measured:
    // 5-7 active lanes per instruction).  The warp flattens instead: exclusive scan of the 32 range lengths, then the
    // concatenated postings are taken 4 x 32 at a time, every lane locating its (entry, offset) by a 5-step search over
    // the scanned lengths.  The next step's entries and directory lookups are in flight while this step's postings are walked.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kTailThreads / 32;
    constexpr int kUnroll = 4;
    auto load_ent = [&](int e) { return e < e_hi ? __ldg(T.qe_ent + e) : make_uint2(0, 0); };
    int e = e_lo + warp * 32 + lane;
    uint2 x_cur = load_ent(e), x_nxt = load_ent(e + kWarps * 32);
    uint32_t a_cur = 0, b_cur = 0;
    if (e < e_hi) { a_cur = __ldg(dir + (x_cur.x & 0xFFFFFFu)); b_cur = __ldg(dir + (x_cur.x & 0xFFFFFFu) + 1); }
    for (; __any_sync(0xffffffffu, e < e_hi); e += kWarps * 32) {
        const uint2 x = x_cur;
        const uint32_t pa = a_cur;
        const int len = (int)(b_cur - a_cur);
        // prefetch: directory of the next step, entries of the one after
        x_cur = x_nxt;
        x_nxt = load_ent(e + 2 * kWarps * 32);
        a_cur = b_cur = 0;
        if (e + kWarps * 32 < e_hi) { a_cur = __ldg(dir + (x_cur.x & 0xFFFFFFu)); b_cur = __ldg(dir + (x_cur.x & 0xFFFFFFu) + 1); }
        // exclusive scan of the lengths
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - len;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t ql = x.x >> 24;
        const float qw = __uint_as_float(x.y);
        for (int base = 0; base < total; base += 32 * kUnroll) {
            int src[kUnroll];
            uint2 f[kUnroll];
            bool on[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int idx = base + 32 * u + lane;
                on[u] = idx < total;
                // largest lane s with excl[s] <= idx (lengths may be 0: the search lands on the last such lane, whose range
                // then contains idx because excl[s + 1] > idx)
                int s = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int cand = s + step;
                    const int ev = __shfl_sync(0xffffffffu, excl, cand & 31);
                    if (cand < 32 && ev <= idx) s = cand;
                }
                src[u] = s;
                const uint32_t sa = __shfl_sync(0xffffffffu, pa, s);
                const int se = __shfl_sync(0xffffffffu, excl, s);
                f[u] = on[u] ? __ldg(post + sa + (uint32_t)(idx - se)) : make_uint2(0, 0);      // kUnroll loads in flight per lane
                if (base + 32 * (u + 1) >= total) break;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint32_t sq = __shfl_sync(0xffffffffu, ql, src[u]);
                const float sw = __shfl_sync(0xffffffffu, qw, src[u]);
                if (on[u]) hit(f[u].x, sq, __fmul_ru(sw, __uint_as_float(f[u].y)));
                if (base + 32 * (u + 1) >= total) break;
            }
        }
    }
    __syncthreads();
    // ---- codes -> global (the part of the tile inside the round): per 32-doc chunk the block's queries are contiguous
    const long long d_lo = (long long)tile * T.tail_tile;
    const long long c_lo = max(d_lo, T.r_lo), c_hi = min(d_lo + T.tail_tile, T.r_hi_pad);       // multiples of 256
    const int i_lo = (int)((c_lo - d_lo) >> 5), i_hi = (int)((c_hi - d_lo) >> 5);
    const int nq_here = min(kQB, T.q_pad - q0);
    for (int i = i_lo * kQB + threadIdx.x; i < i_hi * kQB; i += kTailThreads) {
        const int c = i / kQB, ql = i - c * kQB;
        if (ql < nq_here) T.codes[(size_t)(((d_lo - T.r_lo) >> 5) + c) * T.q_pad + q0 + ql] = reinterpret_cast<const uint4*>(cs)[i];
    }
    if (bad_q >= 0 && q0 + bad_q < T.n_queries) atomicOr(&T.status[q0 + bad_q], FZ_STATUS_FALLBACK);
}

// ----------------------------------------------------------------------------------- exact rescoring
// One CTA per query.  The query's terms sit in a small open-addressing table in shared memory; a warp takes one pending
// candidate (id flagged by the filter), streams the doc's (term, weight) pairs with coalesced 8-byte loads and sums
// w_q * w_d over the shared terms in fp32.  ~0.75 KB per candidate; a few thousand candidates per query and pass.
constexpr int kRsWarps = 8;
constexpr int kRsSlots = 512;       // >= 4 x FZ_MAX_QUERY_TERMS
__global__ void __launch_bounds__(kRsWarps * 32)
sparse_rescore_kernel(const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_term, const float* __restrict__ q_weight,
                      const int64_t* __restrict__ doc_ptr, const uint2* __restrict__ doc_post, CandState<float> st, int force_all) {
    __shared__ int s_key[kRsSlots];
    __shared__ float s_val[kRsSlots];
    const int q = blockIdx.x;
    const int n = min(st.cnt[q], st.cap);
    const size_t off = (size_t)q * st.cap;
    for (int i = threadIdx.x; i < kRsSlots; i += blockDim.x) { s_key[i] = -1; s_val[i] = 0.f; }
    __syncthreads();
    const int qb = q_ptr[q], qe = min(q_ptr[q + 1], qb + FZ_MAX_QUERY_TERMS);
    for (int i = qb + threadIdx.x; i < qe; i += blockDim.x) {
        const int term = q_term[i];
        if (term < 0) continue;
        const float w = q_weight ? q_weight[i] : 1.0f;
        unsigned slot = ((unsigned)term * 2654435761u) >> 23;      // top 9 bits
        while (true) {
            const int prev = atomicCAS(&s_key[slot], -1, term);
            if (prev == -1 || prev == term) { atomicAdd(&s_val[slot], w); break; }      // a repeated query term adds up
            slot = (slot + 1) & (kRsSlots - 1);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Software pipeline over this warp's candidates: the id of candidate i + 2 and the posting range of candidate i + 1 are
    // in flight while candidate i's postings are summed (each step of the chain id -> doc_ptr -> postings is a dependent
    // global load of ~1 us under load).
    auto load_id = [&](int i) { return i < n ? st.id[off + i] : 0; };
    auto is_pending = [&](int32_t id) { return force_all != 0 || id < 0; };
    int32_t id1 = load_id(warp), id2 = load_id(warp + kRsWarps);
    long long a1 = 0, b1 = 0;
    if (warp < n && is_pending(id1)) {
        const int d = (int)((uint32_t)id1 & ~kPendingBit);
        a1 = __ldg(doc_ptr + d);
        b1 = __ldg(doc_ptr + d + 1);
    }
    for (int i = warp; i < n; i += kRsWarps) {
        const int32_t id = id1;
        const long long p0 = a1, p1 = b1;
        id1 = id2;
        id2 = load_id(i + 2 * kRsWarps);
        if (i + kRsWarps < n && is_pending(id1)) {
            const int d = (int)((uint32_t)id1 & ~kPendingBit);
            a1 = __ldg(doc_ptr + d);
            b1 = __ldg(doc_ptr + d + 1);
        }
        if (!is_pending(id)) continue;                              // rescored in an earlier round
        float acc = 0.f;
        for (long long p = p0 + lane; p < p1; p += 128) {           // up to four independent loads in flight per lane
            uint2 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = p + 32 * u < p1 ? __ldg(doc_post + p + 32 * u) : make_uint2(0xffffffffu, 0u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (e[u].x == 0xffffffffu) continue;
                unsigned slot = (e[u].x * 2654435761u) >> 23;
                while (true) {
                    const int key = s_key[slot];
                    if (key == (int)e[u].x) { acc = fmaf(__uint_as_float(e[u].y), s_val[slot], acc); break; }
                    if (key == -1) break;
                    slot = (slot + 1) & (kRsSlots - 1);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            st.score[off + i] = acc;
            st.id[off + i] = (int32_t)((uint32_t)id & ~kPendingBit);
        }
    }
}

// after the final select: a query whose threshold is not positive has fewer than k positive-score docs (zero-score docs
// would have to fill up in doc-id order: the general path does that)
__global__ void splade_finalize_kernel(CandState<float> st, int n_queries) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_queries && !(st.tau[q] > 0.f)) st.status[q] |= FZ_STATUS_FALLBACK;
}

struct SpladeWs {
    CandState<float> st;
    __nv_bfloat16* qh;
    float2* qparam;
    uint32_t* codes;
    size_t code_bytes;
};

static size_t splade_fixed_bytes(int n_queries, int cap, int head_dim) {
    const size_t q_pad = (size_t)ceil_div(n_queries, kBM) * kBM;
    return align_up(cand_state_bytes<float>(n_queries, cap), 1024) + align_up(q_pad * head_dim * 2, 1024) +
           align_up((size_t)n_queries * sizeof(float2), 1024) + align_up(((size_t)n_queries + 1) * sizeof(int32_t), 1024) +
           align_up((size_t)n_queries * FZ_MAX_QUERY_TERMS * sizeof(uint2), 1024);          // qe_start, qe_ent
}

}  // namespace fz

using namespace fz;

extern "C" {

size_t fz_splade_topk_workspace_bytes(int n_queries, int k, int cap, int head_dim, int64_t max_round_docs) {
    (void)k;
    const size_t q_pad = (size_t)ceil_div(n_queries, kBM) * kBM;
    const size_t tiles = (size_t)ceil_div<long long>(max_round_docs > 256 ? max_round_docs : 256, 256);
    return splade_fixed_bytes(n_queries, cap, head_dim) + tiles * q_pad * 128;
}

int fz_splade_topk(const fz_splade_head_t* head, const fz_postings_t* boot, const int32_t* q_ptr, const int32_t* q_term,
                   const float* q_weight, int n_queries, int k, int64_t doc_base, int cap, int growth, float* out_scores,
                   int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync,
                   fz_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FZ_REQUIRE(head && q_ptr && q_term && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE(head->head_bf16 && head->term_head && head->term_max && head->doc_ptr && head->doc_post && head->tail_base &&
               head->tail_dir, "null head pointer");
    FZ_REQUIRE(head->tail_tile >= 256 && head->tail_tile <= 2048 && head->tail_tile % 256 == 0, "tail_tile=%d must be 256..2048, a multiple of 256",
               head->tail_tile);
    FZ_REQUIRE(head->n_terms >= 1 && head->n_terms < (1 << 24), "n_terms out of range (tail records hold 24-bit term ids)");
    FZ_REQUIRE(head->head_dim >= 64 && head->head_dim <= 256 && head->head_dim % 64 == 0, "head_dim=%d must be 64, 128, 192 or 256",
               head->head_dim);
    FZ_REQUIRE(head->n_docs >= 1 && head->n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192, "need 1 <= k, 2k <= cap <= 8192 (k=%d cap=%d)", k, cap);
    FZ_REQUIRE(growth >= 2 && growth <= 64, "growth=%d out of range", growth);
    if (n_queries == 0) return FZ_OK;
    const int head_dim = head->head_dim;
    const long long N = head->n_docs;
    const int q_pad = ceil_div(n_queries, kBM) * kBM;
    const size_t fixed = splade_fixed_bytes(n_queries, cap, head_dim);
    FZ_REQUIRE(ws && ws_bytes >= fixed + (size_t)q_pad * 128, "workspace too small");
    char* p = (char*)ws;
    CandState<float> st = cand_state_carve<float>(p, n_queries, cap, out_status);
    p += align_up(cand_state_bytes<float>(n_queries, cap), 1024);
    __nv_bfloat16* qh = (__nv_bfloat16*)p;
    p += align_up((size_t)q_pad * head_dim * 2, 1024);
    float2* qparam = (float2*)p;
    p += align_up((size_t)n_queries * sizeof(float2), 1024);
    const int n_qb = ceil_div(n_queries, kQB);
    int32_t* qe_start = (int32_t*)p;
    p += align_up(((size_t)n_queries + 1) * sizeof(int32_t), 1024);
    uint2* qe_ent = (uint2*)p;
    p += align_up((size_t)n_queries * FZ_MAX_QUERY_TERMS * sizeof(uint2), 1024);
    uint32_t* codes = (uint32_t*)p;
    const long long max_round = (long long)((ws_bytes - fixed) / ((size_t)q_pad * 128)) * 256;

    int rc = cand_init<float>(st, n_queries, stream);
    if (rc) return rc;
    {
        ProfScope prof("splade_query_prep", stream);
        splade_query_prep_kernel<<<n_queries, kPrepThreads, 0, stream>>>(q_ptr, q_term, q_weight, head->term_head, head->term_max,
                                                                          head->n_terms, head_dim, qh, qparam, out_status);
        FZ_LAUNCH_CHECK();
        // the batch's (query, tail term) entries, grouped by query (hence by block of kQB queries)
        qe_count_kernel<<<ceil_div(n_queries, 128), 128, 0, stream>>>(q_ptr, q_term, head->term_head, head->n_terms, n_queries, qe_start);
        qe_scan_kernel<<<1, 1024, 0, stream>>>(qe_start, n_queries);
        qe_fill_kernel<<<ceil_div(n_queries, 128), 128, 0, stream>>>(q_ptr, q_term, q_weight, head->term_head, head->n_terms, n_queries,
                                                                      qparam, qe_start, qe_ent);
        FZ_LAUNCH_CHECK();
    }
    CUtensorMap tmap_q, tmap_d;
    rc = make_bf16_tile_map(&tmap_q, qh, (uint64_t)n_queries, (uint64_t)head_dim, kBM);
    if (rc) return rc;
    rc = make_bf16_tile_map(&tmap_d, head->head_bf16, (uint64_t)N, (uint64_t)head_dim, kBN / kPair);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(filter_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        attr = true;
    }
    GemmArgs G;
    memset(&G, 0, sizeof(G));
    G.n_queries = n_queries;
    G.num_k_blocks = head_dim / kBK;
    G.m_tiles = ceil_div(n_queries, kBM);
    G.st = st;
    G.stats = (unsigned long long*)g_debug_stats;
    G.codes = (const uint4*)codes;
    G.q_pad = q_pad;
    G.qparam = qparam;
    TailArgs T;
    memset(&T, 0, sizeof(T));
    T.tile_base = head->tail_base;
    T.tile_dir = head->tail_dir;
    T.tail_post = (const uint2*)head->tail_post;
    T.qe_start = qe_start;
    T.qe_ent = qe_ent;
    T.n_terms = head->n_terms;
    T.n_queries = n_queries;
    T.q_pad = q_pad;
    T.n_qblocks = n_qb;
    T.tail_tile = head->tail_tile;
    T.codes = (uint4*)codes;
    T.status = out_status;
    const size_t tail_smem = (size_t)(head->tail_tile / 32) * kQB * 16;
    FZ_CUDA(cudaFuncSetAttribute(tail_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem));

    const bool synced = sync && sync->hook;
    FZ_REQUIRE(!synced || (sync->exchange && sync->n_shards >= 1 && sync->sched_docs >= N), "bad shard sync");
    const long long SN = synced ? (long long)sync->sched_docs : N;     // the schedule every shard follows
    auto rescore = [&](int force_all) -> int {
        ProfScope prof("splade_rescore", stream);
        sparse_rescore_kernel<<<n_queries, kRsWarps * 32, 0, stream>>>(q_ptr, q_term, q_weight, head->doc_ptr,
                                                                        (const uint2*)head->doc_post, st, force_all);
        FZ_LAUNCH_CHECK();
        return FZ_OK;
    };
    long long lo = 0, hi;
    if (boot) {
        // threshold bootstrap on the shard's first documents with the general kernel (no rescoring per emitted doc), then
        // its survivors are rescored exactly so that every score in the buffer - and the threshold - is an exact one
        FZ_REQUIRE(boot->n_docs >= 256 && boot->n_docs % 256 == 0 && boot->n_docs <= N && boot->n_terms == head->n_terms,
                   "bootstrap index must cover the first n (multiple of 256) docs of the shard");
        rc = sparse_bootstrap_f32(boot, q_ptr, q_term, q_weight, n_queries, k, st, (head->flags & FZ_SPLADE_UNIT_ROWS) != 0, stream);
        if (rc) return rc;
        rc = rescore(1);
        if (rc) return rc;
        rc = cand_select<float>(st, n_queries, k, 0.f, false, doc_base, nullptr, nullptr, nullptr, stream, nullptr);
        if (rc) return rc;
        lo = boot->n_docs;
        hi = lo * growth;
        if (hi - lo > max_round) hi = lo + max_round;
    } else {
        // the first round takes every doc that shares a term with the query (no threshold yet): it must fit the buffer
        hi = (long long)(cap / 256) * 256;
        if (hi > 2048 && 2048 >= 2 * k) hi = 2048;
        if (hi < 256) hi = 256;
        if (hi > max_round) hi = max_round;
        FZ_REQUIRE(hi >= 256 && hi <= cap, "cap=%d too small for a first round of 256 docs", cap);
    }
    if (hi > SN) hi = SN;
    while (true) {
        const long long r_lo = lo < N ? lo : N, r_hi = hi < N ? hi : N;
        if (r_hi > r_lo) {
            {
                T.r_lo = r_lo;
                T.r_hi_pad = (r_hi + 255) / 256 * 256;
                const long long tiles = ceil_div<long long>(r_hi, head->tail_tile) - r_lo / head->tail_tile;
                FZ_REQUIRE(tiles * n_qb < (1ll << 31), "grid too large");
                ProfScope prof("splade_tail_codes", stream);
                tail_codes_kernel<<<(unsigned)(tiles * n_qb), kTailThreads, tail_smem, stream>>>(T);
                FZ_LAUNCH_CHECK();
            }
            G.r_lo = r_lo;
            G.r_hi = r_hi;
            G.n_tiles = (int)ceil_div<long long>(r_hi - r_lo, kBN);
            const long long pair_tiles = (long long)ceil_div(G.m_tiles, kPair) * G.n_tiles;
            const int max_clusters = num_sms() / kPair;
            const int grid = kPair * (int)(pair_tiles < max_clusters ? pair_tiles : max_clusters);
            {
                ProfScope prof("splade_head_gemm", stream);
                filter_gemm_kernel<true><<<grid, kGemmThreads, kGemmSmem, stream>>>(tmap_q, tmap_d, G);
                FZ_LAUNCH_CHECK();
            }
            rc = rescore(0);
            if (rc) return rc;
        }
        const bool last = hi >= SN;
        const float* floor = last ? nullptr : shard_floor<float>(sync, st, n_queries, k, 0.f, stream, &rc);
        if (rc) return rc;
        rc = cand_select<float>(st, n_queries, k, 0.f, last, doc_base, out_scores, out_ids, nullptr, stream, floor);
        if (rc) return rc;
        if (last) break;
        lo = hi;
        hi = hi * growth;
        if (hi - lo > max_round) hi = lo + max_round;
        if (hi > SN) hi = SN;
    }
    splade_finalize_kernel<<<ceil_div(n_queries, 256), 256, 0, stream>>>(st, n_queries);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
