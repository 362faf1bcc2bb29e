// K2: BM25 / TF-IDF / ATIRE and SPLADE scoring as sparse accumulation over a term-major CSR inverted index.
//
// The document axis is cut into tiles whose score accumulators live in shared memory.  One CTA scores one
// (doc tile, query) pair: it walks the query's terms IN QUERY ORDER (fp64 lexical scores are summed left to
// right like src/retrievers/bm25.py:152-155; postings of one term hit distinct docs, so there are no atomics and
// the sum is deterministic), then scans the tile and appends every doc that beats the query's running threshold
// to the candidate buffer (topk_state.cuh).  CTAs are ordered tile-major, so all queries stream the same slice of
// the posting lists at the same time and the slice is served from L2 after its first read.
//
// Lexical postings carry the precomputed fp64 impact of the (term, doc) pair — the reference's per-pair formula
// evaluated once at index time with the same operation order and no FMA contraction (fz_lexical_impacts).
#include "common.cuh"
#include "topk_state.cuh"
#include "splade_internal.cuh"

#include <limits>

namespace fz {


// ------------------------------------------------------------------------------------ index-time helpers
__global__ void lexical_impacts_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_doc,
                                       const int32_t* __restrict__ post_tf, const int32_t* __restrict__ doc_len,
                                       const double* __restrict__ idf, int n_terms, long long nnz, double avgdl,
                                       double k1, double k1p1, double one_minus_b, double b, int variant,
                                       double* __restrict__ out) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < nnz;
         p += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_terms;     // last t with term_ptr[t] <= p
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (term_ptr[mid] <= p) lo = mid; else hi = mid;
        }
        const double w = idf[lo];
        const double tf = (double)post_tf[p];
        double v;
        if (variant == FZ_LEX_TFIDF) {
            v = __dmul_rn(tf, w);                                                     // bm25.py:114
        } else {
            const double dl = (double)doc_len[post_doc[p]];
            const double num = __dmul_rn(w, __dmul_rn(tf, k1p1));                     // idf * (tf * (k1 + 1))
            const double inner = __dadd_rn(one_minus_b, __ddiv_rn(__dmul_rn(b, dl), avgdl));   // 1 - b + b*dl/avgdl
            const double den = __dadd_rn(tf, __dmul_rn(k1, inner));                   // tf + k1 * (...)
            v = __ddiv_rn(num, den);                                                  // bm25.py:155
        }
        out[p] = v;
    }
}

// ------------------------------------------------------------------------------------ tile scoring
template <typename AccT>
struct SparseArgs {
    fz_postings_t ix;
    const int32_t* q_ptr;
    const int32_t* q_term;
    const float* q_weight;    // f32 only
    int n_queries;
    int tile_lo, tile_hi;     // tiles of this launch
    int group_lo;             // first tile group of this launch
    long long r_lo, r_hi;     // doc range of this round: the whole tile is accumulated, only this range is emitted
    int sign_mode;            // +1: emit score > 0, -1: emit score < 0
    CandState<AccT> st;
    AccT* out_full;           // full mode: [n_queries, n_docs]
    int accumulate;           // full mode: start from the row's current values (later term chunks of queries > kMaxTerms terms)
    // zero-fill
    int k;
    long long doc_base;
    AccT* out_scores;
    int32_t* out_ids;
};

constexpr int kChunks = 4;      // a thread owns one 16-byte vector of docs (4 fp32 / 2 fp64 sums) in each of 4 chunks
constexpr int kMaxTerms = 128;  // query terms a CTA can hold (longer queries are rejected)
constexpr int kGroupTiles = FZ_COARSE_TILES;    // consecutive doc tiles one CTA walks for its query
constexpr int kMaxSparseThreads = 512;
template <typename AccT> struct AccTraits { static constexpr int kVec = 16 / (int)sizeof(AccT); };

enum { kKindNone = -1, kKindShort = 0, kKindTiled = 1, kKindDense = 2 };

// The query's terms, resolved ONCE per CTA (term id -> storage form, base, weight; short lists: the few postings that
// fall into the CTA's tile group, found through the coarse marks).  Under load a dependent global load costs thousands
// of cycles, so nothing of this chain (q_ptr -> q_term -> term_slot -> base) is repeated per tile.
struct TermStatic {
    long long base[kMaxTerms];   // tiled: first posting of the term; dense: row * stride; short: first posting of the group's slice
    int aux[kMaxTerms];          // tiled: row of the tile-offset table; short: postings in the group's slice
    int kind[kMaxTerms];
    float w[kMaxTerms];
    int n;
};
// The terms with at least one posting in the current tile, dense ones first for fp32 (query order for fp64).
struct TermList {
    long long lo[kMaxTerms];
    int lenkind[kMaxTerms];                  // postings in the tile | storage form << 24
    float w[kMaxTerms];                      // fp32 only
    unsigned ballot[kMaxTerms / 32][2];      // [warp][0 = active dense, 1 = active scatter]
    int n;
};

__device__ __forceinline__ float acc_add(float a, float v, float w) { return __fmaf_rn(v, w, a); }
__device__ __forceinline__ double acc_add(double a, double v, float) { return __dadd_rn(a, v); }

template <typename AccT>
__device__ __forceinline__ void resolve_terms(const SparseArgs<AccT>& A, int q, int group, TermStatic& S) {
    const int t = threadIdx.x;
    const int qb = A.q_ptr[q];
    const int nt = min(A.q_ptr[q + 1] - qb, kMaxTerms);
    if (t < kMaxTerms) {
        int kind = kKindNone, aux = 0;
        long long base = 0;
        float w = 1.0f;
        if (t < nt) {
            const int term = A.q_term[qb + t];
            if (A.q_weight) w = A.q_weight[qb + t];
            if (term >= 0 && term < A.ix.n_terms) {
                const int slot = A.ix.term_slot[term];
                if (slot >= 0) {
                    kind = kKindTiled;
                    base = A.ix.tiled_base[slot];
                    aux = slot;
                } else if (slot <= -2) {
                    kind = kKindDense;
                    base = (long long)(-2 - slot) * A.ix.dense_stride;
                } else {
                    const long long b0 = A.ix.term_ptr[term];
                    if (A.ix.term_ptr[term + 1] > b0) {
                        const uint16_t* cm = A.ix.short_coarse + (size_t)term * (A.ix.n_coarse + 1) + group;
                        const int c0 = cm[0], c1 = cm[1];
                        if (c1 > c0) {
                            kind = kKindShort;
                            base = b0 + c0;
                            aux = c1 - c0;
                        }
                    }
                }
            }
        }
        S.base[t] = base;
        S.aux[t] = aux;
        S.kind[t] = kind;
        S.w[t] = w;
    }
    if (t == 0) {
        S.n = nt;
        // never truncate silently: the top-k entry points flag the query, the caller scores it in chunks (fz_sparse_scores_*)
        if (A.q_ptr[q + 1] - qb > kMaxTerms && A.st.status) atomicOr(&A.st.status[q], FZ_STATUS_TOO_LONG);
    }
}

// Accumulate one query's postings that fall into tile `tile` (docs [d_lo, d_lo + tile_docs)) into acc[0 .. tile_docs).
//
// A thread owns one 16-byte vector of docs in every chunk of kVec * blockDim docs.  Dense terms are added to those docs'
// running sums in REGISTERS with coalesced 16-byte loads; tiled and short terms scatter into shared memory, one term
// after the other: the barrier between terms keeps every doc's sum in query order (fp64 lexical scores are summed left
// to right like src/retrievers/bm25.py:152-155) and makes atomics unnecessary (one term never hits a doc twice).  The
// running sums move between registers and shared memory whenever the storage form of the next term changes, so the
// order of the additions is exactly the query order for fp64; fp32 (SPLADE, order-free within the stated tolerance)
// takes all dense terms first.  Terms with no posting in the tile are dropped when the tile's term list is built.
// (o0, o1) are this thread's term's segment bounds in the tile (tiled terms), loaded one tile ahead by the caller.
// acc needs kVec * kChunks * blockDim + 4 slots; slot tile_docs is the dump slot of the padding postings.
// The final sums of this thread's docs are returned in racc (kToSmem = false: the caller scans them in registers, so a
// tile that ends on a dense term never goes through shared memory) or left in acc behind a barrier (kToSmem = true).
template <typename AccT, bool kToSmem>
__device__ __forceinline__ void accumulate_tile(const SparseArgs<AccT>& A, int tile, long long d_lo, AccT* acc,
                                                const TermStatic& S, TermList& L, uint32_t o0, uint32_t o1,
                                                typename std::conditional<std::is_same<AccT, double>::value, double2, float4>::type (&racc)[kChunks],
                                                const AccT* __restrict__ init = nullptr, int n_init = 0) {
    constexpr int kVec = AccTraits<AccT>::kVec;
    constexpr bool kF64 = std::is_same<AccT, double>::value;
    using Vec = typename std::conditional<kF64, double2, float4>::type;
    using OffVec = typename std::conditional<kF64, uint32_t, uint2>::type;      // kVec uint16 offsets
    const AccT* __restrict__ short_val = reinterpret_cast<const AccT*>(A.ix.post_val);
    const AccT* __restrict__ tiled_val = reinterpret_cast<const AccT*>(A.ix.tiled_val);
    const AccT* __restrict__ dense_val = reinterpret_cast<const AccT*>(A.ix.dense_val);
    const int T = blockDim.x, t = threadIdx.x;
    const int tile_docs = A.ix.tile_docs;
    const int dl = (int)d_lo;

    // ---- the tile's term list: thread i < n_terms holds term i; only the warps that hold terms take part, and a query
    // of <= 32 terms (every lexical query) needs no exchange between warps, so its list costs one barrier per tile
    const int n_tw = S.n <= 32 ? 1 : (S.n + 31) >> 5;
    long long lo = 0;
    int len = 0, kind = kKindNone;
    unsigned bd = 0, bs = 0;
    if (t < 32 * n_tw) {
        kind = S.kind[t];
        if (kind == kKindTiled) {
            lo = S.base[t] + o0;
            len = (int)(o1 - o0);
        } else if (kind == kKindDense) {
            lo = S.base[t] + (long long)tile * tile_docs;
            len = tile_docs;
        } else if (kind == kKindShort) {
            lo = S.base[t];
            len = S.aux[t];         // the group's slice: the scatter checks the tile bounds
        }
        bd = __ballot_sync(0xffffffffu, len > 0 && kind == kKindDense);
        bs = __ballot_sync(0xffffffffu, len > 0 && kind != kKindDense);
        if (n_tw > 1 && (t & 31) == 0) { L.ballot[t >> 5][0] = bd; L.ballot[t >> 5][1] = bs; }
    }
    if (n_tw > 1) __syncthreads();
    if (t < 32 * n_tw) {
        const int wi = t >> 5;
        const unsigned below = (1u << (t & 31)) - 1;
        int dense_before = __popc(bd & below), scat_before = __popc(bs & below), n_dense = __popc(bd), n_scat = __popc(bs);
        if (n_tw > 1) {
            n_dense = 0;
            n_scat = 0;
            for (int w2 = 0; w2 < n_tw; ++w2) {
                const int nd = __popc(L.ballot[w2][0]), ns = __popc(L.ballot[w2][1]);
                if (w2 < wi) { dense_before += nd; scat_before += ns; }
                n_dense += nd;
                n_scat += ns;
            }
        }
        if (len > 0) {
            int pos;
            if (kF64) pos = dense_before + scat_before;                       // query order
            else pos = kind == kKindDense ? dense_before : n_dense + scat_before;
            L.lo[pos] = lo;
            L.lenkind[pos] = len | (kind << 24);
            if (!kF64) L.w[pos] = S.w[t];
        }
        if (t == 0) L.n = n_dense + n_scat;
    }
    __syncthreads();

#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        if constexpr (kF64) racc[c] = make_double2(0.0, 0.0); else racc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (init) {         // continue the sums of an earlier chunk of this query's terms (same left-to-right order)
            AccT v[kVec];
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                const int i = (c * T + t) * kVec + u;
                v[u] = i < n_init ? init[i] : (AccT)0;
            }
            racc[c] = *reinterpret_cast<const Vec*>(v);
        }
    }
    bool in_reg = true;       // the running sums of this thread's docs live in racc (shared memory is stale)

    // (Issuing the next scatter term's global loads before the current term's read-modify-write, and pairing dense rows,
    // were both measured slower: the extra registers cost more occupancy than the shorter dependency chain wins.)
    const int n_active = L.n;
    for (int j = 0; j < n_active; ++j) {
        const int jlk = L.lenkind[j];
        const int jlen = jlk & 0xffffff;
        const int jkind = jlk >> 24;
        const float jw = kF64 ? 1.0f : L.w[j];
        const long long jlo = L.lo[j];
        if (jkind == kKindDense) {
            if (!in_reg) {      // the last scatter ended with a barrier
#pragma unroll
                for (int c = 0; c < kChunks; ++c) racc[c] = *reinterpret_cast<const Vec*>(acc + (size_t)(c * T + t) * kVec);
                in_reg = true;
            }
            const AccT* __restrict__ src = dense_val + jlo;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int i = (c * T + t) * kVec;
                if (i < tile_docs) {
                    const Vec a = __ldg(reinterpret_cast<const Vec*>(src + i));
                    if constexpr (kF64) {
                        racc[c].x = __dadd_rn(racc[c].x, a.x); racc[c].y = __dadd_rn(racc[c].y, a.y);
                    } else {
                        racc[c].x = __fmaf_rn(a.x, jw, racc[c].x); racc[c].y = __fmaf_rn(a.y, jw, racc[c].y);
                        racc[c].z = __fmaf_rn(a.z, jw, racc[c].z); racc[c].w = __fmaf_rn(a.w, jw, racc[c].w);
                    }
                }
            }
            continue;
        }
        if (in_reg) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) *reinterpret_cast<Vec*>(acc + (size_t)(c * T + t) * kVec) = racc[c];
            __syncthreads();
            in_reg = false;
        }
        if (jkind == kKindTiled) {
            // kVec postings per thread and step: one load of uint16 offsets, one 16-byte load of values
            const uint16_t* __restrict__ op = A.ix.tiled_off + jlo;
            const AccT* __restrict__ vp = tiled_val + jlo;
            for (int i = kVec * t; i < jlen; i += kVec * T) {
                const OffVec o = __ldg(reinterpret_cast<const OffVec*>(op + i));
                const Vec v = __ldg(reinterpret_cast<const Vec*>(vp + i));
                if constexpr (kF64) {
                    const unsigned i0 = o & 0xffffu, i1 = o >> 16;
                    acc[i0] = __dadd_rn(acc[i0], v.x);
                    acc[i1] = __dadd_rn(acc[i1], v.y);
                } else {
                    const unsigned i0 = o.x & 0xffffu, i1 = o.x >> 16, i2 = o.y & 0xffffu, i3 = o.y >> 16;
                    acc[i0] = __fmaf_rn(v.x, jw, acc[i0]);
                    acc[i1] = __fmaf_rn(v.y, jw, acc[i1]);
                    acc[i2] = __fmaf_rn(v.z, jw, acc[i2]);
                    acc[i3] = __fmaf_rn(v.w, jw, acc[i3]);
                }
            }
        } else {
            const int32_t* __restrict__ dj = A.ix.post_doc + jlo;
            const AccT* __restrict__ vj = short_val + jlo;
            for (int p = t; p < jlen; p += T) {
                const unsigned o = (unsigned)(__ldg(dj + p) - dl);
                if (o < (unsigned)tile_docs) acc[o] = acc_add(acc[o], __ldg(vj + p), jw);
            }
        }
        __syncthreads();   // the next term may hit the same docs
    }
    if constexpr (kToSmem) {
        if (in_reg) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) *reinterpret_cast<Vec*>(acc + (size_t)(c * T + t) * kVec) = racc[c];
        }
        __syncthreads();
    } else {
        // the caller scans this thread's docs in registers: a tile that ends on a dense term (or holds only dense
        // terms) never touches shared memory again
        if (!in_reg) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) racc[c] = *reinterpret_cast<const Vec*>(acc + (size_t)(c * T + t) * kVec);
        }
    }
}

// segment bounds of this thread's term in `tile` (tiled terms only)
template <typename AccT>
__device__ __forceinline__ uint32_t tile_offset(const SparseArgs<AccT>& A, const TermStatic& S, int tile) {
    const int t = threadIdx.x;
    if (t < max(S.n, 1) && S.kind[t] == kKindTiled)
        return __ldg(A.ix.tiled_tile_off + (size_t)S.aux[t] * (A.ix.n_tiles + 1) + tile);
    return 0;
}

// One CTA = one query x one group of kGroupTiles consecutive doc tiles (CTAs are ordered group-major, so all queries
// stream the same slice of the posting lists at the same time and the slice is served from L2 / L1).
template <typename AccT, int MODE>   // MODE 0: threshold emit, 1: store every score
__global__ void __launch_bounds__(kMaxSparseThreads) sparse_tile_kernel(const SparseArgs<AccT> A) {
    constexpr int kVec = AccTraits<AccT>::kVec;
    using Vec = typename std::conditional<std::is_same<AccT, double>::value, double2, float4>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AccT* acc = reinterpret_cast<AccT*>(smem_raw);
    __shared__ TermStatic S;
    __shared__ TermList L[2];     // double-buffered: a warp may build the next tile's list while others still read this one

    const int group = A.group_lo + blockIdx.x / A.n_queries;
    const int q = blockIdx.x % A.n_queries;
    const int t_begin = max(group * kGroupTiles, A.tile_lo), t_end = min((group + 1) * kGroupTiles, A.tile_hi);
    resolve_terms<AccT>(A, q, group, S);
    const AccT tau = MODE == 0 ? A.st.tau[q] : (AccT)0;
    const AccT lim = A.sign_mode > 0 ? (tau > (AccT)0 ? tau : (AccT)0) : tau;
    __syncthreads();
    uint32_t o0 = tile_offset<AccT>(A, S, t_begin), o1 = tile_offset<AccT>(A, S, t_begin + 1);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const uint32_t o2 = tile + 1 < t_end ? tile_offset<AccT>(A, S, tile + 2) : 0;      // next tile's bound, in flight
        const long long d_lo = (long long)tile * A.ix.tile_docs;
        const long long d_hi = min(d_lo + A.ix.tile_docs, (long long)A.ix.n_docs);
        Vec racc[kChunks];
        const bool cont = MODE == 1 && A.accumulate;
        accumulate_tile<AccT, false>(A, tile, d_lo, acc, S, L[tile & 1], o0, o1, racc,
                                     cont ? A.out_full + (size_t)q * A.ix.n_docs + d_lo : nullptr, (int)(d_hi - d_lo));
        o0 = o1;
        o1 = o2;
        // every thread scans its own docs in registers; survivors are rare once tau has risen
        // (a tile that straddles a round boundary is accumulated by both rounds and emitted once: [e_lo, e_hi))
        const int e_lo = MODE == 1 ? 0 : (int)(max(d_lo, A.r_lo) - d_lo);
        const int e_hi = MODE == 1 ? (int)(d_hi - d_lo) : (int)(min(d_hi, A.r_hi) - d_lo);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int i0 = (c * (int)blockDim.x + (int)threadIdx.x) * kVec;
            AccT v[kVec];
            *reinterpret_cast<Vec*>(v) = racc[c];
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                const AccT sc = v[u];
                const bool inside = i0 + u >= e_lo && i0 + u < e_hi;
                if (MODE == 1) {
                    if (inside) A.out_full[(size_t)q * A.ix.n_docs + d_lo + i0 + u] = sc;
                } else {
                    const bool want = A.sign_mode > 0 ? (sc > lim) : (sc < (AccT)0 && sc > lim);
                    if (want && inside) cand_append<AccT>(A.st, q, sc, (int32_t)(d_lo + i0 + u));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------ EXPERIMENTAL (off by default)
// fp32 variant with FIXED-POINT scatter accumulators (FZ_SPLADE_FIXED=1; not yet validated on hardware - the default path
// is sparse_tile_kernel above).  Why: per (query, 2048-doc tile) a SPLADE query has ~12 active scatter terms that move a
// median of 32 postings each, so the barrier-per-term scheme keeps ~19 % of the CTA's posting slots busy; fp32 atomicAdd
// on shared memory is a CAS loop, but a shared INTEGER add is one native instruction (ATOMS.ADD).  Here the scatter terms
// add round(v * w * 2^26) into int32 accumulators, every warp working through its own terms with no barrier between
// terms (integer adds commute: the result is deterministic), dense rows stay in fp32 registers, and the scan adds the two
// parts.  |scatter sum| must stay below 2^31 / scale = 32 (cos_sim: <= 1); absolute error <= n_terms * 2^-27 (emulated on
// SPLADE-shaped data: 6e-7 relative on the top-1000 scores, fp32 left-to-right: 3e-7).
struct TermListFx {
    long long lo[kMaxTerms];
    int lenkind[kMaxTerms];
    float w[kMaxTerms];
    unsigned ballot[kMaxTerms / 32][2];
    int n, n_dense;
};
constexpr float kFxScale = 67108864.0f;         // 2^26
constexpr float kFxInv = 1.0f / 67108864.0f;

template <int MODE>
__global__ void __launch_bounds__(kMaxSparseThreads) sparse_tile_fx_kernel(const SparseArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* fx = reinterpret_cast<int*>(smem_raw);         // [tile_docs + 4] fixed-point scatter sums, slot tile_docs = dump
    __shared__ TermStatic S;
    __shared__ TermListFx L[2];
    const float* __restrict__ short_val = reinterpret_cast<const float*>(A.ix.post_val);
    const float* __restrict__ tiled_val = reinterpret_cast<const float*>(A.ix.tiled_val);
    const float* __restrict__ dense_val = reinterpret_cast<const float*>(A.ix.dense_val);
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, n_warps = T >> 5;
    const int tile_docs = A.ix.tile_docs;

    const int group = A.group_lo + blockIdx.x / A.n_queries;
    const int q = blockIdx.x % A.n_queries;
    const int t_begin = max(group * kGroupTiles, A.tile_lo), t_end = min((group + 1) * kGroupTiles, A.tile_hi);
    resolve_terms<float>(A, q, group, S);
    const float tau = MODE == 0 ? A.st.tau[q] : 0.0f;
    const float lim = A.sign_mode > 0 ? (tau > 0.0f ? tau : 0.0f) : tau;
    if (t == 0) fx[tile_docs] = 0;
    __syncthreads();
    uint32_t o0 = tile_offset<float>(A, S, t_begin), o1 = tile_offset<float>(A, S, t_begin + 1);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const uint32_t o2 = tile + 1 < t_end ? tile_offset<float>(A, S, tile + 2) : 0;
        const long long d_lo = (long long)tile * tile_docs;
        const long long d_hi = min(d_lo + tile_docs, (long long)A.ix.n_docs);
        const int dl = (int)d_lo;
        TermListFx& Lt = L[tile & 1];
        // ---- zero this thread's own accumulator slots (only it reads them back), then the tile's term list
#pragma unroll
        for (int c = 0; c < kChunks; ++c) *reinterpret_cast<int4*>(fx + (c * T + t) * 4) = make_int4(0, 0, 0, 0);
        const int n_tw = S.n <= 32 ? 1 : (S.n + 31) >> 5;
        long long lo = 0;
        int len = 0, kind = kKindNone;
        unsigned bd = 0, bs = 0;
        if (t < 32 * n_tw) {
            kind = S.kind[t];
            if (kind == kKindTiled) {
                lo = S.base[t] + o0;
                len = (int)(o1 - o0);
            } else if (kind == kKindDense) {
                lo = S.base[t] + (long long)tile * tile_docs;
                len = tile_docs;
            } else if (kind == kKindShort) {
                lo = S.base[t];
                len = S.aux[t];
            }
            bd = __ballot_sync(0xffffffffu, len > 0 && kind == kKindDense);
            bs = __ballot_sync(0xffffffffu, len > 0 && kind != kKindDense);
            if (n_tw > 1 && lane == 0) { Lt.ballot[warp][0] = bd; Lt.ballot[warp][1] = bs; }
        }
        if (n_tw > 1) __syncthreads();
        if (t < 32 * n_tw) {
            const unsigned below = (1u << lane) - 1;
            int dense_before = __popc(bd & below), scat_before = __popc(bs & below), n_dense = __popc(bd), n_scat = __popc(bs);
            if (n_tw > 1) {
                n_dense = 0;
                n_scat = 0;
                for (int w2 = 0; w2 < n_tw; ++w2) {
                    const int nd = __popc(Lt.ballot[w2][0]), ns = __popc(Lt.ballot[w2][1]);
                    if (w2 < warp) { dense_before += nd; scat_before += ns; }
                    n_dense += nd;
                    n_scat += ns;
                }
            }
            if (len > 0) {
                const int pos = kind == kKindDense ? dense_before : n_dense + scat_before;
                Lt.lo[pos] = lo;
                Lt.lenkind[pos] = len | (kind << 24);
                Lt.w[pos] = S.w[t];
            }
            if (t == 0) { Lt.n = n_dense + n_scat; Lt.n_dense = n_dense; }
        }
        __syncthreads();        // list + zeroed accumulators visible
        o0 = o1;
        o1 = o2;

        // ---- dense rows: fp32 registers, every thread its own docs
        float4 racc[kChunks];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) racc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n_active = Lt.n, n_dense = Lt.n_dense;
        for (int j = 0; j < n_dense; ++j) {
            const float jw = Lt.w[j];
            const float* __restrict__ src = dense_val + Lt.lo[j];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int i = (c * T + t) * 4;
                if (i < tile_docs) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
                    racc[c].x = __fmaf_rn(a.x, jw, racc[c].x); racc[c].y = __fmaf_rn(a.y, jw, racc[c].y);
                    racc[c].z = __fmaf_rn(a.z, jw, racc[c].z); racc[c].w = __fmaf_rn(a.w, jw, racc[c].w);
                }
            }
        }
        // ---- scatter terms: warp w takes terms w, w + n_warps, ...; integer atomics, no barrier between terms
        for (int j = n_dense + warp; j < n_active; j += n_warps) {
            const int jlk = Lt.lenkind[j];
            const int jlen = jlk & 0xffffff;
            const float jw = Lt.w[j] * kFxScale;
            const long long jlo = Lt.lo[j];
            if ((jlk >> 24) == kKindTiled) {
                const uint16_t* __restrict__ op = A.ix.tiled_off + jlo;
                const float* __restrict__ vp = tiled_val + jlo;
                for (int i = 4 * lane; i < jlen; i += 128) {
                    const uint2 o = __ldg(reinterpret_cast<const uint2*>(op + i));
                    const float4 v = __ldg(reinterpret_cast<const float4*>(vp + i));
                    atomicAdd(&fx[o.x & 0xffffu], __float2int_rn(v.x * jw));
                    atomicAdd(&fx[o.x >> 16], __float2int_rn(v.y * jw));
                    atomicAdd(&fx[o.y & 0xffffu], __float2int_rn(v.z * jw));
                    atomicAdd(&fx[o.y >> 16], __float2int_rn(v.w * jw));
                }
            } else {
                const int32_t* __restrict__ dj = A.ix.post_doc + jlo;
                const float* __restrict__ vj = short_val + jlo;
                for (int p = lane; p < jlen; p += 32) {
                    const unsigned o = (unsigned)(__ldg(dj + p) - dl);
                    if (o < (unsigned)tile_docs) atomicAdd(&fx[o], __float2int_rn(__ldg(vj + p) * jw));
                }
            }
        }
        __syncthreads();        // every term's adds have landed

        // ---- scan: registers (dense part) + this thread's own fixed-point slots
        const int e_lo = MODE == 1 ? 0 : (int)(max(d_lo, A.r_lo) - d_lo);
        const int e_hi = MODE == 1 ? (int)(d_hi - d_lo) : (int)(min(d_hi, A.r_hi) - d_lo);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int i0 = (c * T + t) * 4;
            const int4 f = *reinterpret_cast<const int4*>(fx + i0);
            const float v[4] = {racc[c].x + (float)f.x * kFxInv, racc[c].y + (float)f.y * kFxInv,
                                racc[c].z + (float)f.z * kFxInv, racc[c].w + (float)f.w * kFxInv};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float sc = v[u];
                const bool inside = i0 + u >= e_lo && i0 + u < e_hi;
                if (MODE == 1) {
                    if (inside) A.out_full[(size_t)q * A.ix.n_docs + d_lo + i0 + u] = sc;
                } else {
                    const bool want = A.sign_mode > 0 ? (sc > lim) : (sc < 0.0f && sc > lim);
                    if (want && inside) cand_append<float>(A.st, q, sc, (int32_t)(d_lo + i0 + u));
                }
            }
        }
    }
}

static bool splade_fixed_point_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("FZ_SPLADE_FIXED");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

template <typename AccT> static int check_index(const fz_postings_t* ix);

// ------------------------------------------------------------------------------------ SPLADE tail codes (splade.cu)
// The SPLADE pipeline scores the HEAD terms (the ~200 most frequent ones: 95+ % of all (query term, posting) pairs) on the
// tensor cores (filter_gemm.cuh) and only needs an UPPER BOUND of the remaining tail sum per (query, doc) to decide which
// docs can still reach the top-k; the survivors are rescored exactly.  This kernel produces that bound: the tail terms'
// postings are scattered into fixed-point shared-memory accumulators (integer atomics: native, commutative, every addend
// rounded up), one warp per term and no barrier between terms, and every doc's sum is rounded up into a 4-bit code
// (filter_gemm.cuh: kCodeBase) written where the GEMM epilogue reads it: 16 bytes per (query, 32-doc chunk).
constexpr uint32_t kTailCodeBase = 0x3C000000u;   // == kCodeBase (filter_gemm.cuh)

constexpr float kTailScale = 67108864.0f;         // 2^26 fixed point of (tail * g) <= kCodeTop

// Round up a fixed-point tail sum into its 4-bit code: smallest c with decode(c) - B0 >= f * 2^-26 (filter_gemm.cuh).
__device__ __forceinline__ uint32_t tail_code(int f, bool& bad) {
    const float tv = __fadd_ru(__int2float_ru(f) * (1.0f / kTailScale), 0.0078125f);
    uint32_t code = (__float_as_uint(tv) - kTailCodeBase + 0x3FFFFFu) >> 22;
    if (code > 15u || f < 0) { bad = true; code = 15u; }
    return code;
}

// The tail terms' postings are scattered into fixed-point shared-memory accumulators (integer atomics: native, commutative,
// every addend rounded up), one warp per term and no barrier between terms.
// 256 threads whatever the tile size: per tile a warp's fixed work (loop set-up, barriers) weighs more than the postings
// it moves, so fewer warps and larger tiles (up to 32768 docs) are what makes this kernel cheap.  The accumulators are
// zeroed once per CTA: the encode pass takes every touched doc's sum with atomicExch(.., 0) (a doc hit by several terms is
// encoded by the first reader, the others see 0 and add nothing), which leaves them zero for the next tile.
constexpr int kTailThreads = 256;
__global__ void __launch_bounds__(kTailThreads) tail_codes_kernel(const TailCodeArgs T, const SparseArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* fx = reinterpret_cast<int*>(smem_raw);         // [tile_docs + 4], slot tile_docs = dump of the padding postings
    uint32_t* cs = reinterpret_cast<uint32_t*>(fx + A.ix.tile_docs + 4);     // [tile_docs / 8] the tile's codes, 8 per word
    __shared__ TermStatic S;
    __shared__ TermListFx L[2];
    const float* __restrict__ short_val = reinterpret_cast<const float*>(A.ix.post_val);
    const float* __restrict__ tiled_val = reinterpret_cast<const float*>(A.ix.tiled_val);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int n_warps = kTailThreads / 32;
    const int tile_docs = A.ix.tile_docs;

    const int group = A.group_lo + blockIdx.x / A.n_queries;
    const int q = blockIdx.x % A.n_queries;
    const int t_begin = max(group * kGroupTiles, A.tile_lo), t_end = min((group + 1) * kGroupTiles, A.tile_hi);
    resolve_terms<float>(A, q, group, S);
    for (int i = t * 4; i < tile_docs + 4; i += kTailThreads * 4) *reinterpret_cast<int4*>(fx + i) = make_int4(0, 0, 0, 0);
    for (int i = t * 4; i < tile_docs / 8; i += kTailThreads * 4) *reinterpret_cast<uint4*>(cs + i) = make_uint4(0, 0, 0, 0);
    const float gain = T.qparam[q].y * kTailScale;
    const long long r_hi_pad = (T.r_hi + 255) / 256 * 256;
    uint4* __restrict__ out4 = reinterpret_cast<uint4*>(T.codes);           // [((doc - r_lo) / 256) * q_pad + q][8] x 16 bytes
    __syncthreads();
    const int n_tw = S.n <= 32 ? 1 : (S.n + 31) >> 5;
    uint32_t o0 = tile_offset<float>(A, S, t_begin), o1 = tile_offset<float>(A, S, t_begin + 1);
    bool bad = false;
    for (int tile = t_begin; tile < t_end; ++tile) {
        TermListFx& Lt = L[tile & 1];
        const long long d_lo = (long long)tile * tile_docs;
        if (warp < n_tw) {      // the warps that hold the query's terms build the tile's list of active terms
            const uint32_t o2 = tile + 1 < t_end ? tile_offset<float>(A, S, tile + 2) : 0;
            long long lo = 0;
            int len = 0;
            const int kind = S.kind[t];
            if (kind == kKindTiled) {
                lo = S.base[t] + o0;
                len = (int)(o1 - o0);
            } else if (kind == kKindShort) {
                lo = S.base[t];
                len = S.aux[t];
            }
            o0 = o1;
            o1 = o2;
            const unsigned bs = __ballot_sync(0xffffffffu, len > 0);
            int before = __popc(bs & ((1u << lane) - 1));
            if (n_tw > 1) {
                // (n_tw <= 4: a named barrier among the term-holding warps only)
                if (lane == 0) Lt.ballot[warp][1] = bs;
                asm volatile("bar.sync 1, %0;" ::"r"(32 * n_tw) : "memory");
                int n_act = 0;
                for (int w2 = 0; w2 < n_tw; ++w2) {
                    const int ns = __popc(Lt.ballot[w2][1]);
                    if (w2 < warp) before += ns;
                    n_act += ns;
                }
                if (t == 0) Lt.n = n_act;
            } else if (t == 0) {
                Lt.n = __popc(bs);
            }
            if (len > 0) {
                Lt.lo[before] = lo;
                Lt.lenkind[before] = len | (kind << 24);
                Lt.w[before] = S.w[t] * gain;
            }
        }
        __syncthreads();        // list visible (accumulators are zero: set-up / the previous tile's encode pass)

        // ---- scatter: warp w takes terms w, w + 8, ...; integer atomics, no barrier between terms
        const int n_active = (T.debug & 2) ? 0 : Lt.n;
        const int dl = (int)d_lo;
        for (int j = warp; j < n_active; j += n_warps) {
            const int jlk = Lt.lenkind[j];
            const int jlen = jlk & 0xffffff;
            const float jw = Lt.w[j];
            const long long jlo = Lt.lo[j];
            if ((jlk >> 24) == kKindTiled) {
                const uint16_t* __restrict__ op = A.ix.tiled_off + jlo;
                const float* __restrict__ vp = tiled_val + jlo;
                for (int i = 4 * lane; i < jlen; i += 128) {
                    const uint2 o = __ldg(reinterpret_cast<const uint2*>(op + i));
                    const float4 v = __ldg(reinterpret_cast<const float4*>(vp + i));
                    atomicAdd(&fx[o.x & 0xffffu], __float2int_ru(v.x * jw));
                    atomicAdd(&fx[o.x >> 16], __float2int_ru(v.y * jw));
                    atomicAdd(&fx[o.y & 0xffffu], __float2int_ru(v.z * jw));
                    atomicAdd(&fx[o.y >> 16], __float2int_ru(v.w * jw));
                }
            } else {
                const int32_t* __restrict__ dj = A.ix.post_doc + jlo;
                const float* __restrict__ vj = short_val + jlo;
                for (int p = lane; p < jlen; p += 32) {
                    const unsigned o = (unsigned)(__ldg(dj + p) - dl);
                    if (o < (unsigned)tile_docs) atomicAdd(&fx[o], __float2int_ru(__ldg(vj + p) * jw));
                }
            }
        }
        __syncthreads();        // every term's adds have landed

        // ---- encode only the docs a posting touched (~10 % of the tile): walk the postings again (their offsets come
        // from L1 now), take each touched doc's sum (leaving 0), round it up into its code, OR it into the code words
        for (int j = warp; j < n_active; j += n_warps) {
            const int jlk = Lt.lenkind[j];
            const int jlen = jlk & 0xffffff;
            const long long jlo = Lt.lo[j];
            if ((jlk >> 24) == kKindTiled) {
                const uint16_t* __restrict__ op = A.ix.tiled_off + jlo;
                for (int i = 4 * lane; i < jlen; i += 128) {
                    const uint2 o = __ldg(reinterpret_cast<const uint2*>(op + i));
                    const unsigned oo[4] = {o.x & 0xffffu, o.x >> 16, o.y & 0xffffu, o.y >> 16};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (oo[u] >= (unsigned)tile_docs) continue;
                        const int f = atomicExch(&fx[oo[u]], 0);
                        if (f != 0) atomicOr(&cs[oo[u] >> 3], tail_code(f, bad) << (4 * (oo[u] & 7)));
                    }
                }
            } else {
                const int32_t* __restrict__ dj = A.ix.post_doc + jlo;
                for (int p = lane; p < jlen; p += 32) {
                    const unsigned o = (unsigned)(__ldg(dj + p) - dl);
                    if (o >= (unsigned)tile_docs) continue;
                    const int f = atomicExch(&fx[o], 0);
                    if (f != 0) atomicOr(&cs[o >> 3], tail_code(f, bad) << (4 * (o & 7)));
                }
            }
        }
        __syncthreads();

        // ---- the tile's code words -> the round's code buffer [256-doc tile][query][8 chunks]: 8 consecutive threads store
        // one whole 128-byte line (a [chunk][query] layout cost 1.9 G scattered 16-byte partial-sector stores per step: 10
        // of this kernel's 46 ms); the words are zeroed again by the thread that copied them
        {
            const long long c_lo = max(d_lo, T.r_lo), c_hi = min(d_lo + tile_docs, r_hi_pad);      // multiples of 256
            const int i_lo = (int)((c_lo - d_lo) >> 5), i_hi = (int)((c_hi - d_lo) >> 5);
            const long long rel = (d_lo - T.r_lo) >> 5;          // chunk index of the tile's first doc in the round (may be < 0)
            for (int i = i_lo + t; i < i_hi; i += kTailThreads) {
                uint4* src = reinterpret_cast<uint4*>(cs) + i;
                const long long ch = rel + i;
                if (!(T.debug & 1)) out4[((ch >> 3) * T.q_pad + q) * 8 + (ch & 7)] = *src;
                *src = make_uint4(0, 0, 0, 0);
            }
        }
    }
    if (bad) atomicOr(&T.status[q], FZ_STATUS_FALLBACK);
}

int launch_tail_codes(const TailCodeArgs& T, cudaStream_t stream) {
    const fz_postings_t* ix = &T.ix;
    FZ_REQUIRE(ix->term_ptr && ix->term_slot && ix->short_coarse, "null index pointer");
    FZ_REQUIRE(ix->n_tiled == 0 || (ix->tiled_base && ix->tiled_tile_off && ix->tiled_off && ix->tiled_val), "tiled postings missing");
    FZ_REQUIRE(ix->tile_docs >= 256 && ix->tile_docs <= 32768 && ix->tile_docs % 256 == 0 && ix->n_dense == 0,
               "tail index: tile_docs must be a multiple of 256 in [256, 32768], no dense rows");
    FZ_REQUIRE(ix->n_docs >= 1 && ix->n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(ix->n_tiles == (int)ceil_div<long long>(ix->n_docs, ix->tile_docs), "n_tiles inconsistent with n_docs");
    FZ_REQUIRE(ix->n_coarse == (ix->n_tiles + FZ_COARSE_TILES - 1) / FZ_COARSE_TILES, "n_coarse inconsistent with n_tiles");
    FZ_REQUIRE(T.r_lo % 256 == 0 && T.r_hi > T.r_lo, "tail round must start on a multiple of 256");
    const size_t smem = ((size_t)ix->tile_docs + 4) * sizeof(int) + (size_t)ix->tile_docs / 2;
    static bool attr = false;
    if (!attr) {
        FZ_CUDA(cudaFuncSetAttribute(tail_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr = true;
    }
    SparseArgs<float> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = T.q_ptr;
    A.q_term = T.q_term;
    A.q_weight = T.q_weight;
    A.n_queries = T.n_queries;
    A.tile_lo = (int)(T.r_lo / ix->tile_docs);
    A.tile_hi = (int)ceil_div<long long>(T.r_hi < ix->n_docs ? T.r_hi : ix->n_docs, ix->tile_docs);
    if (A.tile_hi <= A.tile_lo) A.tile_hi = A.tile_lo + 1;
    A.group_lo = A.tile_lo / kGroupTiles;
    const long long blocks = (long long)(ceil_div(A.tile_hi, kGroupTiles) - A.group_lo) * T.n_queries;
    FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
    ProfScope prof("splade_tail_codes", stream);
    tail_codes_kernel<<<(unsigned)blocks, kTailThreads, smem, stream>>>(T, A);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

// Queries with fewer than k positive-score docs: append zero-score docs in ascending doc-id order
// (the reference ranks every document; unmatched ones score exactly 0.0 and tie by index, bm25.py:103-105).
template <typename AccT>
__global__ void __launch_bounds__(kMaxSparseThreads) sparse_zero_fill_kernel(const SparseArgs<AccT> A) {
    using Vec = typename std::conditional<std::is_same<AccT, double>::value, double2, float4>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AccT* acc = reinterpret_cast<AccT*>(smem_raw);
    __shared__ TermStatic S;
    __shared__ TermList L;
    __shared__ int s_warp[kMaxSparseThreads / 32];
    __shared__ int s_found;

    const int q = blockIdx.x;
    // after the final select cnt[q] is the number of positive-score docs kept: all of them when there are < k
    const int n_have = A.st.cnt[q];
    // negatives never need zero fill; a positive threshold means k positive-score docs exist (on this shard, or - with a
    // cross-shard floor - over all shards), so no zero-score doc can reach the top-k
    if (n_have >= A.k || A.sign_mode < 0 || A.st.tau[q] > (AccT)0) return;
    const int need = A.k - n_have;
    const int n_warps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        s_found = 0;
        A.st.status[q] |= FZ_STATUS_NEED_ZERO;
    }
    __syncthreads();
    for (int tile = 0; tile < A.ix.n_tiles; ++tile) {
        if (tile % kGroupTiles == 0) {
            __syncthreads();
            resolve_terms<AccT>(A, q, tile / kGroupTiles, S);
            __syncthreads();
        }
        const long long d_lo = (long long)tile * A.ix.tile_docs;
        const long long d_hi = min(d_lo + A.ix.tile_docs, (long long)A.ix.n_docs);
        Vec racc[kChunks];
        accumulate_tile<AccT, true>(A, tile, d_lo, acc, S, L, tile_offset<AccT>(A, S, tile), tile_offset<AccT>(A, S, tile + 1), racc);
        const int n = (int)(d_hi - d_lo);
        for (int c0 = 0; c0 < n; c0 += blockDim.x) {
            const int i = c0 + threadIdx.x;
            const int z = (i < n) && (acc[i] == (AccT)0);
            const unsigned bal = __ballot_sync(0xffffffffu, z);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            __syncthreads();
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < n_warps; ++w) {
                const int v = s_warp[w];
                if (w < warp) before += v;
                total += v;
            }
            const int base = s_found;
            if (z) {
                const int r = base + before + __popc(bal & ((1u << lane) - 1));
                if (r < need) {
                    A.out_scores[(size_t)q * A.k + n_have + r] = (AccT)0;
                    A.out_ids[(size_t)q * A.k + n_have + r] = (int32_t)(A.doc_base + d_lo + i);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_found = base + total;
            __syncthreads();
            if (s_found >= need) return;
        }
    }
    if (threadIdx.x == 0 && s_found < need) A.st.status[q] |= FZ_STATUS_NEED_NEG;
}

// threads per CTA: every thread owns 4 docs in each of kChunks chunks
template <typename AccT>
static int sparse_threads(int tile_docs) {
    const int per_thread = AccTraits<AccT>::kVec * kChunks;
    int t = (tile_docs + per_thread - 1) / per_thread;
    t = (t + 31) / 32 * 32;
    return t < kMaxTerms ? kMaxTerms : t;       // kMaxTerms threads resolve the terms
}
template <typename AccT>
static size_t sparse_smem(int tile_docs) {
    return ((size_t)AccTraits<AccT>::kVec * kChunks * sparse_threads<AccT>(tile_docs) + 4) * sizeof(AccT);
}

template <typename AccT>
static int check_index(const fz_postings_t* ix) {
    FZ_REQUIRE(ix && ix->term_ptr && ix->term_slot && ix->short_coarse, "null index pointer");
    FZ_REQUIRE(ix->n_coarse == (ix->n_tiles + FZ_COARSE_TILES - 1) / FZ_COARSE_TILES, "n_coarse inconsistent with n_tiles");
    FZ_REQUIRE(ix->n_tiled == 0 || (ix->tiled_base && ix->tiled_tile_off && ix->tiled_off && ix->tiled_val), "tiled postings missing");
    FZ_REQUIRE(ix->n_dense == 0 || ix->dense_val, "dense rows missing");
    FZ_REQUIRE(ix->tile_docs >= 4 && ix->tile_docs % 4 == 0, "tile_docs=%d must be a positive multiple of 4", ix->tile_docs);
    FZ_REQUIRE(sparse_threads<AccT>(ix->tile_docs) <= kMaxSparseThreads, "tile_docs=%d too large (max %d for this score type)",
               ix->tile_docs, AccTraits<AccT>::kVec * kChunks * kMaxSparseThreads);
    FZ_REQUIRE(ix->n_docs >= 1 && ix->n_docs < (1ll << 31), "n_docs out of range");
    FZ_REQUIRE(ix->n_tiles == (int)ceil_div<long long>(ix->n_docs, ix->tile_docs), "n_tiles inconsistent with n_docs");
    FZ_REQUIRE(ix->dense_stride == (long long)ix->n_tiles * ix->tile_docs, "dense_stride must be n_tiles * tile_docs");
    return FZ_OK;
}

template <typename AccT>
static int set_smem_attrs() {
    static bool done = false;
    if (!done) {
        FZ_CUDA(cudaFuncSetAttribute(sparse_tile_kernel<AccT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(sparse_tile_kernel<AccT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FZ_CUDA(cudaFuncSetAttribute(sparse_zero_fill_kernel<AccT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (std::is_same<AccT, float>::value)
            FZ_CUDA(cudaFuncSetAttribute(sparse_tile_fx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    return FZ_OK;
}

template <typename AccT>
static int sparse_topk(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                       int n_queries, int k, int64_t doc_base, int cap, int growth, int sign_mode, AccT* out_scores,
                       int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync,
                       cudaStream_t stream) {
    int rc = check_index<AccT>(ix);
    if (rc) return rc;
    FZ_REQUIRE(q_ptr && q_term && out_scores && out_ids && out_status, "null pointer");
    FZ_REQUIRE(k >= 1 && cap >= 2 * k && cap <= 8192, "need 1 <= k, 2k <= cap <= 8192 (k=%d cap=%d)", k, cap);
    FZ_REQUIRE(growth >= 1 && growth <= 64, "growth=%d out of range", growth);
    FZ_REQUIRE(sign_mode == 1 || sign_mode == -1, "sign_mode must be +1 or -1");
    FZ_REQUIRE(ws && ws_bytes >= cand_state_bytes<AccT>(n_queries, cap), "workspace too small");
    if (n_queries == 0) return FZ_OK;
    rc = set_smem_attrs<AccT>();
    if (rc) return rc;

    SparseArgs<AccT> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = q_ptr;
    A.q_term = q_term;
    A.q_weight = q_weight;
    A.n_queries = n_queries;
    A.sign_mode = sign_mode;
    A.st = cand_state_carve<AccT>(ws, n_queries, cap, out_status);
    A.k = k;
    A.doc_base = doc_base;
    A.out_scores = out_scores;
    A.out_ids = out_ids;
    rc = cand_init<AccT>(A.st, n_queries, stream);
    if (rc) return rc;

    const int threads = sparse_threads<AccT>(ix->tile_docs);
    const size_t smem = sparse_smem<AccT>(ix->tile_docs);
    const long long N = ix->n_docs;
    const bool synced = sync && sync->hook;
    FZ_REQUIRE(!synced || (sync->exchange && sync->n_shards >= 1 && sync->sched_docs >= N), "bad shard sync");
    const long long SN = synced ? (long long)sync->sched_docs : N;     // the schedule every shard follows
    // Rounds end on tile boundaries where that keeps them overflow-free (a tile cut by a boundary is accumulated twice)
    auto align_hi = [&](long long lo, long long hi) {
        const long long a = hi / ix->tile_docs * ix->tile_docs;
        return a > lo ? a : hi;
    };
    long long lo = 0, hi = SN < cap ? SN : align_hi(0, cap);
    while (true) {
        A.r_lo = lo < N ? lo : N;
        A.r_hi = hi < N ? hi : N;
        if (A.r_hi > A.r_lo) {
            A.tile_lo = (int)(A.r_lo / ix->tile_docs);
            A.tile_hi = (int)ceil_div<long long>(A.r_hi, ix->tile_docs);
            A.group_lo = A.tile_lo / kGroupTiles;
            const long long blocks = (long long)(ceil_div(A.tile_hi, kGroupTiles) - A.group_lo) * n_queries;
            FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
            ProfScope prof(sizeof(AccT) == 8 ? "sparse_tile_f64" : "sparse_tile_f32", stream);
            if constexpr (std::is_same<AccT, float>::value) {
                if (splade_fixed_point_enabled()) sparse_tile_fx_kernel<0><<<(unsigned)blocks, threads, smem, stream>>>(A);
                else sparse_tile_kernel<AccT, 0><<<(unsigned)blocks, threads, smem, stream>>>(A);
            } else {
                sparse_tile_kernel<AccT, 0><<<(unsigned)blocks, threads, smem, stream>>>(A);
            }
        }
        FZ_LAUNCH_CHECK();
        const bool last = hi >= SN;
        const AccT* floor = last ? nullptr : shard_floor<AccT>(sync, A.st, n_queries, k, (AccT)0, stream, &rc);
        if (rc) return rc;
        rc = cand_select<AccT>(A.st, n_queries, k, (AccT)0, last, doc_base, out_scores, out_ids, nullptr, stream, floor);
        if (rc) return rc;
        if (last) break;
        lo = hi;
        hi = growth >= 2 ? hi * growth : hi + (cap - k);
        if (hi >= SN) hi = SN; else hi = align_hi(lo, hi);
    }
    ProfScope prof("sparse_zero_fill", stream);
    sparse_zero_fill_kernel<AccT><<<n_queries, threads, smem, stream>>>(A);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int sparse_bootstrap_f32(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, int k, const CandState<float>& st, bool fixed_point, cudaStream_t stream) {
    int rc = check_index<float>(ix);
    if (rc) return rc;
    rc = set_smem_attrs<float>();
    if (rc) return rc;
    SparseArgs<float> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = q_ptr;
    A.q_term = q_term;
    A.q_weight = q_weight;
    A.n_queries = n_queries;
    A.sign_mode = 1;
    A.st = st;
    A.k = k;
    const int threads = sparse_threads<float>(ix->tile_docs);
    const size_t smem = sparse_smem<float>(ix->tile_docs);
    const long long N = ix->n_docs;
    const int cap = st.cap;
    auto align_hi = [&](long long lo, long long hi) {
        const long long a = hi / ix->tile_docs * ix->tile_docs;
        return a > lo ? a : hi;
    };
    long long lo = 0, hi = N < cap ? N : align_hi(0, cap);
    while (true) {
        A.r_lo = lo;
        A.r_hi = hi;
        A.tile_lo = (int)(A.r_lo / ix->tile_docs);
        A.tile_hi = (int)ceil_div<long long>(A.r_hi, ix->tile_docs);
        A.group_lo = A.tile_lo / kGroupTiles;
        const long long blocks = (long long)(ceil_div(A.tile_hi, kGroupTiles) - A.group_lo) * n_queries;
        FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
        {
            ProfScope prof("splade_bootstrap", stream);
            if (fixed_point) sparse_tile_fx_kernel<0><<<(unsigned)blocks, threads, smem, stream>>>(A);
            else sparse_tile_kernel<float, 0><<<(unsigned)blocks, threads, smem, stream>>>(A);
            FZ_LAUNCH_CHECK();
        }
        rc = cand_select<float>(st, n_queries, k, 0.f, false, 0, nullptr, nullptr, nullptr, stream, nullptr);
        if (rc) return rc;
        if (hi >= N) break;
        lo = hi;
        hi = hi * 4;
        if (hi >= N) hi = N; else hi = align_hi(lo, hi);
    }
    return FZ_OK;
}

template <typename AccT>
static int sparse_scores(const fz_postings_t* ix, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, AccT* out, int accumulate, cudaStream_t stream) {
    int rc = check_index<AccT>(ix);
    if (rc) return rc;
    FZ_REQUIRE(q_ptr && q_term && out, "null pointer");
    if (n_queries == 0) return FZ_OK;
    rc = set_smem_attrs<AccT>();
    if (rc) return rc;
    SparseArgs<AccT> A;
    memset(&A, 0, sizeof(A));
    A.ix = *ix;
    A.q_ptr = q_ptr;
    A.q_term = q_term;
    A.q_weight = q_weight;
    A.n_queries = n_queries;
    A.out_full = out;
    A.accumulate = accumulate;
    A.tile_lo = 0;
    A.tile_hi = ix->n_tiles;
    A.group_lo = 0;
    const long long blocks = (long long)ceil_div(ix->n_tiles, kGroupTiles) * n_queries;
    FZ_REQUIRE(blocks < (1ll << 31), "grid too large");
    sparse_tile_kernel<AccT, 1><<<(unsigned)blocks, sparse_threads<AccT>(ix->tile_docs), sparse_smem<AccT>(ix->tile_docs), stream>>>(A);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // namespace fz

using namespace fz;

extern "C" {

int fz_lexical_impacts(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* post_tf, const int32_t* doc_len,
                       const double* idf, int32_t n_terms, int64_t nnz, double avgdl, double k1, double b, int variant,
                       double* out_impact, fz_stream_t stream) {
    FZ_REQUIRE(term_ptr && post_doc && post_tf && idf && out_impact, "null pointer");
    FZ_REQUIRE(variant == FZ_LEX_TFIDF || variant == FZ_LEX_BM25, "unknown lexical variant %d", variant);
    FZ_REQUIRE(variant == FZ_LEX_TFIDF || doc_len, "BM25 needs doc_len");
    if (nnz == 0) return FZ_OK;
    int blocks = (int)(ceil_div<long long>(nnz, 256) < 148 * 32 ? ceil_div<long long>(nnz, 256) : 148 * 32);
    lexical_impacts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(term_ptr, post_doc, post_tf, doc_len, idf, n_terms,
                                                                     nnz, avgdl, k1, k1 + 1, 1 - b, b, variant, out_impact);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

size_t fz_sparse_topk_workspace_bytes(int n_queries, int k, int cap, int is_f64) {
    (void)k;
    return is_f64 ? cand_state_bytes<double>(n_queries, cap) : cand_state_bytes<float>(n_queries, cap);
}

int fz_sparse_topk_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries, int k,
                       int64_t doc_base, int cap, int growth, int sign_mode, double* out_scores, int32_t* out_ids,
                       int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync, fz_stream_t stream) {
    return sparse_topk<double>(index, q_ptr, q_term, nullptr, n_queries, k, doc_base, cap, growth, sign_mode,
                               out_scores, out_ids, out_status, ws, ws_bytes, sync, (cudaStream_t)stream);
}

int fz_sparse_topk_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                       int n_queries, int k, int64_t doc_base, int cap, int growth, int sign_mode, float* out_scores,
                       int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync,
                       fz_stream_t stream) {
    return sparse_topk<float>(index, q_ptr, q_term, q_weight, n_queries, k, doc_base, cap, growth, sign_mode,
                              out_scores, out_ids, out_status, ws, ws_bytes, sync, (cudaStream_t)stream);
}

int fz_sparse_scores_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries,
                         double* out_scores, int accumulate, fz_stream_t stream) {
    return sparse_scores<double>(index, q_ptr, q_term, nullptr, n_queries, out_scores, accumulate, (cudaStream_t)stream);
}

int fz_sparse_scores_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, float* out_scores, int accumulate, fz_stream_t stream) {
    return sparse_scores<float>(index, q_ptr, q_term, q_weight, n_queries, out_scores, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
