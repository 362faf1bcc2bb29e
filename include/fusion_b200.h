/* fusion_b200 — C ABI of the B200 (sm_100a) retrieval-scoring + rank-fusion library.
 *
 * The reference (maastrichtlawtech/fusion) is pure Python and has no FFI layer: its boundary is the Python
 * call signatures listed per entry point below.  The Python package `fusion_b200` keeps those signatures and
 * calls these functions through ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes.  Every pointer is DEVICE memory unless the name ends in `_h`.
 *   - The caller owns all memory (inputs, outputs, workspace); `*_workspace_bytes` sizes the scratch.
 *   - Functions enqueue work on `stream` and return; they never synchronise, allocate or throw.
 *   - Return 0 on success, FZ_ERR_* (< 0) otherwise; `fz_last_error()` holds the message (thread local).
 *   - Document ids are int32, global = doc_base + row inside the shard.  Padding slots hold id -1.
 */
#ifndef FUSION_B200_H
#define FUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FZ_OK 0
#define FZ_ERR_ARG (-1)
#define FZ_ERR_CUDA (-2)
#define FZ_ERR_UNSUPPORTED (-3)

#define FZ_ABI_VERSION 2

typedef void* fz_stream_t; /* cudaStream_t */

const char* fz_last_error(void);
int fz_abi_version(void);

/* Optional per-kernel timing for the benchmark: CUDA events on the launching stream around every kernel launch.
 * fz_profile_summary synchronises, writes one "name launches total_ms" line per kernel into out, and resets. */
int fz_profile_enable(int on);
int fz_profile_summary(char* out, size_t cap);
/* Debugging aid: device buffer of [n_ctas, 8] uint64 cycle counters the persistent tensor-core kernels fill
 * (barrier wait / issue cycles per warp role); NULL switches it off. */
int fz_debug_set_stats(void* device_buffer);

/* per-query status bits written by the top-k entry points */
#define FZ_STATUS_OVERFLOW 1  /* candidate buffer overflowed: result for this query is NOT valid, re-run with growth=1 */
#define FZ_STATUS_NEED_ZERO 2 /* fewer than k positive-score docs: zero-score docs were appended in doc-id order */
#define FZ_STATUS_NEED_NEG 4  /* positives + zero-score docs < k: negative-score docs are missing from the tail */
#define FZ_STATUS_TOO_LONG 16 /* the query has more than FZ_MAX_QUERY_TERMS terms: result NOT valid (see fz_sparse_scores_*) */
#define FZ_STATUS_FALLBACK 8  /* fz_splade_topk only: this query is outside the fast path's contract (negative weights, fewer than
                                k positive-score docs, a tail sum outside the code range): re-run it with fz_sparse_topk_f32 */

/* ------------------------------------------------------------------------------------------------------------
 * K5  k-way merge of per-shard / per-chunk top-k lists.
 * Replaces the Python heapq merges in src/retrievers/splade/base.py:235-243 and
 * src/utils/sentence_transformers.py:358-364, and is the merge step after the multi-GPU all-gather.
 *   scores/ids: [n_src, n_queries, k_in]  (entries with id < 0 are ignored)
 *   out:        [n_queries, k_out], best first (score desc, ties by lower id), padded with (-inf, -1)
 * ---------------------------------------------------------------------------------------------------------- */
size_t fz_merge_topk_workspace_bytes(int n_src, int n_queries, int k_in);
int fz_merge_topk_f32(const float* scores, const int32_t* ids, int n_src, int n_queries, int k_in, int k_out,
                      float* out_scores, int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream);
int fz_merge_topk_f64(const double* scores, const int32_t* ids, int n_src, int n_queries, int k_in, int k_out,
                      double* out_scores, int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream);

/* Full ranking of dense score rows (the reference's `top_k = len(documents)` mode, src/retrievers/hybrid.py:74,103):
 *   scores [n_queries, n_docs] -> out [n_queries, k] sorted by score desc, ties by lower doc index (stable sort,
 *   src/retrievers/bm25.py:105). */
size_t fz_rank_rows_workspace_bytes(int n_queries, int64_t n_docs);
int fz_rank_rows_f32(const float* scores, int n_queries, int64_t n_docs, int k, int64_t doc_base, float* out_scores,
                     int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream);
int fz_rank_rows_f64(const double* scores, int n_queries, int64_t n_docs, int k, int64_t doc_base, double* out_scores,
                     int32_t* out_ids, void* ws, size_t ws_bytes, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4  rank fusion.  Replaces Aggregator.fuse / convert2dict / transform_scores / weight_scores /
 * aggregate_scores, src/retrievers/hybrid.py:170-307.
 *   ids_h[s], scores_h[s]: host arrays of n_sys DEVICE pointers, each [n_queries, list_stride[s]]
 *   lens_h[s]: device int32 [n_queries] (NULL => every list has list_stride[s] entries)
 *   score_is_f64_h[s]: 1 if scores_h[s] points at doubles, 0 for floats
 *   method: FZ_FUSE_*; normalization: FZ_NORM_* (nsf only); weights_h: host doubles [n_sys] (nsf only)
 *   distr_h[s]: device float32 percentile distribution of length distr_len_h[s], ASCENDING (percentile / NCE only)
 *   out_ids [n_queries, out_stride], out_scores (double) [n_queries, out_stride], out_len [n_queries]:
 *       the union of the lists, fused score descending, ties by first insertion (system order, then rank)
 * ---------------------------------------------------------------------------------------------------------- */
#define FZ_FUSE_BCF 0
#define FZ_FUSE_RRF 1
#define FZ_FUSE_NSF 2
#define FZ_FUSE_KEEP_ORDER 0x100 /* OR into `method`, one system only: the transformed, deduplicated list in first-insertion
                                  order instead of sorted (the per-system normalisation step of fz_fuse_sweep) */
#define FZ_FUSE_PROMOTE_F64 0x200 /* OR into `method` (nsf with a torch normalisation): weight and sum the fp32 normalised scores
                                  in fp64.  weight_scores / aggregate_scores (hybrid.py:283-307) compute np.float32 * python
                                  float: float32 under NumPy >= 2 (NEP 50, the default here), float64 under the NumPy 1.x the
                                  reference pins */
#define FZ_NORM_NONE 0
#define FZ_NORM_MINMAX 1
#define FZ_NORM_ZSCORE 2
#define FZ_NORM_ARCTAN 3
#define FZ_NORM_PERCENTILE 4
#define FZ_NORM_NCE 5
#define FZ_NORM_IDENTITY_F32 6 /* scores taken as fp32, weighted and summed in fp32 (np.float32 inputs to aggregate_scores) */
#define FZ_FUSE_MAX_SYSTEMS 8

size_t fz_fuse_workspace_bytes(int n_sys, int n_queries, const int32_t* list_stride_h);
int fz_fuse(const int32_t* const* ids_h, const void* const* scores_h, const int32_t* const* lens_h,
            const int32_t* score_is_f64_h, const int32_t* list_stride_h, int n_sys, int n_queries, int method,
            int normalization, const double* weights_h, const float* const* distr_h, const int32_t* distr_len_h,
            int32_t* out_ids, double* out_scores, int32_t* out_len, int out_stride, void* ws, size_t ws_bytes,
            fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Evaluation next to the path.
 * fz_rank_metrics: recall@k / map@k / mrr@k / ndcg@k / R-precision of ranked id lists against gold id lists, SUMMED over
 * the queries (divide by n_queries for Metrics.compute_all_metrics, src/utils/metrics.py:40-162; the metric set of
 * run_evaluation, src/retrievers/hybrid.py:24-27, is recall {5,10,20,50,100,200,500,1000}, map/mrr/ndcg {10,100}).
 *   ids [n_queries, stride] (-1 padded), lens [n_queries] or NULL, gold_ptr [n_queries + 1], gold_ids: device memory
 *   *_k_h: host arrays of cut-offs (at most 8 each); a query with more than 256 distinct gold ids turns the outputs into
 *   NaN (never truncated silently)
 *   out_sum: device doubles [n_recall + n_map + n_mrr + n_ndcg + 1] in that order, R-precision last
 *   ws_per_query: device doubles [n_queries, n_metrics] or NULL; with it the per-query values are summed in query order
 *   (bit-reproducible), without it by atomics in arbitrary order
 * fz_fuse_sweep: the linear-fusion weight sweep of src/retrievers/hybrid.py:404-426 - for every weight vector, fuse the
 * systems' lists by weighted sum (union, missing = 0, stable descending order) and evaluate; out_sum [n_weights, M].
 *   ids_h[s] / vals_h[s]: host arrays of device pointers to the ALREADY NORMALISED lists [n_queries, list_stride[s]] without
 *   repeated ids (fz_fuse on the single system with weight 1 produces exactly that); values_are_f32 = 1 when the values
 *   are fp32 numbers (any torch normalisation; summed in fp32 like the reference), 0 for normalization 'none' (fp64);
 *   weights: device doubles [n_weights, n_sys].  The union of one query's lists must fit shared memory (top-k lists).
 * ---------------------------------------------------------------------------------------------------------- */
int fz_rank_metrics(const int32_t* ids, const int32_t* lens, int n_queries, int stride, const int32_t* gold_ptr,
                    const int32_t* gold_ids, const int32_t* recall_k_h, int n_recall, const int32_t* map_k_h, int n_map,
                    const int32_t* mrr_k_h, int n_mrr, const int32_t* ndcg_k_h, int n_ndcg, double* out_sum,
                    double* ws_per_query, fz_stream_t stream);
int fz_fuse_sweep(const int32_t* const* ids_h, const double* const* vals_h, const int32_t* const* lens_h,
                  const int32_t* list_stride_h, int n_sys, int n_queries, int values_are_f32, const double* weights,
                  int n_weights, const int32_t* gold_ptr, const int32_t* gold_ids, const int32_t* recall_k_h, int n_recall,
                  const int32_t* map_k_h, int n_map, const int32_t* mrr_k_h, int n_mrr, const int32_t* ndcg_k_h, int n_ndcg,
                  double* out_sum, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Index-build helpers next to the path.
 * fz_token_starts / fz_hash_tokens: whitespace tokenisation of a UTF-8 buffer with the rules of Python's str.split()
 * (`doc.split()`, src/retrievers/bm25.py:54-60,72,81,101; the Unicode whitespace set of str.isspace()).
 *   out_flags [n_bytes] uint8: 1 where a token starts.  starts [n_tokens] int64: those byte offsets.
 *   out_h1 / out_h2 [n_tokens] int64: two independent 64-bit hashes of the token bytes (equal tokens <=> equal pairs,
 *   up to a 2^-128 collision); out_len [n_tokens] int32 token length in bytes (may be NULL).
 * fz_quantiles_f64: np.percentile(..., method='linear') of an ASCENDING array at linspace(0, 1, n_quantiles) - what
 * pandas Series.quantile computes for the percentile distributions (src/retrievers/hybrid.py:391-398).
 * ---------------------------------------------------------------------------------------------------------- */
int fz_token_starts(const void* utf8, int64_t n_bytes, void* out_flags, fz_stream_t stream);
int fz_hash_tokens(const void* utf8, int64_t n_bytes, const int64_t* starts, int64_t n_tokens, int64_t* out_h1,
                   int64_t* out_h2, int32_t* out_len, fz_stream_t stream);
int fz_quantiles_f64(const double* sorted, int64_t n, int n_quantiles, double* out, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Cross-shard threshold exchange of a corpus sharded over G GPUs (north-star "multi-GPU partitioning": each GPU keeps a
 * local top-k; the reference has no counterpart - its corpus lives in one process).  Without it every shard maintains its
 * own top-k and emits / selects ~k candidates per query and round whatever G is - a per-query cost that does not shrink
 * with the shard.  With it, between two rounds of a threshold-filter top-k every shard writes the ceil(k / n_shards)-th
 * best score it holds into `exchange`, calls `hook` - the caller all-reduces (MIN) the buffer in place on the call's
 * stream (NCCL through torch.distributed in fusion_b200/sharding.py) - and uses the result as a floor: at least k
 * documents over all shards reach it (a shard that holds fewer than floor_rank candidates publishes -inf: no floor
 * that round), so nothing below it can enter the global top-k (ties with it are kept: the global tie-break is by
 * document id).  Every shard then carries ~k / G candidates per query.
 *   sched_docs: the largest shard's n_docs; the round schedule is derived from it so that all shards call the hook the
 *   same number of times (a shard that runs out of documents runs empty rounds).
 * Results are unchanged: the union of the shards' lists still contains the global top-k. */
typedef int (*fz_shard_hook_t)(void* user);
typedef struct fz_shard_sync {
    fz_shard_hook_t hook;   /* NULL = no exchange */
    void* user;
    void* exchange;         /* device [n_queries] of the score type (double for _f64, float otherwise) */
    int32_t n_shards;
    int32_t floor_rank;     /* rank each shard publishes: ceil(K / n_shards) for the K of the GLOBAL top-K (a small shard
                               may run with k < K); 0 = ceil(k / n_shards) of this call's k */
    int64_t sched_docs;
} fz_shard_sync_t;

/* ------------------------------------------------------------------------------------------------------------
 * SPLADE activation head: between the encoder's MLM logits and the CSR vectors K2 consumes.
 * fz_splade_pool: activations[b, v] = sum_l or amax_l log1p(relu(logits[b, l, v] * mask[b, l]))
 *   (SPLADE.forward, src/retrievers/splade/splade.py:88-94).  logits [n_rows, seq_len, vocab] fp32 or bf16, row-major;
 *   mask [n_rows, seq_len] int32 (0 = padding: never loaded); out_act [n_rows, vocab] fp32.
 * fz_prune_topk: keep the keep_topk largest activations of every row and zero the rest (SPLADE._prune_activations,
 *   splade.py:295-306); ties at the cutoff keep the lower term id; out_act may alias act.
 * fz_csr_count / fz_csr_fill: dense [n_rows, vocab] -> CSR with zeros dropped and term ids ascending.  The caller
 *   turns out_nnz into row_ptr (exclusive prefix sum, int64 [n_rows + 1]) between the two calls.
 * ---------------------------------------------------------------------------------------------------------- */
#define FZ_POOL_MAX 0
#define FZ_POOL_SUM 1
int fz_splade_pool(const void* logits, int logits_are_bf16, const int32_t* mask, int n_rows, int seq_len, int vocab,
                   int pooling, float* out_act, fz_stream_t stream);
int fz_prune_topk(const float* act, int n_rows, int vocab, int keep_topk, float* out_act, fz_stream_t stream);
int fz_csr_count(const float* act, int n_rows, int vocab, int32_t* out_nnz, fz_stream_t stream);
int fz_csr_fill(const float* act, int n_rows, int vocab, const int64_t* row_ptr, int32_t* out_term, float* out_weight,
                fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2  sparse scoring over a term-major CSR inverted index, document range tiled for shared-memory accumulators.
 * Replaces TFIDF/BM25/AtireBM25.score + .search (src/retrievers/bm25.py:100-115,149-156) and, for SPLADE,
 * the dense [Q,V]x[V,N] cosine of src/retrievers/hybrid.py:101-103 / splade/base.py:186-197.
 * ---------------------------------------------------------------------------------------------------------- */
/* A term's postings are stored in ONE of three forms, chosen at index time from its document frequency:
 *   short  (df < tiled_min)        doc-ascending (doc, value) pairs plus a coarse table (postings before every 16th
 *                                  tile): a CTA walks 16 tiles and scans the few postings between two marks
 *   tiled  (tiled_min <= df)       per (term, doc tile) segments of (tile-relative uint16 doc offset, value), padded to
 *                                  a multiple of 4 postings with (tile_docs, 0) so that a thread moves 4 postings with
 *                                  one 8-byte and one 16/32-byte load; inside a segment the postings are ordered so that
 *                                  the 32 lanes of a warp hit 32 different shared-memory banks (the order inside a
 *                                  segment is free: a term never hits a doc twice)
 *   dense  (df >= dense_frac * N)  one value per document (0 where the doc lacks the term): no ids, no scatter -
 *                                  accumulated in registers with 16-byte coalesced loads                               */
typedef struct fz_postings {
    const int64_t* term_ptr;        /* [n_terms + 1] short lists: ranges into post_doc / post_val (empty otherwise) */
    const int32_t* post_doc;        /* short postings: local doc row, ascending inside a term */
    const void* post_val;           /* short postings: double impacts (lexical) or float weights (SPLADE) */
    const uint16_t* short_coarse;   /* [n_terms, n_coarse + 1] postings of the term below tile 16 * c (rows of non-short terms unused) */
    const int32_t* term_slot;       /* [n_terms] -1 = short, r >= 0 = tiled row r, v <= -2 = dense row (-2 - v) */
    const int64_t* tiled_base;      /* [n_tiled] first posting of the term in tiled_off / tiled_val (multiple of 4) */
    const uint32_t* tiled_tile_off; /* [n_tiled, n_tiles + 1] segment starts relative to tiled_base (multiples of 4) */
    const uint16_t* tiled_off;      /* tile-relative doc offsets, == tile_docs for padding */
    const void* tiled_val;
    const void* dense_val;          /* [n_dense, dense_stride] */
    int64_t dense_stride;           /* n_tiles * tile_docs */
    int32_t n_terms;
    int32_t n_tiled;
    int32_t n_dense;
    int32_t tile_docs;              /* docs per shared-memory accumulator tile: multiple of 4, <= 8192 (f32) / 4096 (f64) */
    int32_t n_tiles;
    int32_t n_coarse;               /* ceil(n_tiles / FZ_COARSE_TILES) */
    int64_t n_docs;
} fz_postings_t;
#ifndef FZ_COARSE_TILES
#define FZ_COARSE_TILES 16 /* also the number of consecutive tiles one CTA walks */
#endif

#define FZ_MAX_QUERY_TERMS 128
#define FZ_LEX_TFIDF 0
#define FZ_LEX_BM25 1 /* also ATIRE: only the idf table differs */

/* per-posting fp64 impact, evaluated with the reference's operation order and no FMA contraction:
 *   TF-IDF: tf * idf                                           (bm25.py:114)
 *   BM25:   idf * (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * dl / avgdl))   (bm25.py:155) */
int fz_lexical_impacts(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* post_tf, const int32_t* doc_len,
                       const double* idf, int32_t n_terms, int64_t nnz, double avgdl, double k1, double b, int variant,
                       double* out_impact, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Index build ("fz_build_*"): everything between "token ids / sparse vectors per document" and the
 * device arrays the scoring kernels read, so that a non-Python host can build an index.  Replaces the dict-of-dicts
 * index construction of the reference: TFIDF._build_index + BM25._build_index (src/retrievers/bm25.py:53-83,141-143:
 * vocabulary df, per-doc term frequencies, document lengths) and, for SPLADE, the [N, V] activation matrix of
 * splade/base.py:186-197 (held here as CSR and turned into an inverted index).  All pointers are device pointers unless
 * the name ends in _h.  Every builder is "plan -> the caller allocates -> fill": a *_plan call returns the array sizes
 * to the host and synchronises the stream once; nothing in the library allocates device memory.  The plan and the fill
 * call of one build must be given the SAME workspace, untouched in between (the plan leaves its sort / scan results there).
 *
 * fz_build_lexical_*   token ids per document (doc_ptr [n_docs + 1], doc_tok [n_tokens] in [0, vocab)) -> term-major
 *                      postings (term_ptr [vocab + 1], post_doc, post_tf: doc-ascending inside a term, tf = occurrences of
 *                      the term in the doc) and doc_len [n_docs].  plan returns nnz = distinct (term, doc) pairs.
 * fz_build_term_major  (row, term) per entry of a doc-major CSR -> term_ptr [n_terms + 1] and the permutation out_order
 *                      [nnz] (entry indices in term-major / row-ascending order) - the transpose used for SPLADE vectors.
 * fz_build_postings_*  term-major CSR (term_ptr, post_doc ascending inside a term, post_val float or double) -> the three
 *                      storage forms of fz_postings_t.  tiled_min / dense_min: df from which a term is stored as tile
 *                      segments / as a dense row (dense_min = 2^62: never).  plan writes short_ptr (= fz_postings_t.term_ptr)
 *                      and term_slot and returns the sizes of the other arrays:
 *                        post_doc / post_val [n_short]; short_coarse [n_terms, n_coarse + 1]; tiled_base [n_tiled];
 *                        tiled_tile_off [n_tiled, n_tiles + 1]; tiled_off / tiled_val [n_tiled_entries];
 *                        dense_val [n_dense, dense_stride].
 * fz_build_csr_normalize   w / max(|row|_2, 1e-12) per CSR row (cos_sim: hybrid.py:101-103 normalises both sides)
 * fz_build_term_stats      df [n_terms] and the largest weight of every term (0 for unused terms); out_flags (device
 *                          int32): bit 0 = a negative weight exists (no head/tail split), bit 1 = a term id out of range
 * fz_build_splade_head     the head matrix of fz_splade_head_t: head[d, term_head[t]] = bf16(w) for the head terms
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct fz_build_plan {
    int64_t n_short;         /* postings kept as short lists */
    int64_t n_tiled_entries; /* tile-segment entries incl. padding */
    int64_t dense_stride;    /* n_tiles * tile_docs */
    int32_t n_tiled;
    int32_t n_dense;
    int32_t n_tiles;
    int32_t n_coarse;
} fz_build_plan_t;

size_t fz_build_lexical_workspace_bytes(int64_t n_tokens, int32_t vocab);
int fz_build_lexical_plan(const int64_t* doc_ptr, const int32_t* doc_tok, int64_t n_docs, int64_t n_tokens, int32_t vocab,
                          int64_t* out_nnz_h, void* ws, size_t ws_bytes, fz_stream_t stream);
int fz_build_lexical_fill(const int64_t* doc_ptr, int64_t n_docs, int64_t n_tokens, int32_t vocab, int64_t nnz,
                          int64_t* out_term_ptr, int32_t* out_post_doc, int32_t* out_post_tf, int32_t* out_doc_len, void* ws,
                          size_t ws_bytes, fz_stream_t stream);
size_t fz_build_term_major_workspace_bytes(int64_t nnz, int32_t n_terms);
int fz_build_term_major(const int32_t* row, const int32_t* term, int64_t nnz, int64_t n_rows, int32_t n_terms,
                        int64_t* out_term_ptr, int64_t* out_order, void* ws, size_t ws_bytes, fz_stream_t stream);
size_t fz_build_postings_workspace_bytes(int32_t n_terms);
int fz_build_postings_plan(const int64_t* term_ptr, const int32_t* post_doc, int32_t n_terms, int64_t n_docs, int32_t tile_docs,
                           int64_t tiled_min, int64_t dense_min, int64_t* out_short_ptr, int32_t* out_term_slot,
                           fz_build_plan_t* out_plan_h, void* ws, size_t ws_bytes, fz_stream_t stream);
int fz_build_postings_fill(const int64_t* term_ptr, const int32_t* post_doc, const void* post_val, int value_bytes,
                           int32_t n_terms, int64_t n_docs, int32_t tile_docs, const int64_t* short_ptr,
                           const int32_t* term_slot, const fz_build_plan_t* plan_h, int64_t nnz, int32_t* out_short_doc,
                           void* out_short_val, uint16_t* out_short_coarse, int64_t* out_tiled_base,
                           uint32_t* out_tiled_tile_off, uint16_t* out_tiled_off, void* out_tiled_val, void* out_dense_val,
                           void* ws, size_t ws_bytes, fz_stream_t stream);
int fz_build_csr_normalize(const int64_t* doc_ptr, const float* weight, int64_t n_docs, float* out_weight, fz_stream_t stream);
int fz_build_term_stats(const int32_t* term, const float* weight, int64_t nnz, int32_t n_terms, int64_t* out_df,
                        float* out_term_max, int32_t* out_flags, fz_stream_t stream);
int fz_build_splade_head(const int64_t* doc_ptr, const int32_t* term, const float* weight, int64_t n_docs, int64_t nnz,
                         const int32_t* term_head, int32_t head_dim, void* out_head_bf16, fz_stream_t stream);

/* top-k: out [n_queries, k] (score desc, ties by lower doc id); zero-score docs fill up in doc-id order.
 *   q_ptr [n_queries+1], q_term [nq] (term ids in query-token order, duplicates kept, -1 = out of vocabulary;
 *   at most FZ_MAX_QUERY_TERMS per query: a longer query gets FZ_STATUS_TOO_LONG in out_status and an unspecified row -
 *   never a silently truncated score; score it with fz_sparse_scores_* in chunks of FZ_MAX_QUERY_TERMS terms
 *   (accumulate = 1 from the second chunk on) + fz_rank_rows_*),
 *   q_weight [nq] float (f32 variant only; NULL => 1)
 *   growth: >= 2 geometric round growth (fast path), 1 = conservative rounds that can never overflow
 *   out_status [n_queries]: FZ_STATUS_* bits                                                              */
size_t fz_sparse_topk_workspace_bytes(int n_queries, int k, int cap, int is_f64);
int fz_sparse_topk_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries, int k,
                       int64_t doc_base, int cap, int growth, int sign_mode, double* out_scores, int32_t* out_ids,
                       int32_t* out_status, void* ws, size_t ws_bytes, const fz_shard_sync_t* sync /* may be NULL */,
                       fz_stream_t stream);
int fz_sparse_topk_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                       int n_queries, int k, int64_t doc_base, int cap, int growth, int sign_mode, float* out_scores,
                       int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes,
                       const fz_shard_sync_t* sync /* may be NULL */, fz_stream_t stream);
/* every document's score: out [n_queries, n_docs] (full-ranking mode and tests).  accumulate != 0: the rows already hold the
 * sums of the query's EARLIER terms and this call continues them in the same left-to-right order - how a query of more than
 * FZ_MAX_QUERY_TERMS terms is scored in several calls (bit-identical to one pass over all its terms) */
int fz_sparse_scores_f64(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, int n_queries,
                         double* out_scores, int accumulate, fz_stream_t stream);
int fz_sparse_scores_f32(const fz_postings_t* index, const int32_t* q_ptr, const int32_t* q_term, const float* q_weight,
                         int n_queries, float* out_scores, int accumulate, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2c  SPLADE top-k as head GEMM + tail bound + exact rescoring (non-negative weights: SPLADE activations are
 * log1p(relu(.)) >= 0, src/retrievers/splade/splade.py:94).  Same contract as fz_sparse_topk_f32 - the dense [Q,V]x[V,N]
 * cosine of src/retrievers/hybrid.py:101-103 / splade/base.py:186-251, scores in fp32 to 1e-5 - for the queries it can serve.
 *
 * A Zipfian vocabulary puts > 95 % of all (query term, posting) pairs into the ~200 most frequent terms.  Those HEAD terms
 * are stored as a dense doc-major bf16 matrix and scored on the tensor cores (the K1 filter GEMM); the remaining TAIL terms
 * keep the inverted index (fz_postings_t built from the tail terms only, no dense rows, tile_docs % 256 == 0), but their
 * sum is only needed as an upper bound: a 4-bit code per (query, doc).  A doc whose head score + tail bound can still beat
 * the query's running k-th EXACT score is appended to the candidate buffer and rescored exactly in fp32 from the doc-major
 * CSR copy; between rounds cand_select tightens the threshold on exact scores.  The returned scores are the exact ones.
 *   head_bf16 [n_docs, head_dim]: weight of head term j in doc d (bf16, row-major, head_dim % 64 == 0, 64..256)
 *   term_head [n_terms]: head column of a term, -1 for tail terms;  term_max [n_terms]: max weight of the term in the shard
 *   doc_ptr [n_docs + 1], doc_post [nnz] (int32 term, float weight) pairs: the doc-major copy, terms in any order
 * boot_index (may be NULL): a general inverted index (all terms, any storage forms) over the FIRST boot_index->n_docs docs of
 * the shard.  Those docs are scored with the K2 kernel first - its exact scores cost no rescoring - so that the head/tail
 * rounds start with a threshold that few docs beat: a round emits about k docs per doubling of the range whatever its size,
 * and every emitted doc costs a ~0.75 KB rescoring gather.
 * Queries that leave the fast path's contract get FZ_STATUS_FALLBACK (or FZ_STATUS_OVERFLOW) in out_status and an
 * unspecified row: the caller re-runs those with fz_sparse_topk_f32 on the full index.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct fz_splade_head {
    const void* head_bf16;
    const int32_t* term_head;
    const float* term_max;
    const int64_t* doc_ptr;
    const void* doc_post;
    int32_t head_dim;
    int32_t n_terms;
    int64_t n_docs;
    int32_t flags;                /* FZ_SPLADE_UNIT_ROWS: every doc vector has norm <= 1 (cos_sim index) */
    int32_t reserved;
} fz_splade_head_t;
#define FZ_SPLADE_UNIT_ROWS 1

/* max_round_docs bounds the code buffer (n_queries * max_round_docs / 2 bytes): rounds never span more documents */
size_t fz_splade_topk_workspace_bytes(int n_queries, int k, int cap, int head_dim, int64_t max_round_docs);
int fz_splade_topk(const fz_postings_t* tail_index, const fz_splade_head_t* head, const fz_postings_t* boot_index,
                   const int32_t* q_ptr, const int32_t* q_term,
                   const float* q_weight, int n_queries, int k, int64_t doc_base, int cap, int growth, float* out_scores,
                   int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes,
                   const fz_shard_sync_t* sync /* may be NULL */, fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1  dense exhaustive inner-product scoring with top-k fused into the tcgen05 GEMM epilogue.
 * Replaces sentence_transformers.util.semantic_search (src/retrievers/hybrid.py:103), BaseModel.search /
 * compute_batchwise_similarity (src/retrievers/splade/base.py:186-251) and the scoring loop of
 * InformationRetrievalEvaluatorCustom.compute_metrices (src/utils/sentence_transformers.py:334-364).
 * Inputs are already L2-normalised for cos_sim (fz_normalize_rows).
 *   q_bf16 [n_queries, dim], d_bf16 [n_docs, dim]: tensor-core operands (dim % 64 == 0)
 *   q_f32 / d_f32: fp32 originals for exact rescoring of the survivors, or NULL for the bf16 throughput mode
 *   margin: candidates within `margin` below the running k-th bf16 score are kept for rescoring
 * ---------------------------------------------------------------------------------------------------------- */
size_t fz_dense_topk_workspace_bytes(int n_queries, int k, int cap);
int fz_dense_topk(const void* q_bf16, const void* d_bf16, const float* q_f32, const float* d_f32, int n_queries,
                  int64_t n_docs, int dim, int k, float margin, int64_t doc_base, int cap, int growth,
                  float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws, size_t ws_bytes,
                  fz_stream_t stream);
/* The exact mode in two calls, for corpus-sharded runs over G shards: _filter leaves the surviving candidates in `ws` and
 * writes, per query, (the floor_rank-th best bf16 score of this shard) - margin to out_tau (-inf if the shard has fewer).
 * With floor_rank = ceil(k / G) the MINIMUM of out_tau over the shards (one all-reduce) is a lower bound of
 * (global k-th bf16 score) - margin: the G shards' best ceil(k / G) are >= k documents.  Passed to _finish as tau_floor,
 * it spares the fp32 rescoring (a 3 KB row each) of every candidate that cannot reach the global top-k.
 * fz_dense_topk == _filter + _finish(tau_floor = NULL). */
int fz_dense_topk_filter(const void* q_bf16, const void* d_bf16, int n_queries, int64_t n_docs, int dim, int k, float margin,
                         int64_t doc_base, int cap, int growth, int floor_rank, float* out_tau, int32_t* out_status,
                         void* ws, size_t ws_bytes, const fz_shard_sync_t* sync /* may be NULL */, fz_stream_t stream);
int fz_dense_topk_finish(const float* q_f32, const float* d_f32, const float* tau_floor, int n_queries, int dim, int k,
                         int64_t doc_base, int cap, float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws,
                         size_t ws_bytes, fz_stream_t stream);
/* exact fp32 scores of every (query, doc) pair on CUDA cores: out [n_queries, n_docs] (full-ranking mode) */
int fz_dense_scores_f32(const float* q_f32, const float* d_f32, int n_queries, int64_t n_docs, int dim,
                        float* out_scores, fz_stream_t stream);
/* out[i] = <q[i,:], d[i,:]> (compute_pairwise_similarity, src/retrievers/splade/base.py:173-184) */
int fz_pairwise_dot_f32(const float* q_f32, const float* d_f32, int64_t n_rows, int dim, float* out, fz_stream_t stream);
/* rows / max(||row||, 1e-12) (torch.nn.functional.normalize, base.py:195-196); either output may be NULL;
 * normalize == 0 only converts. */
int fz_normalize_rows(const float* x, int64_t n_rows, int dim, int normalize, float* out_f32, void* out_bf16,
                      fz_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  ColBERT MaxSim over candidate token tiles: S(q, d) = sum_i max_j <Q_i, D_j>  (tcgen05, bf16 in, fp32 acc).
 * Replaces colbert-ai `colbert_score` reached through CustomSearcher.search_all (src/utils/colbert_ir.py:245-255)
 * and Ranker.multi_vector_search (src/retrievers/hybrid.py:109-137).
 *
 * The token store is kept PACKED: per passage the two 64-dim halves as 128-byte rows, rows padded to a multiple of 8,
 * 16-byte chunks pre-swizzled for the UMMA shared-memory layout, so a passage is fetched with two plain bulk copies.
 *   fz_maxsim_pack: tok_ptr [n_docs + 1] int64 token offsets, tok_emb [n_tokens, 128] bf16 (16-byte aligned),
 *                   pk_ptr [n_docs + 1] int64 packed-row offsets (pk_ptr[d+1] - pk_ptr[d] = tokens of d rounded up to 8),
 *                   packed [pk_ptr[n_docs] * 256 bytes], 1024-byte aligned.
 *   fz_maxsim_bf16: q_tok [n_queries * lq, 128] bf16, cand_ids [n_queries, n_cand] global ids (ids outside
 *                   [doc_base, doc_base + n_docs) are skipped, score 0), out_scores [n_queries, n_cand] fp32,
 *                   ws: fz_maxsim_workspace_bytes(...) of scratch, 16-byte aligned (the candidates this shard owns, per query).
 * ---------------------------------------------------------------------------------------------------------- */
size_t fz_maxsim_workspace_bytes(int n_queries, int n_cand);
int fz_maxsim_pack(const int64_t* tok_ptr, const void* tok_emb, const int64_t* pk_ptr, int64_t n_docs, void* packed,
                   fz_stream_t stream);
int fz_maxsim_bf16(const void* q_tok, int lq, const int32_t* cand_ids, const int64_t* tok_ptr, const int64_t* pk_ptr,
                   const void* packed, int64_t n_docs, int64_t doc_base, int n_queries, int n_cand, float* out_scores,
                   void* ws, size_t ws_bytes, fz_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FUSION_B200_H */
