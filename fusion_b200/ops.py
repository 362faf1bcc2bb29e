"""Tensor API: torch CUDA tensors in, torch CUDA tensors out, every op one or more calls into the C ABI.

torch is plumbing here (device memory, the current stream); the arithmetic is in libfusion_b200.so.  This is
the level ``bench.py`` measures; the drop-in adapters in ``fusion_b200.retrievers`` wrap it into the
reference's list-of-dict shapes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (FZ_STATUS_FALLBACK, FZ_STATUS_NEED_NEG, FZ_STATUS_OVERFLOW, FZ_STATUS_TOO_LONG, FusionB200Error,
                   check)

DEFAULT_CAP = 8192
COARSE_TILES = 16          # FZ_COARSE_TILES
DEFAULT_GROWTH = 4


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise FusionB200Error(f"{name} must be a CUDA tensor (fusion_b200 has no CPU path)")
    if t.dtype != dtype:
        raise FusionB200Error(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------- K5
def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k_out: int):
    """k-way merge of ``n_src`` top-k lists: scores/ids [n_src, Q, k_in] -> ([Q, k_out], [Q, k_out]),
    best first (score desc, ties by lower id), padded with (-inf, -1); entries with id < 0 are ignored."""
    lib = _lib.load()
    f64 = scores.dtype == torch.float64
    scores = _req(scores, torch.float64 if f64 else torch.float32, "scores")
    ids = _req(ids, torch.int32, "ids")
    g, q, k_in = scores.shape
    out_s = torch.empty((q, k_out), dtype=scores.dtype, device=scores.device)
    out_i = torch.empty((q, k_out), dtype=torch.int32, device=scores.device)
    ws = _ws(lib.fz_merge_topk_workspace_bytes(g, q, k_in), scores.device)
    fn = lib.fz_merge_topk_f64 if f64 else lib.fz_merge_topk_f32
    check(fn(_ptr(scores), _ptr(ids), g, q, k_in, k_out, _ptr(out_s), _ptr(out_i), _ptr(ws), ws.numel(),
             _stream(scores)), "fz_merge_topk")
    return out_s, out_i


def rank_rows(scores: torch.Tensor, k: int, doc_base: int = 0, max_ws_bytes: int = 1 << 30):
    """Rank every column of each row: scores [Q, N] -> top-k (score desc, ties by lower index)."""
    lib = _lib.load()
    f64 = scores.dtype == torch.float64
    scores = _req(scores, torch.float64 if f64 else torch.float32, "scores")
    q, n = scores.shape
    out_s = torch.empty((q, k), dtype=scores.dtype, device=scores.device)
    out_i = torch.empty((q, k), dtype=torch.int32, device=scores.device)
    fn = lib.fz_rank_rows_f64 if f64 else lib.fz_rank_rows_f32
    per_q = max(1, lib.fz_rank_rows_workspace_bytes(1, n))
    step = max(1, min(q, max_ws_bytes // per_q))
    for lo in range(0, q, step):
        hi = min(q, lo + step)
        ws = _ws(lib.fz_rank_rows_workspace_bytes(hi - lo, n), scores.device)
        check(fn(_ptr(scores[lo:hi]), hi - lo, n, k, doc_base, _ptr(out_s[lo:hi]), _ptr(out_i[lo:hi]), _ptr(ws),
                 ws.numel(), _stream(scores)), "fz_rank_rows")
    return out_s, out_i


# ----------------------------------------------------------------------------------------------- K4
def fuse(lists, method: str, normalization: str | None = None, weights=None, distributions=None,
         out_stride: int | None = None, max_ws_bytes: int = 2 << 30, keep_order: bool = False, promote_f64: bool = False):
    """Fuse ``S`` ranked-list systems.

    lists: sequence of (ids int32 [Q, n_s], scores f32|f64 [Q, n_s], lens int32 [Q] | None), rank order.
    -> (ids int32 [Q, U], scores f64 [Q, U], lens int32 [Q]); U = out_stride or sum(n_s).  Rows are the union of
    the lists, fused score descending, ties by first insertion; padded with (-1, -inf).  ``keep_order`` (one system
    only) returns the transformed, deduplicated list in first-insertion order instead.  ``promote_f64`` (nsf with a torch
    normalisation): weight and sum the fp32 normalised scores in fp64 - NumPy 1.x's ``np.float32 * float`` - instead of fp32
    (NumPy >= 2, NEP 50).
    """
    lib = _lib.load()
    if method not in _lib.FUSE_METHODS:
        raise FusionB200Error(f"unknown fusion method {method!r}")
    if method == "nsf" and normalization not in _lib.FUSE_NORMS:
        raise FusionB200Error(f"unknown normalization {normalization!r}")
    s = len(lists)
    dev = lists[0][0].device
    q = lists[0][0].shape[0]
    ids_t, sc_t, len_t, strides, is64 = [], [], [], [], []
    for i, (ids, sc, lens) in enumerate(lists):
        f64 = sc.dtype == torch.float64
        ids_t.append(_req(ids, torch.int32, f"ids[{i}]"))
        sc_t.append(_req(sc, torch.float64 if f64 else torch.float32, f"scores[{i}]"))
        len_t.append(None if lens is None else _req(lens, torch.int32, f"lens[{i}]"))
        if ids.shape != sc.shape or ids.shape[0] != q:
            raise FusionB200Error("ranked lists of different systems cover different numbers of queries")
        strides.append(ids.shape[1])
        is64.append(1 if f64 else 0)
    total = sum(strides)
    u = out_stride or total
    norm_code = _lib.FUSE_NORMS[normalization] if method == "nsf" else 0
    need_distr = method == "nsf" and norm_code in (4, 5)
    distr_t = []
    if need_distr:
        if distributions is None or len(distributions) != s:
            raise FusionB200Error("percentile normalisation needs one distribution per system")
        for d in distributions:
            d = torch.as_tensor(d).to(device=dev, dtype=torch.float32).contiguous()
            if d.numel() > 1 and bool((d[1:] < d[:-1]).any()):
                raise FusionB200Error("percentile distributions must be ascending")
            distr_t.append(d)
    arr_p = C.c_void_p * s
    arr_i = C.c_int32 * s
    arr_d = C.c_double * s
    w = [1.0] * s if weights is None else [float(x) for x in weights]
    out_ids = torch.empty((q, u), dtype=torch.int32, device=dev)
    out_sc = torch.empty((q, u), dtype=torch.float64, device=dev)
    out_len = torch.empty((q,), dtype=torch.int32, device=dev)
    stride_arr = arr_i(*strides)
    per_q = lib.fz_fuse_workspace_bytes(s, 1, stride_arr)
    step = q if per_q == 0 else max(1, min(q, max_ws_bytes // per_q))
    for lo in range(0, q, step):
        hi = min(q, lo + step)
        ws = _ws(lib.fz_fuse_workspace_bytes(s, hi - lo, stride_arr), dev)
        check(lib.fz_fuse(
            arr_p(*[t[lo:hi].data_ptr() for t in ids_t]), arr_p(*[t[lo:hi].data_ptr() for t in sc_t]),
            arr_p(*[0 if t is None else t[lo:hi].data_ptr() for t in len_t]), arr_i(*is64), stride_arr, s, hi - lo,
            _lib.FUSE_METHODS[method] | (0x100 if keep_order else 0) | (0x200 if promote_f64 else 0), norm_code, arr_d(*w),
            arr_p(*[d.data_ptr() for d in distr_t]) if need_distr else None,
            arr_i(*[d.numel() for d in distr_t]) if need_distr else None,
            _ptr(out_ids[lo:hi]), _ptr(out_sc[lo:hi]), _ptr(out_len[lo:hi]), u, _ptr(ws), ws.numel(), _stream(out_ids)),
            "fz_fuse")
    return out_ids, out_sc, out_len


# ----------------------------------------------------------------------------------------------- metrics / weight sweep
RECALL_KS = (5, 10, 20, 50, 100, 200, 500, 1000)      # run_evaluation, src/retrievers/hybrid.py:27
MAP_KS = MRR_KS = NDCG_KS = (10, 100)


def metric_names(recall_ks=RECALL_KS, map_ks=MAP_KS, mrr_ks=MRR_KS, ndcg_ks=NDCG_KS) -> list[str]:
    return ([f"recall@{k}" for k in recall_ks] + [f"map@{k}" for k in map_ks] + [f"mrr@{k}" for k in mrr_ks] +
            [f"ndcg@{k}" for k in ndcg_ks] + ["r-precision"])


def _ks(*lists):
    out = []
    for ks in lists:
        arr = (C.c_int32 * max(1, len(ks)))(*ks)
        out += [arr, len(ks)]
    return out


MAX_GOLD_PER_QUERY = 256      # kMaxGold in csrc/metrics.cu


def _check_gold_limit(out: torch.Tensor) -> None:
    if bool(torch.isnan(out).any()):
        raise FusionB200Error(f"a query has more than {MAX_GOLD_PER_QUERY} distinct relevant documents: the device metric "
                              "kernels hold a query's gold ids in shared memory (the result would be wrong, not truncated)")


def rank_metrics(ids: torch.Tensor, lens: torch.Tensor | None, gold_ptr: torch.Tensor, gold_ids: torch.Tensor,
                 recall_ks=RECALL_KS, map_ks=MAP_KS, mrr_ks=MRR_KS, ndcg_ks=NDCG_KS) -> torch.Tensor:
    """Mean retrieval metrics of ranked id lists [Q, n] against gold id lists (CSR) -> float64 [M], ``metric_names`` order
    (src/utils/metrics.py:40-58)."""
    lib = _lib.load()
    ids = _req(ids, torch.int32, "ids")
    gold_ptr = _req(gold_ptr, torch.int32, "gold_ptr")
    gold_ids = _req(gold_ids, torch.int32, "gold_ids")
    if lens is not None:
        lens = _req(lens, torch.int32, "lens")
    q, n = ids.shape
    m = len(recall_ks) + len(map_ks) + len(mrr_ks) + len(ndcg_ks) + 1
    out = torch.empty(m, dtype=torch.float64, device=ids.device)
    per_q = torch.empty((max(q, 1), m), dtype=torch.float64, device=ids.device)       # summed in query order: reproducible bits
    check(lib.fz_rank_metrics(_ptr(ids), _ptr(lens), q, n, _ptr(gold_ptr), _ptr(gold_ids), *_ks(recall_ks, map_ks, mrr_ks, ndcg_ks),
                              _ptr(out), _ptr(per_q), _stream(ids)), "fz_rank_metrics")
    _check_gold_limit(out)
    return out / max(q, 1)


def fuse_sweep(lists, normalization: str | None, weights: torch.Tensor, gold_ptr: torch.Tensor, gold_ids: torch.Tensor,
               distributions=None, recall_ks=RECALL_KS, map_ks=MAP_KS, mrr_ks=MRR_KS, ndcg_ks=NDCG_KS) -> torch.Tensor:
    """Linear-fusion weight sweep (src/retrievers/hybrid.py:404-426): mean metrics of the nsf fusion of ``lists`` for every
    row of ``weights`` [W, S] -> float64 [W, M].  ``lists`` as for :func:`fuse`; each system is normalised once
    (weight-independent), then one kernel evaluates all W weight vectors per query."""
    lib = _lib.load()
    s = len(lists)
    dev = lists[0][0].device
    q = lists[0][0].shape[0]
    norm = [], [], []
    for i, triple in enumerate(lists):
        nid, nval, nlen = fuse([triple], "nsf", normalization, [1.0], None if distributions is None else [distributions[i]],
                               keep_order=True)
        norm[0].append(nid)
        norm[1].append(nval)
        norm[2].append(nlen)
    weights = _req(torch.as_tensor(weights, dtype=torch.float64, device=dev), torch.float64, "weights")
    if weights.dim() != 2 or weights.shape[1] != s:
        raise FusionB200Error("weights must be [n_weights, n_systems]")
    gold_ptr = _req(gold_ptr, torch.int32, "gold_ptr")
    gold_ids = _req(gold_ids, torch.int32, "gold_ids")
    m = len(recall_ks) + len(map_ks) + len(mrr_ks) + len(ndcg_ks) + 1
    out = torch.empty((weights.shape[0], m), dtype=torch.float64, device=dev)
    arr_p, arr_i = C.c_void_p * s, C.c_int32 * s
    f32_path = 0 if normalization in (None, "none") else 1
    check(lib.fz_fuse_sweep(arr_p(*[t.data_ptr() for t in norm[0]]), arr_p(*[t.data_ptr() for t in norm[1]]),
                            arr_p(*[t.data_ptr() for t in norm[2]]), arr_i(*[t.shape[1] for t in norm[0]]), s, q, f32_path,
                            _ptr(weights), weights.shape[0], _ptr(gold_ptr), _ptr(gold_ids),
                            *_ks(recall_ks, map_ks, mrr_ks, ndcg_ks), _ptr(out), _stream(out)), "fz_fuse_sweep")
    _check_gold_limit(out)
    return out / max(q, 1)


# ----------------------------------------------------------------------------------------------- K2
@dataclass
class PostingsView:
    """Device postings in the three storage forms of ``fz_postings_t`` (built by fusion_b200.index.build_postings)."""
    term_ptr: torch.Tensor        # int64 [V+1]   short lists
    post_doc: torch.Tensor        # int32
    post_val: torch.Tensor        # float64 | float32
    short_coarse: torch.Tensor    # uint16 payload in int16 [V, n_coarse+1]
    term_slot: torch.Tensor       # int32 [V]
    tiled_base: torch.Tensor      # int64 [n_tiled]
    tiled_tile_off: torch.Tensor  # uint32 payload in int32 [n_tiled, n_tiles+1]
    tiled_off: torch.Tensor       # uint16 payload in int16
    tiled_val: torch.Tensor
    dense_val: torch.Tensor       # [n_dense, n_tiles*tile_docs]
    n_docs: int
    tile_docs: int

    @property
    def n_tiles(self) -> int:
        return (self.n_docs + self.tile_docs - 1) // self.tile_docs

    @property
    def dtype(self) -> torch.dtype:
        return self.post_val.dtype

    def tensors(self):
        return (self.term_ptr, self.post_doc, self.post_val, self.short_coarse, self.term_slot, self.tiled_base, self.tiled_tile_off,
                self.tiled_off, self.tiled_val, self.dense_val)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def c_struct(self) -> _lib.Postings:
        p = lambda t: t.data_ptr() if t.numel() else 0
        return _lib.Postings(self.term_ptr.data_ptr(), p(self.post_doc), p(self.post_val), p(self.short_coarse),
                             self.term_slot.data_ptr(),
                             p(self.tiled_base), p(self.tiled_tile_off), p(self.tiled_off), p(self.tiled_val),
                             p(self.dense_val), self.n_tiles * self.tile_docs, self.term_ptr.numel() - 1,
                             self.tiled_base.numel(), self.dense_val.shape[0], self.tile_docs, self.n_tiles,
                             (self.n_tiles + COARSE_TILES - 1) // COARSE_TILES, self.n_docs)


MAX_QUERY_TERMS = 128     # kMaxTerms in csrc/sparse.cu: the terms of a query are held once per CTA


@dataclass
class ShardSync:
    """Cross-shard threshold exchange of a sharded corpus (``fz_shard_sync_t``): ``reduce_min`` all-reduces a [Q] tensor
    over the shards in place (element-wise MIN, on the current stream); ``sched_docs`` is the largest shard's size, so
    that every shard runs the same number of rounds and therefore the same number of collectives."""
    reduce_min: object
    n_shards: int
    sched_docs: int


class _SyncCall:
    """Owns the exchange tensor and the ctypes callback of one library call."""

    def __init__(self, sync: "ShardSync | None", nq: int, dtype, device, k_global: int = 0):
        self.error = None
        self.struct = None
        if sync is None or sync.n_shards <= 1:
            return
        self.exchange = torch.empty((nq,), dtype=dtype, device=device)

        def hook(_user):
            try:
                sync.reduce_min(self.exchange)
                return 0
            except BaseException as e:      # an exception must not unwind through the C frames
                self.error = e
                return 1

        self._cb = _lib.SHARD_HOOK(hook)
        # the published rank refers to the GLOBAL k: a shard smaller than k runs with k_eff < k and must not count as
        # ceil(k_eff / G) documents towards the k the floor has to cover
        floor_rank = -(-int(k_global) // int(sync.n_shards)) if k_global else 0
        self.struct = _lib.ShardSync(self._cb, None, self.exchange.data_ptr(), int(sync.n_shards), floor_rank,
                                     int(sync.sched_docs))

    def ref(self):
        return C.byref(self.struct) if self.struct is not None else None

    def reraise(self):
        if self.error is not None:
            raise self.error


FULL_ROWS_MAX_BYTES = 4 << 30         # score rows materialised for queries longer than MAX_QUERY_TERMS


def _long_queries(q_ptr: torch.Tensor) -> torch.Tensor:
    """Indices of the queries with more than MAX_QUERY_TERMS terms (tokens, duplicates counted); usually empty."""
    if q_ptr.numel() <= 1:
        return torch.zeros(0, dtype=torch.long, device=q_ptr.device)
    return torch.nonzero((q_ptr[1:] - q_ptr[:-1]) > MAX_QUERY_TERMS).flatten()


def _long_query_topk(pv: "PostingsView", q_ptr, q_term, q_weight, sel, k: int, doc_base: int):
    """Queries longer than the kernels' per-CTA term table: score every document in chunks of MAX_QUERY_TERMS terms
    (``sparse_scores`` continues the sums in query order, bit-identical to one pass) and rank the rows.  The reference has
    no length limit (bm25.py:149-156 loops over query.split()); long legal questions and unpruned SPLADE queries reach it."""
    per_q = pv.n_docs * pv.post_val.element_size()
    step = max(1, FULL_ROWS_MAX_BYTES // max(per_q, 1))
    out_s, out_i = [], []
    for lo in range(0, sel.numel(), step):
        p2, t2, w2 = _subset_queries(q_ptr, q_term, q_weight, sel[lo:lo + step])
        s2, i2 = rank_rows(sparse_scores(pv, p2, t2, w2), k, doc_base)
        out_s.append(s2)
        out_i.append(i2)
    return torch.cat(out_s), torch.cat(out_i)


def _sparse_topk_once(pv: PostingsView, q_ptr, q_term, q_weight, k, doc_base, cap, growth, sign_mode, sync=None,
                      k_global: int = 0):
    lib = _lib.load()
    f64 = pv.dtype == torch.float64
    dev = pv.term_ptr.device
    nq = q_ptr.numel() - 1
    sc = _SyncCall(sync, nq, pv.dtype, dev, k_global)
    out_s = torch.empty((nq, k), dtype=pv.dtype, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    status = torch.empty((nq,), dtype=torch.int32, device=dev)
    ws = _ws(lib.fz_sparse_topk_workspace_bytes(nq, k, cap, 1 if f64 else 0), dev)
    st = pv.c_struct()
    if f64:
        rc = lib.fz_sparse_topk_f64(C.byref(st), _ptr(q_ptr), _ptr(q_term), nq, k, doc_base, cap, growth, sign_mode,
                                    _ptr(out_s), _ptr(out_i), _ptr(status), _ptr(ws), ws.numel(), sc.ref(), _stream(out_s))
    else:
        rc = lib.fz_sparse_topk_f32(C.byref(st), _ptr(q_ptr), _ptr(q_term), _ptr(q_weight), nq, k, doc_base, cap,
                                    growth, sign_mode, _ptr(out_s), _ptr(out_i), _ptr(status), _ptr(ws), ws.numel(),
                                    sc.ref(), _stream(out_s))
    sc.reraise()
    check(rc, "fz_sparse_topk_f64" if f64 else "fz_sparse_topk_f32")
    return out_s, out_i, status


def _subset_queries(q_ptr, q_term, q_weight, sel):
    """CSR rows ``sel`` of the query matrix (host-side index arithmetic on a handful of flagged queries)."""
    ptr = q_ptr.cpu()
    sel_l = sel.cpu().tolist()
    lens = [int(ptr[i + 1] - ptr[i]) for i in sel_l]
    idx = torch.cat([torch.arange(int(ptr[i]), int(ptr[i + 1])) for i in sel_l]) if sum(lens) else torch.zeros(0, dtype=torch.long)
    new_ptr = torch.zeros(len(sel_l) + 1, dtype=torch.int32)
    new_ptr[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
    idx = idx.to(q_term.device)
    return (new_ptr.to(q_term.device), q_term[idx].contiguous(),
            None if q_weight is None else q_weight[idx].contiguous())


def sparse_topk(pv: PostingsView, q_ptr, q_term, q_weight, k: int, doc_base: int = 0, cap: int = DEFAULT_CAP,
                growth: int = DEFAULT_GROWTH, sync: ShardSync | None = None, defer: bool = False):
    """Top-k of sum_t w_qt * val[t, d] over an inverted index -> (scores [Q,k], ids [Q,k] int32).

    Order: score desc, ties by lower doc id; docs that match nothing score 0 and fill up in doc-id order; negative
    scores (BM25 idf < 0 for df > N/2) rank after the zeros - exactly the reference's `sorted(..., reverse=True)`
    over every document (src/retrievers/bm25.py:100-106).

    ``sync`` (corpus sharded over several GPUs): the shards agree on a per-query score floor between rounds, so each
    keeps only what can still reach the GLOBAL top-k; the local list may then hold fewer than k real entries (padded
    with (-inf, -1)), the merge of the shards' lists is unchanged.  The rare re-runs below never use it: they are
    decided per rank and must not issue collectives.

    ``defer``: return ``(scores, ids, fixup)`` without reading the per-query status back; ``fixup()`` reads it (the one
    host synchronisation of the call) and repairs the flagged rows in place.  A pipeline of several retrievers launches
    them all first and then runs the fix-ups: one pipeline drain per step instead of one per retriever."""
    q_ptr = _req(q_ptr, torch.int32, "q_ptr")
    q_term = _req(q_term, torch.int32, "q_term")
    if q_weight is not None:
        q_weight = _req(q_weight, torch.float32, "q_weight")
    k_eff = min(k, pv.n_docs)
    cap = max(cap, 2 * k_eff)
    out_s, out_i, status = _sparse_topk_once(pv, q_ptr, q_term, q_weight, k_eff, doc_base, cap, growth, +1, sync, k)

    def fixup():
        _sparse_topk_fixup(pv, q_ptr, q_term, q_weight, k_eff, doc_base, cap, out_s, out_i, status)
    if defer:
        return out_s, out_i, fixup
    fixup()
    return out_s, out_i


def _sparse_topk_fixup(pv, q_ptr, q_term, q_weight, k_eff, doc_base, cap, out_s, out_i, status):
    st = status.cpu()
    too_long = (st & FZ_STATUS_TOO_LONG) != 0
    if bool(too_long.any()):        # rare: more terms than a CTA holds - every document scored in term chunks, rows ranked
        sel = torch.nonzero(too_long).flatten()
        s2, i2 = _long_query_topk(pv, q_ptr, q_term, q_weight, sel, k_eff, doc_base)
        seld = sel.to(out_s.device)
        out_s[seld], out_i[seld] = s2, i2
        st[sel] = 0
    over = (st & FZ_STATUS_OVERFLOW) != 0
    if bool(over.any()):            # rare: redo those queries with rounds that cannot overflow
        sel = torch.nonzero(over).flatten()
        p2, t2, w2 = _subset_queries(q_ptr, q_term, q_weight, sel)
        s2, i2, st2 = _sparse_topk_once(pv, p2, t2, w2, k_eff, doc_base, cap, 1, +1)
        if bool(((st2.cpu() & FZ_STATUS_OVERFLOW) != 0).any()):
            raise FusionB200Error("sparse top-k overflowed even with conservative rounds")
        seld = sel.to(out_s.device)
        out_s[seld], out_i[seld] = s2, i2
        st[sel] = st2.cpu()
    need_neg = (st & FZ_STATUS_NEED_NEG) != 0
    if bool(need_neg.any()):        # positives + zero-score docs < k: the tail is the best of the negative scores
        sel = torch.nonzero(need_neg).flatten()
        p2, t2, w2 = _subset_queries(q_ptr, q_term, q_weight, sel)
        s2, i2, st2 = _sparse_topk_once(pv, p2, t2, w2, k_eff, doc_base, cap, 1, -1)
        for j, qi in enumerate(sel.tolist()):
            have = int((out_i[qi] >= 0).sum())
            take = k_eff - have
            out_s[qi, have:] = s2[j, :take]
            out_i[qi, have:] = i2[j, :take]


@dataclass
class SpladeHeadView:
    """Device tensors of ``fz_splade_head_t`` + the tail-only inverted index (built by fusion_b200.index.SparseIndex)."""
    head_bf16: torch.Tensor       # bf16 [N, head_dim]
    term_head: torch.Tensor       # int32 [V]
    term_max: torch.Tensor        # float32 [V]
    doc_ptr: torch.Tensor         # int64 [N+1]
    doc_post: torch.Tensor        # int32 [nnz, 2]: (term, weight bits)
    tail: PostingsView            # tail terms only, no dense rows, tile_docs % 256 == 0
    head_dim: int
    n_terms: int
    n_docs: int
    unit_rows: bool = False       # cos_sim index: every doc vector has norm <= 1
    boot: "PostingsView | None" = None    # general index over the first boot.n_docs docs (threshold bootstrap)

    def nbytes(self) -> int:
        return (self.tail.nbytes() + (self.boot.nbytes() if self.boot is not None else 0) +
                sum(t.numel() * t.element_size() for t in (self.head_bf16, self.term_head, self.term_max)))

    def c_struct(self) -> _lib.SpladeHead:
        return _lib.SpladeHead(self.head_bf16.data_ptr(), self.term_head.data_ptr(), self.term_max.data_ptr(),
                               self.doc_ptr.data_ptr(), self.doc_post.data_ptr(), self.head_dim, self.n_terms, self.n_docs,
                               1 if self.unit_rows else 0, 0)


SPLADE_GROWTH = 2               # rounds grow 2x: a round emits the docs whose score UPPER BOUND beats the running k-th score
SPLADE_MAX_ROUND_DOCS = 1 << 21    # bounds the code buffer: n_queries * max_round_docs / 2 bytes


def splade_topk(index, q_ptr, q_term, q_weight, k: int, doc_base: int = 0, cap: int = DEFAULT_CAP,
                growth: int = SPLADE_GROWTH, sync: ShardSync | None = None, max_round_docs: int = SPLADE_MAX_ROUND_DOCS,
                defer: bool = False):
    """SPLADE top-k through ``fz_splade_topk``: head terms on the tensor cores, a 4-bit upper bound of the tail sum per
    (query, doc), exact fp32 rescoring of the survivors.  ``index``: a ``fusion_b200.index.SparseIndex`` with a head/tail
    split.  Same result contract as :func:`sparse_topk` (scores exact fp32 sparse dot products, score desc, ties by lower
    doc id, zero-score docs fill up in doc-id order).  Queries the fast path hands back (negative weights, fewer than k
    positive-score docs, candidate-buffer overflow) are re-run on the general inverted index - never with ``sync``."""
    lib = _lib.load()
    hv: SpladeHeadView = index.head
    q_ptr = _req(q_ptr, torch.int32, "q_ptr")
    q_term = _req(q_term, torch.int32, "q_term")
    if q_weight is not None:
        q_weight = _req(q_weight, torch.float32, "q_weight")
    dev = hv.head_bf16.device
    nq = q_ptr.numel() - 1
    k_eff = min(k, hv.n_docs)
    cap = max(cap, 2 * k_eff)
    out_s = torch.empty((nq, k_eff), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k_eff), dtype=torch.int32, device=dev)
    status = torch.empty((nq,), dtype=torch.int32, device=dev)
    if nq == 0:
        return (out_s, out_i, lambda: None) if defer else (out_s, out_i)
    round_docs = max(256, min(int(max_round_docs), (hv.n_docs + 255) // 256 * 256))
    ws = _ws(lib.fz_splade_topk_workspace_bytes(nq, k_eff, cap, hv.head_dim, round_docs), dev)
    sc = _SyncCall(sync, nq, torch.float32, dev, k)
    tail, head = hv.tail.c_struct(), hv.c_struct()
    boot = hv.boot.c_struct() if hv.boot is not None else None
    rc = lib.fz_splade_topk(C.byref(tail), C.byref(head), C.byref(boot) if boot is not None else None, _ptr(q_ptr), _ptr(q_term), _ptr(q_weight), nq, k_eff, doc_base, cap,
                            growth, _ptr(out_s), _ptr(out_i), _ptr(status), _ptr(ws), ws.numel(), sc.ref(), _stream(out_s))
    sc.reraise()
    check(rc, "fz_splade_topk")
    def fixup():
        bad = (status & (FZ_STATUS_OVERFLOW | FZ_STATUS_FALLBACK)) != 0
        bad[_long_queries(q_ptr)] = True   # more terms than the kernels' per-query tables: the general path scores them in chunks
        if bool(bad.any()):
            sel = torch.nonzero(bad).flatten()
            p2, t2, w2 = _subset_queries(q_ptr, q_term, q_weight, sel)
            s2, i2 = sparse_topk(index.view(), p2, t2, w2, k_eff, doc_base, cap=cap)
            out_s[sel], out_i[sel] = s2, i2
    if defer:
        return out_s, out_i, fixup
    fixup()
    return out_s, out_i


def sparse_scores(pv: PostingsView, q_ptr, q_term, q_weight=None):
    """Score of every document for every query: [Q, n_docs] (fp64 for lexical impacts, fp32 for SPLADE weights).
    Queries of more than MAX_QUERY_TERMS terms are scored in several passes over chunks of their terms; every pass
    continues the row's sums in query order, so the result is bit-identical to one pass over all terms."""
    lib = _lib.load()
    q_ptr = _req(q_ptr, torch.int32, "q_ptr")
    q_term = _req(q_term, torch.int32, "q_term")
    if q_weight is not None:
        q_weight = _req(q_weight, torch.float32, "q_weight")
    nq = q_ptr.numel() - 1
    f64 = pv.dtype == torch.float64
    out = torch.empty((nq, pv.n_docs), dtype=pv.dtype, device=pv.term_ptr.device)
    st = pv.c_struct()
    max_len = int((q_ptr[1:] - q_ptr[:-1]).max()) if nq else 0
    n_pass = max(1, -(-max_len // MAX_QUERY_TERMS))
    for p in range(n_pass):
        if n_pass == 1:
            ptr_p, term_p, w_p = q_ptr, q_term, q_weight
        else:       # pass p: terms [128 p, 128 (p + 1)) of every query (empty for the short ones)
            lens = (q_ptr[1:] - q_ptr[:-1]).long()
            lo = torch.clamp(lens, max=p * MAX_QUERY_TERMS)
            hi = torch.clamp(lens, max=(p + 1) * MAX_QUERY_TERMS)
            cnt = hi - lo
            ptr_p = torch.zeros(nq + 1, dtype=torch.int32, device=q_ptr.device)
            ptr_p[1:] = torch.cumsum(cnt, 0).to(torch.int32)
            row = torch.repeat_interleave(torch.arange(nq, device=q_ptr.device), cnt)
            pos = torch.arange(int(cnt.sum()), device=q_ptr.device) - ptr_p[:-1].long()[row] + lo[row] + q_ptr[:-1].long()[row]
            term_p = q_term[pos].contiguous() if pos.numel() else torch.full((1,), -1, dtype=torch.int32, device=q_ptr.device)
            w_p = None if q_weight is None else (q_weight[pos].contiguous() if pos.numel() else torch.zeros(1, device=q_ptr.device))
        if f64:
            check(lib.fz_sparse_scores_f64(C.byref(st), _ptr(ptr_p), _ptr(term_p), nq, _ptr(out), 1 if p else 0, _stream(out)),
                  "fz_sparse_scores_f64")
        else:
            check(lib.fz_sparse_scores_f32(C.byref(st), _ptr(ptr_p), _ptr(term_p), _ptr(w_p), nq, _ptr(out), 1 if p else 0,
                                           _stream(out)), "fz_sparse_scores_f32")
    return out


def lexical_impacts(term_ptr, post_doc, post_tf, doc_len, idf, avgdl: float, k1: float, b: float, variant: int):
    lib = _lib.load()
    out = torch.empty(post_doc.numel(), dtype=torch.float64, device=post_doc.device)
    check(lib.fz_lexical_impacts(_ptr(_req(term_ptr, torch.int64, "term_ptr")), _ptr(_req(post_doc, torch.int32, "post_doc")),
                                 _ptr(_req(post_tf, torch.int32, "post_tf")),
                                 _ptr(None if doc_len is None else _req(doc_len, torch.int32, "doc_len")),
                                 _ptr(_req(idf, torch.float64, "idf")), term_ptr.numel() - 1, post_doc.numel(),
                                 float(avgdl), float(k1), float(b), variant, _ptr(out), _stream(out)), "fz_lexical_impacts")
    return out


# ----------------------------------------------------------------------------------------------- index build
def build_lexical_postings(doc_ptr: torch.Tensor, doc_tok: torch.Tensor, vocab: int):
    """Token ids per document -> term-major postings: (term_ptr int64 [V+1], post_doc int32, post_tf int32, doc_len int32)
    (``fz_build_lexical_*``; the reference builds dicts, bm25.py:53-83)."""
    lib = _lib.load()
    doc_ptr = _req(doc_ptr, torch.int64, "doc_ptr")
    doc_tok = _req(doc_tok, torch.int32, "doc_tok")
    dev = doc_ptr.device
    n_docs, n_tok = doc_ptr.numel() - 1, doc_tok.numel()
    nbytes = lib.fz_build_lexical_workspace_bytes(n_tok, vocab)
    ws = _ws(nbytes, dev)
    nnz = C.c_int64(0)
    check(lib.fz_build_lexical_plan(_ptr(doc_ptr), _ptr(doc_tok), n_docs, n_tok, vocab, C.byref(nnz), _ptr(ws), nbytes,
                                    _stream(doc_ptr)), "fz_build_lexical_plan")
    nnz = int(nnz.value)
    term_ptr = torch.empty(vocab + 1, dtype=torch.int64, device=dev)
    post_doc = torch.empty(nnz, dtype=torch.int32, device=dev)
    post_tf = torch.empty(nnz, dtype=torch.int32, device=dev)
    doc_len = torch.empty(n_docs, dtype=torch.int32, device=dev)
    check(lib.fz_build_lexical_fill(_ptr(doc_ptr), n_docs, n_tok, vocab, nnz, _ptr(term_ptr), _ptr(post_doc), _ptr(post_tf),
                                    _ptr(doc_len), _ptr(ws), nbytes, _stream(doc_ptr)), "fz_build_lexical_fill")
    return term_ptr, post_doc, post_tf, doc_len


def build_term_major(row: torch.Tensor, term: torch.Tensor, n_rows: int, n_terms: int):
    """(row, term) per entry -> (order int64 [nnz]: entry indices term-major / row-ascending, term_ptr int64 [V+1])."""
    lib = _lib.load()
    row, term = _req(row, torch.int32, "row"), _req(term, torch.int32, "term")
    dev, nnz = row.device, row.numel()
    nbytes = lib.fz_build_term_major_workspace_bytes(nnz, n_terms)
    ws = _ws(nbytes, dev)
    term_ptr = torch.empty(n_terms + 1, dtype=torch.int64, device=dev)
    order = torch.empty(nnz, dtype=torch.int64, device=dev)
    check(lib.fz_build_term_major(_ptr(row), _ptr(term), nnz, n_rows, n_terms, _ptr(term_ptr), _ptr(order), _ptr(ws), nbytes,
                                  _stream(row)), "fz_build_term_major")
    return order, term_ptr


def build_postings(term_ptr: torch.Tensor, post_doc: torch.Tensor, post_val: torch.Tensor, n_docs: int, tile_docs: int,
                   tiled_min: int, dense_min: int) -> "PostingsView":
    """Term-major CSR -> the three storage forms of ``fz_postings_t`` (``fz_build_postings_plan`` / ``_fill``)."""
    lib = _lib.load()
    term_ptr = _req(term_ptr, torch.int64, "term_ptr")
    post_doc = _req(post_doc, torch.int32, "post_doc")
    if post_val.dtype not in (torch.float32, torch.float64):
        raise FusionB200Error(f"post_val must be float32 or float64, got {post_val.dtype}")
    post_val = _req(post_val, post_val.dtype, "post_val")
    dev, n_terms, nnz = term_ptr.device, term_ptr.numel() - 1, post_doc.numel()
    nbytes = lib.fz_build_postings_workspace_bytes(n_terms)
    ws = _ws(nbytes, dev)
    short_ptr = torch.empty(n_terms + 1, dtype=torch.int64, device=dev)
    term_slot = torch.empty(n_terms, dtype=torch.int32, device=dev)
    plan = _lib.BuildPlan()
    check(lib.fz_build_postings_plan(_ptr(term_ptr), _ptr(post_doc), n_terms, n_docs, tile_docs, int(tiled_min), int(dense_min),
                                     _ptr(short_ptr), _ptr(term_slot), C.byref(plan), _ptr(ws), nbytes, _stream(term_ptr)),
          "fz_build_postings_plan")
    vt = post_val.dtype
    short_doc = torch.empty(plan.n_short, dtype=torch.int32, device=dev)
    short_val = torch.empty(plan.n_short, dtype=vt, device=dev)
    short_coarse = torch.empty((n_terms, plan.n_coarse + 1), dtype=torch.int16, device=dev)
    tiled_base = torch.empty(plan.n_tiled, dtype=torch.int64, device=dev)
    tile_off = torch.empty((plan.n_tiled, plan.n_tiles + 1), dtype=torch.int32, device=dev)
    tiled_off = torch.empty(plan.n_tiled_entries, dtype=torch.int16, device=dev)
    tiled_val = torch.empty(plan.n_tiled_entries, dtype=vt, device=dev)
    dense_val = torch.empty((plan.n_dense, plan.dense_stride), dtype=vt, device=dev)
    check(lib.fz_build_postings_fill(_ptr(term_ptr), _ptr(post_doc), _ptr(post_val), post_val.element_size(), n_terms, n_docs,
                                     tile_docs, _ptr(short_ptr), _ptr(term_slot), C.byref(plan), nnz, _ptr(short_doc),
                                     _ptr(short_val), _ptr(short_coarse), _ptr(tiled_base), _ptr(tile_off), _ptr(tiled_off),
                                     _ptr(tiled_val), _ptr(dense_val), _ptr(ws), nbytes, _stream(term_ptr)),
          "fz_build_postings_fill")
    return PostingsView(short_ptr, short_doc, short_val, short_coarse, term_slot, tiled_base, tile_off, tiled_off, tiled_val,
                        dense_val, n_docs, tile_docs)


def build_csr_normalize(doc_ptr: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty_like(weight)
    if weight.numel() == 0:          # every row is empty: nothing to normalise
        return out
    check(lib.fz_build_csr_normalize(_ptr(_req(doc_ptr, torch.int64, "doc_ptr")), _ptr(_req(weight, torch.float32, "weight")),
                                     doc_ptr.numel() - 1, _ptr(out), _stream(weight)), "fz_build_csr_normalize")
    return out


def build_term_stats(term: torch.Tensor, weight: torch.Tensor, n_terms: int):
    """-> (df int64 [V], term_max float32 [V], flags int: 1 = negative weight present, 2 = term id out of range)."""
    lib = _lib.load()
    term, weight = _req(term, torch.int32, "term"), _req(weight, torch.float32, "weight")
    dev = term.device
    df = torch.empty(n_terms, dtype=torch.int64, device=dev)
    tmax = torch.empty(n_terms, dtype=torch.float32, device=dev)
    flags = torch.empty(1, dtype=torch.int32, device=dev)
    check(lib.fz_build_term_stats(_ptr(term), _ptr(weight), term.numel(), n_terms, _ptr(df), _ptr(tmax), _ptr(flags),
                                  _stream(term)), "fz_build_term_stats")
    return df, tmax, int(flags.item())


def build_splade_head(doc_ptr: torch.Tensor, term: torch.Tensor, weight: torch.Tensor, term_head: torch.Tensor,
                      head_dim: int) -> torch.Tensor:
    lib = _lib.load()
    n_docs = doc_ptr.numel() - 1
    head = torch.empty((n_docs, head_dim), dtype=torch.bfloat16, device=doc_ptr.device)
    check(lib.fz_build_splade_head(_ptr(_req(doc_ptr, torch.int64, "doc_ptr")), _ptr(_req(term, torch.int32, "term")),
                                   _ptr(_req(weight, torch.float32, "weight")), n_docs, term.numel(),
                                   _ptr(_req(term_head, torch.int32, "term_head")), head_dim, _ptr(head), _stream(term)),
          "fz_build_splade_head")
    return head


# ----------------------------------------------------------------------------------------------- K1
def normalize_rows(x: torch.Tensor, normalize: bool = True, want_f32: bool = True, want_bf16: bool = True):
    """rows / max(||row||, 1e-12) -> (fp32 copy | None, bf16 copy | None)."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    n, d = x.shape
    o32 = torch.empty_like(x) if want_f32 else None
    o16 = torch.empty((n, d), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    check(lib.fz_normalize_rows(_ptr(x), n, d, 1 if normalize else 0, _ptr(o32), _ptr(o16), _stream(x)),
          "fz_normalize_rows")
    return o32, o16


def dense_topk(q_bf16, d_bf16, q_f32, d_f32, k: int, margin: float = 0.0, doc_base: int = 0,
               cap: int = DEFAULT_CAP, growth: int = DEFAULT_GROWTH, tau_reduce=None, n_shards: int = 1,
               sched_docs: int | None = None, defer: bool = False):
    """Exhaustive inner-product top-k (tcgen05 GEMM with the threshold filter in its epilogue).

    q_bf16 [Q, d], d_bf16 [N, d] are the tensor-core operands; with q_f32 / d_f32 the survivors within ``margin`` of
    the running k-th bf16 score are rescored in fp32 (exact mode).  -> (scores f32 [Q,k], ids int32 [Q,k]).
    ``tau_reduce`` (exact mode, corpus sharded over ``n_shards``): a callable that takes the element-wise MINIMUM of a [Q]
    tensor over the shards (one all-reduce).  Every shard reports its ceil(k / n_shards)-th best score; the minimum
    bounds the global k-th score from below, and candidates under it (minus the margin) are not rescored.  With
    ``sched_docs`` (largest shard size) the same exchange also runs between the filter rounds (``ShardSync``).
    ``defer``: return ``(scores, ids, fixup)``; the overflow check (a host synchronisation) and the rare re-run happen in
    ``fixup()`` - the first pass is finished optimistically, a re-run never issues a collective."""
    lib = _lib.load()
    q_bf16 = _req(q_bf16, torch.bfloat16, "q_bf16")
    d_bf16 = _req(d_bf16, torch.bfloat16, "d_bf16")
    exact = d_f32 is not None
    if exact:
        q_f32 = _req(q_f32, torch.float32, "q_f32")
        d_f32 = _req(d_f32, torch.float32, "d_f32")
    nq, dim = q_bf16.shape
    n = d_bf16.shape[0]
    k_eff = min(k, n)
    cap = max(cap, 2 * k_eff)
    out_s = torch.empty((nq, k_eff), dtype=torch.float32, device=q_bf16.device)
    out_i = torch.empty((nq, k_eff), dtype=torch.int32, device=q_bf16.device)
    status = torch.empty((nq,), dtype=torch.int32, device=q_bf16.device)
    ws = _ws(lib.fz_dense_topk_workspace_bytes(nq, k_eff, cap), q_bf16.device)
    staged = exact and tau_reduce is not None
    tau = torch.empty((nq,), dtype=torch.float32, device=q_bf16.device) if staged else None

    def run(g, first=False):
        if staged:
            # only the first attempt exchanges thresholds: a re-run is decided per rank and must not issue collectives
            sc = _SyncCall(ShardSync(tau_reduce, n_shards, sched_docs) if (first and sched_docs is not None) else None,
                           nq, torch.float32, q_bf16.device, k)
            rc = lib.fz_dense_topk_filter(_ptr(q_bf16), _ptr(d_bf16), nq, n, dim, k_eff, float(margin), doc_base, cap, g,
                                          min(k_eff, -(-k // max(1, n_shards))), _ptr(tau), _ptr(status), _ptr(ws),
                                          ws.numel(), sc.ref(), _stream(out_s))
            sc.reraise()
            check(rc, "fz_dense_topk_filter")
        else:
            check(lib.fz_dense_topk(_ptr(q_bf16), _ptr(d_bf16), _ptr(q_f32 if exact else None),
                                    _ptr(d_f32 if exact else None), nq, n, dim, k_eff, float(margin), doc_base, cap, g,
                                    _ptr(out_s), _ptr(out_i), _ptr(status), _ptr(ws), ws.numel(), _stream(out_s)),
                  "fz_dense_topk")

    # With a margin every round emits everything within `margin` of the running k-th score, about twice the k
    # survivors of the plain filter, so the doc ranges may only grow 3x per round instead of 4x.
    if margin > 0:
        growth = min(growth, 3)
    def finish(floor):
        check(lib.fz_dense_topk_finish(_ptr(q_f32), _ptr(d_f32), _ptr(floor), nq, dim, k_eff, doc_base, cap, _ptr(out_s),
                                       _ptr(out_i), _ptr(status), _ptr(ws), ws.numel(), _stream(out_s)),
              "fz_dense_topk_finish")

    run(growth, first=True)
    if staged:
        finish(tau_reduce(tau))             # every shard calls the collective exactly once, whatever happens below

    def fixup():
        if not bool(((status & FZ_STATUS_OVERFLOW) != 0).any()):
            return
        if margin > 0:
            run(2)
            if bool(((status & FZ_STATUS_OVERFLOW) != 0).any()):
                raise FusionB200Error("dense top-k candidate buffer overflowed: lower `margin` or raise `cap`")
        else:
            run(1)      # conservative rounds never overflow when margin == 0
        if staged:
            finish(None)                    # decided per rank: no floor, no collective
    if defer:
        return out_s, out_i, fixup
    fixup()
    return out_s, out_i


def dense_scores(q_f32, d_f32):
    """Exact fp32 [Q, N] score matrix on CUDA cores (full-ranking mode on small corpora)."""
    lib = _lib.load()
    q_f32 = _req(q_f32, torch.float32, "q_f32")
    d_f32 = _req(d_f32, torch.float32, "d_f32")
    out = torch.empty((q_f32.shape[0], d_f32.shape[0]), dtype=torch.float32, device=q_f32.device)
    check(lib.fz_dense_scores_f32(_ptr(q_f32), _ptr(d_f32), q_f32.shape[0], d_f32.shape[0], q_f32.shape[1], _ptr(out),
                                  _stream(out)), "fz_dense_scores_f32")
    return out


def pairwise_dot(q_f32, d_f32):
    """Row-wise dot products of two [B, d] fp32 matrices -> fp32 [B]."""
    lib = _lib.load()
    q_f32 = _req(q_f32, torch.float32, "q_f32")
    d_f32 = _req(d_f32, torch.float32, "d_f32")
    if q_f32.shape != d_f32.shape:
        raise FusionB200Error("pairwise similarity needs two matrices of the same shape")
    out = torch.empty((q_f32.shape[0],), dtype=torch.float32, device=q_f32.device)
    check(lib.fz_pairwise_dot_f32(_ptr(q_f32), _ptr(d_f32), q_f32.shape[0], q_f32.shape[1], _ptr(out), _stream(out)),
          "fz_pairwise_dot_f32")
    return out


# ----------------------------------------------------------------------------------------------- K3
def pack_tokens(tok_ptr: torch.Tensor, tok_emb_bf16: torch.Tensor):
    """Token store -> the packed image the MaxSim kernel streams (per passage two 64-dim halves of 128-byte rows, rows
    padded to 8, 16-byte chunks pre-swizzled for the UMMA layout).  -> (pk_ptr int64 [N+1] packed-row offsets, packed uint8)."""
    lib = _lib.load()
    tok_ptr = _req(tok_ptr, torch.int64, "tok_ptr")
    tok_emb_bf16 = _req(tok_emb_bf16, torch.bfloat16, "tok_emb")
    if tok_emb_bf16.shape[1] != 128:
        raise FusionB200Error("maxsim needs 128-dimensional token embeddings")
    lens = tok_ptr[1:] - tok_ptr[:-1]
    pk_ptr = torch.zeros_like(tok_ptr)
    pk_ptr[1:] = torch.cumsum((lens + 7) // 8 * 8, 0)
    total = int(pk_ptr[-1])
    if total >= (1 << 31):
        raise FusionB200Error("token store too large for one shard (2^31 packed rows)")
    packed = torch.empty(max(total, 8) * 256 + 1024, dtype=torch.uint8, device=tok_emb_bf16.device)
    shift = (-packed.data_ptr()) % 1024
    packed = packed[shift: shift + max(total, 8) * 256]                     # 1024-byte aligned view
    n_docs = tok_ptr.numel() - 1
    if n_docs:
        check(lib.fz_maxsim_pack(_ptr(tok_ptr), _ptr(tok_emb_bf16), _ptr(pk_ptr), n_docs, _ptr(packed), _stream(packed)),
              "fz_maxsim_pack")
    return pk_ptr, packed


def maxsim(q_tok_bf16, tok_ptr, tok_emb_bf16, cand_ids, doc_base: int = 0, packed=None):
    """ColBERT MaxSim of every (query, candidate) pair -> fp32 [Q, C].

    q_tok_bf16 [Q, Lq, 128], tok_ptr int64 [N+1], cand_ids int32 [Q, C] (global ids; candidates outside
    [doc_base, doc_base+N) are skipped and score 0).  ``packed`` = ``pack_tokens(tok_ptr, tok_emb)`` (TokenStore keeps
    it); without it the plain ``tok_emb_bf16`` [T, 128] is packed on the fly."""
    lib = _lib.load()
    q_tok_bf16 = _req(q_tok_bf16, torch.bfloat16, "q_tok")
    tok_ptr = _req(tok_ptr, torch.int64, "tok_ptr")
    cand_ids = _req(cand_ids, torch.int32, "cand_ids")
    nq, lq, dim = q_tok_bf16.shape
    if dim != 128:
        raise FusionB200Error("maxsim needs 128-dimensional token embeddings")
    if packed is None:
        packed = pack_tokens(tok_ptr, tok_emb_bf16)
    pk_ptr, pk = packed
    out = torch.empty(cand_ids.shape, dtype=torch.float32, device=cand_ids.device)
    ws = _ws(lib.fz_maxsim_workspace_bytes(nq, cand_ids.shape[1]), cand_ids.device)
    check(lib.fz_maxsim_bf16(_ptr(q_tok_bf16), lq, _ptr(cand_ids), _ptr(tok_ptr), _ptr(pk_ptr), _ptr(pk),
                             tok_ptr.numel() - 1, doc_base, nq, cand_ids.shape[1], _ptr(out), _ptr(ws), ws.numel(),
                             _stream(out)), "fz_maxsim_bf16")
    return out
