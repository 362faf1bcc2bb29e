"""Tensor-level hybrid retrieve-and-fuse pipeline: BM25 + DPR + SPLADE + ColBERT-MaxSim -> rank fusion.

This is the batched form of ``hybrid.main`` (src/retrievers/hybrid.py:344-358 retrieval, :455 fusion): all queries
of a batch go through the four scorers and one fusion kernel, with corpus shards spread over the ranks of a
``torch.distributed`` group.  Per retriever and rank: local top-k on the shard -> one all-to-all of k candidates per
query -> k-way merge of this rank's query slice; ColBERT rescoring and fusion then run query-sharded.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, fields

import torch

from . import ops, sharding
from .index import DenseIndex, LexicalIndex, SparseIndex, TokenStore

# launch order of the first-stage retrievers (every rank must use the same order: their collectives interleave)
STAGE_ORDER = tuple(os.environ.get("FZ_STAGE_ORDER", "dpr,splade,bm25").replace(":", ",").split(","))
assert sorted(STAGE_ORDER) == ["bm25", "dpr", "splade"], STAGE_ORDER


@dataclass
class HybridQueries:
    """One batch of queries in the form each scorer consumes (host-pinned or device tensors)."""
    lex_ptr: torch.Tensor | None = None      # int32 [Q+1]   BM25 query tokens (CSR), -1 = out of vocabulary
    lex_term: torch.Tensor | None = None     # int32 [nq]
    sp_ptr: torch.Tensor | None = None       # int32 [Q+1]   SPLADE query terms (CSR)
    sp_term: torch.Tensor | None = None      # int32 [ns]
    sp_weight: torch.Tensor | None = None    # float32 [ns]  (already L2-normalised for cos_sim)
    dense: torch.Tensor | None = None        # float32 [Q, d] raw query embeddings
    colbert: torch.Tensor | None = None      # bfloat16 [Q, Lq, 128] query token embeddings

    def to(self, device, non_blocking: bool = True) -> "HybridQueries":
        return HybridQueries(**{f.name: (None if getattr(self, f.name) is None else
                                         getattr(self, f.name).to(device, non_blocking=non_blocking)) for f in fields(self)})

    def to_sharded(self, device, group=None) -> tuple["HybridQueries", int]:
        """Host (pinned) -> device with the corpus sharded over the ranks of ``group``: every rank needs EVERY query, but
        uploading the whole batch on each rank sends it over PCIe G times.  The large row-major tensors (dense embeddings,
        ColBERT query tokens: 97 % of the bytes) are uploaded 1/G per rank and all-gathered over NVLink; the small CSR
        tensors are uploaded whole.  -> (device queries, bytes this rank copied host -> device)."""
        world, rank = sharding._world(group)
        if world == 1:
            return self.to(device), self.nbytes()
        out, h2d = {}, 0
        for f in fields(self):
            t = getattr(self, f.name)
            if t is None:
                out[f.name] = None
            elif f.name in ("dense", "colbert"):
                per = (t.shape[0] + world - 1) // world
                lo, hi = min(t.shape[0], rank * per), min(t.shape[0], (rank + 1) * per)
                part = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
                part[: hi - lo].copy_(t[lo:hi], non_blocking=True)
                h2d += (hi - lo) * t[0].numel() * t.element_size()
                out[f.name] = sharding.allgather_rows(part, group)[: t.shape[0]]
            else:
                out[f.name] = t.to(device, non_blocking=True)
                h2d += t.numel() * t.element_size()
        return HybridQueries(**out), h2d

    def pin(self) -> "HybridQueries":
        return HybridQueries(**{f.name: (None if getattr(self, f.name) is None else getattr(self, f.name).cpu().pin_memory())
                                for f in fields(self)})

    def nbytes(self) -> int:
        return sum(getattr(self, f.name).numel() * getattr(self, f.name).element_size()
                   for f in fields(self) if getattr(self, f.name) is not None)

    @property
    def n_queries(self) -> int:
        for t in (self.dense, self.colbert):
            if t is not None:
                return t.shape[0]
        for t in (self.lex_ptr, self.sp_ptr):
            if t is not None:
                return t.numel() - 1
        return 0


class HybridSearcher:
    """Owns this rank's shard of every index and runs retrieve -> (exchange + merge) -> rescore -> fuse."""

    def __init__(self, lexical: LexicalIndex | None = None, sparse: SparseIndex | None = None,
                 dense: DenseIndex | None = None, tokens: TokenStore | None = None, k: int = 1000,
                 fusion: str = "nsf", normalization: str | None = "z-score", weights: dict | None = None,
                 colbert_pool: int | None = None, dense_exact: bool = True, group=None, shard_sync: bool = True):
        self.lexical, self.sparse, self.dense, self.tokens = lexical, sparse, dense, tokens
        self.k, self.fusion, self.normalization, self.weights = k, fusion, normalization, weights
        self.colbert_pool, self.dense_exact, self.group = colbert_pool, dense_exact, group
        self.world, self.rank = sharding._world(group)
        # cross-shard threshold exchange (ops.ShardSync): every shard follows the round schedule of the largest shard
        self.shard_sync = shard_sync and self.world > 1
        self._sync = {}
        if self.shard_sync:
            idx = {"bm25": lexical, "splade": sparse, "dpr": dense}
            names = [n for n, ix in idx.items() if ix is not None]
            if names:           # (a fusion-only searcher holds no index: nothing to exchange)
                dev = torch.device("cuda", torch.cuda.current_device())
                sizes = sharding.allreduce_max_ints([idx[n].n_docs for n in names], dev, group)
                reduce = lambda t: sharding.allreduce_min(t, self.group)       # noqa: E731
                self._sync = {n: ops.ShardSync(reduce, self.world, m) for n, m in zip(names, sizes)}
        self.stage_ms: dict[str, float] = {}
        self.timing = False
        self._events = []

    # -- helpers ------------------------------------------------------------------------------------------
    def _timed(self, name, fn):
        if not self.timing:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self._events.append((name, a, b))
        return out

    def _merge(self, scores, ids):
        if self.world == 1:
            return scores, ids
        return sharding.exchange_merge_topk(scores, ids, self.k, self.group)

    # -- pipeline -----------------------------------------------------------------------------------------
    def retrieve(self, q: HybridQueries) -> dict[str, tuple[torch.Tensor, torch.Tensor]]:
        """-> {system: (scores [Qs, k], ids int32 [Qs, k])} for this rank's query slice, best first."""
        out = {}
        self._events = []
        # Execution order: the tensor-core GEMM first, then SPLADE, BM25 and MaxSim (which needs the GEMM's list) last:
        # the kernel that follows the GEMM starts under its power cap, and SPLADE's is the one with the most headroom.
        # The returned dict keeps the system order bm25, splade, dpr, colbert (fusion breaks ties by first insertion,
        # hybrid.py:294-307).
        # All three first-stage retrievers are LAUNCHED before any per-query status is read back: their fix-ups (overflow /
        # fallback re-runs, rare) run after one host synchronisation per step instead of one per retriever; then the merges.
        local, fixups = {}, []

        def run_dense():
            q32, q16 = self.dense.prepare_queries(q.dense)
            exact = self.dense_exact and self.dense.d_f32 is not None
            # sharded exact mode: the shards' ceil(k/G)-th scores bound the global k-th one from below, so they agree on
            # a floor (one all-reduce of Q floats) and rescore only what can still reach the global top-k
            reduce = (lambda t: sharding.allreduce_min(t, self.group)) if (exact and self.world > 1) else None
            return ops.dense_topk(q16, self.dense.d_bf16, q32 if exact else None, self.dense.d_f32 if exact else None,
                                  self.k, margin=self.dense.exact_margin(q32, q16) if exact else 0.0, doc_base=self.dense.doc_base,
                                  tau_reduce=reduce, n_shards=self.world,
                                  sched_docs=self._sync["dpr"].sched_docs if (reduce and "dpr" in self._sync) else None,
                                  defer=True)

        stages = {
            "dpr": (self.dense, run_dense),
            "splade": (self.sparse, lambda: self.sparse.topk(q.sp_ptr, q.sp_term, q.sp_weight, self.k,
                                                             sync=self._sync.get("splade"), defer=True)),
            "bm25": (self.lexical, lambda: ops.sparse_topk(self.lexical.view(), q.lex_ptr, q.lex_term, None, self.k,
                                                           self.lexical.doc_base, sync=self._sync.get("bm25"), defer=True)),
        }
        for name in STAGE_ORDER:
            index, run = stages[name]
            if index is not None:
                s, i, fx = self._timed(name, run)
                local[name] = (s, i)
                fixups.append(fx)
        self._timed("status_fixups", lambda: [fx() for fx in fixups])
        for name in STAGE_ORDER:
            if name in local:
                s, i = local[name]
                out[name] = self._timed(f"{name}_merge", lambda: self._merge(s, i))
        if self.tokens is not None:
            out["colbert"] = self._timed("colbert", lambda: self._colbert(q, out))
        return {name: out[name] for name in ("bm25", "splade", "dpr", "colbert") if name in out}

    def _colbert(self, q: HybridQueries, lists):
        """MaxSim-rescore the first available system's merged top-k candidates (north-star config 4)."""
        src = next(n for n in ("dpr", "bm25", "splade") if n in lists)
        cand = lists[src][1]                                            # [Qs, k] global ids of this rank's query slice
        cand_all = sharding.allgather_rows(cand, self.group)[: q.colbert.shape[0]] if self.world > 1 else cand
        pool_ids = cand_all if self.colbert_pool is None else torch.where(cand_all >= 0, cand_all % self.colbert_pool, cand_all)
        part = ops.maxsim(q.colbert, self.tokens.tok_ptr, None, pool_ids.contiguous(), self.tokens.doc_base,
                          packed=self.tokens.packed())
        sc = sharding.reduce_scatter_scores(part, self.group) if self.world > 1 else part
        sc = torch.where(cand >= 0, sc[: cand.shape[0]], torch.full_like(sc[: cand.shape[0]], float("-inf")))
        order_s, order_i = ops.rank_rows(sc.contiguous(), cand.shape[1], 0)
        return order_s, torch.gather(cand, 1, order_i.long())

    def fuse(self, lists: dict, out_k: int | None = None):
        names = list(lists.keys())
        w = None
        if self.fusion == "nsf":
            w = [(self.weights or {}).get(n, 1.0 / len(names)) for n in names]     # equal weights (hybrid.py:448)
        triples = []
        for n in names:
            s, i = lists[n]
            lens = (i >= 0).sum(dim=1).to(torch.int32)
            triples.append((i, s, lens))
        return ops.fuse(triples, self.fusion, self.normalization, w, out_stride=out_k or self.k)

    def search(self, q: HybridQueries, out_k: int | None = None):
        """Device tensors in -> (fused ids [Qs, k], fused scores f64 [Qs, k], lens [Qs]) for this rank's query slice."""
        lists = self.retrieve(q)
        fused = self._timed("fuse", lambda: self.fuse(lists, out_k))
        self.collect_stage_ms()
        return fused

    def collect_stage_ms(self):
        """Fold the CUDA-event stage timings of the last pass into ``stage_ms`` (only when ``timing`` is on)."""
        if self.timing:
            torch.cuda.synchronize()
            for name, a, b in self._events:
                self.stage_ms[name] = self.stage_ms.get(name, 0.0) + a.elapsed_time(b)
            self._events = []

    def search_host(self, q_host: HybridQueries, out_ids: torch.Tensor, out_scores: torch.Tensor, out_k: int | None = None):
        """End-to-end call with HOST buffers: pinned inputs -> device, search, fused top-k -> pinned outputs."""
        dev = torch.device("cuda", torch.cuda.current_device())
        q_dev, self.last_h2d_bytes = q_host.to_sharded(dev, self.group)
        ids, scores, lens = self.search(q_dev, out_k)
        out_ids[: ids.shape[0]].copy_(ids, non_blocking=True)
        out_scores[: scores.shape[0]].copy_(scores, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_ids, out_scores
