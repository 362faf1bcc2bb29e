"""Mirror of ``CustomSearcher.search_all`` (``src/utils/colbert_ir.py:211-255``) for the late-interaction arithmetic.

The reference builds a PLAID index (k-means centroids + 2-bit residuals) and delegates candidate generation and
MaxSim to colbert-ai.  Here the token embeddings are kept as a bf16 store and candidates are scored exactly by the
tcgen05 MaxSim kernel; candidate generation comes from another retriever (top-k rescoring, north-star config 4) or
is exhaustive on small corpora.
"""
from __future__ import annotations

import time

import torch

from ..index import TokenStore
from ..retrievers.hybrid import Ranker


class CustomSearcher:
    def __init__(self, store: TokenStore, encoder=None):
        self.store = store
        self.encoder = encoder

    def encode(self, queries: list[str]) -> torch.Tensor:
        return self.encoder.encode_queries(queries)

    def search_all_tensors(self, q_tok: torch.Tensor, k: int, cand_ids: torch.Tensor | None = None):
        return Ranker.maxsim_search_tensors(q_tok.cuda(), self.store, k, cand_ids)

    def search_all(self, queries: dict, k: int = 10, cand_ids: torch.Tensor | None = None):
        """queries: {qid: text}.  -> {qid: [(pid, rank, score), ...]} like ``Ranking.todict()`` (colbert_ir.py:245-255)."""
        qids = list(queries.keys())
        t0 = time.perf_counter()
        q_tok = self.encode(list(queries.values()))
        t1 = time.perf_counter()
        scores, ids = self.search_all_tensors(q_tok, k, cand_ids)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        n = max(len(qids), 1)
        print(f"Avg. latency (ms/query): {(t2 - t0) / n * 1000:.2f} (Encoding: {(t1 - t0) / n * 1000:.2f}; Scoring: {(t2 - t1) / n * 1000:.2f})")
        return {qid: [(pid, r + 1, s) for r, (pid, s) in enumerate(zip(ri, rs))]
                for qid, ri, rs in zip(qids, ids.cpu().tolist(), scores.cpu().tolist())}
