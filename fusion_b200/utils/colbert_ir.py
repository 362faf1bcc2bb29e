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
    """``search_all`` without a PLAID index: exhaustive MaxSim over the token store (this rank's shard of the collection when
    a process ``group`` is given; the shards' top-k lists are then merged over NCCL), or rescoring of given candidates."""

    def __init__(self, store: TokenStore, encoder=None, group=None):
        self.store = store
        self.encoder = encoder
        self.group = group

    def encode(self, queries: list[str]) -> torch.Tensor:
        return self.encoder.encode_queries(queries)

    def search_all_tensors(self, q_tok: torch.Tensor, k: int, cand_ids: torch.Tensor | None = None):
        return Ranker.maxsim_search_tensors(q_tok.cuda(), self.store, k, cand_ids, group=self.group)

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


# ------------------------------------------------------------------------------------------------------------------
# Ranking files and the MS MARCO evaluation (src/utils/colbert_ir.py:261-345)
# ------------------------------------------------------------------------------------------------------------------
MSMARCO_DEPTHS = (5, 10, 20, 50, 100, 200, 500, 1000)


def save_ranking_tsv(path: str, qids, ids: torch.Tensor, scores: torch.Tensor | None = None) -> int:
    """Write ranked lists as the ``qid \\t pid \\t rank [\\t score]`` lines that colbert-ai's ``Ranking.save`` produces and
    ``evaluate`` reads back (colbert_ir.py:283-293).  ``ids`` [Q, k] (-1 = padding), ranks start at 1.  -> lines written."""
    ids_l = ids.cpu().tolist()
    sc_l = scores.cpu().tolist() if scores is not None else None
    n = 0
    with open(path, "w") as f:
        for qi, qid in enumerate(qids):
            for r, pid in enumerate(ids_l[qi]):
                if pid < 0:
                    break
                f.write(f"{qid}\t{pid}\t{r + 1}\t{sc_l[qi][r]}\n" if sc_l is not None else f"{qid}\t{pid}\t{r + 1}\n")
                n += 1
    return n


def load_ranking_tsv(path: str):
    """-> (qids list, ids int32 [Q, k] padded with -1, scores float64 [Q, k] or None), rows ordered by rank."""
    by_q: dict = {}
    with open(path) as f:
        for line in f:
            qid, pid, rank, *score = line.strip().split("\t")                      # colbert_ir.py:285
            by_q.setdefault(int(qid), []).append((int(rank), int(pid), float(score[0]) if score else None))
    qids = list(by_q.keys())
    k = max((len(v) for v in by_q.values()), default=0)
    ids = torch.full((len(qids), k), -1, dtype=torch.int32)
    has_scores = all(s is not None for v in by_q.values() for _, _, s in v)
    scores = torch.full((len(qids), k), float("-inf"), dtype=torch.float64) if has_scores else None
    for qi, qid in enumerate(qids):
        for j, (_, pid, s) in enumerate(by_q[qid]):
            ids[qi, j] = pid
            if has_scores:
                scores[qi, j] = s
    return qids, ids, scores


def evaluate_ranking_tensors(ids: torch.Tensor, gold_ptr: torch.Tensor, gold_ids: torch.Tensor,
                             lens: torch.Tensor | None = None, depths=MSMARCO_DEPTHS) -> dict:
    """The numbers of ``evaluate`` (colbert_ir.py:303-340) - recall@depth, MRR@10 and R-precision, averaged over the
    ranked queries - from device tensors through ``fz_rank_metrics`` (no per-query Python loops)."""
    from .. import ops
    vals = ops.rank_metrics(ids, lens, gold_ptr, gold_ids, tuple(depths), (), (10,), ()).cpu().tolist()
    names = ops.metric_names(tuple(depths), (), (10,), ())
    out = dict(zip(names, vals))
    out["rp"] = out.pop("r-precision")
    return out
