"""Drop-in mirror of ``src/utils/metrics.py::Metrics`` (:25-162): same constructor, ``compute_all_metrics`` returns the
same ``{'recall@k', 'map@k', 'mrr@k', 'ndcg@k', 'r-precision'}`` means, computed by ``fz_rank_metrics`` on the device
from the ranked id lists (the reference loops over Python lists per query and cut-off).

The per-query helpers (``recall``, ``precision``, ``average_precision``, ...) are kept for API compatibility; they score
ONE query through the same kernel.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .. import ops


def _gold_csr(all_ground_truths, device):
    ptr = np.zeros(len(all_ground_truths) + 1, dtype=np.int32)
    np.cumsum([len(g) for g in all_ground_truths], out=ptr[1:])
    flat = np.fromiter((int(x) for g in all_ground_truths for x in g), dtype=np.int64, count=int(ptr[-1]))
    return torch.from_numpy(ptr).to(device), torch.from_numpy(flat.astype(np.int32)).to(device)


def _result_matrix(all_results, device):
    n = max(1, max((len(r) for r in all_results), default=1))
    ids = np.full((len(all_results), n), -1, dtype=np.int32)
    lens = np.zeros(len(all_results), dtype=np.int32)
    for i, r in enumerate(all_results):
        ids[i, :len(r)] = r
        lens[i] = len(r)
    return torch.from_numpy(ids).to(device), torch.from_numpy(lens).to(device)


class Metrics:
    """Evaluation metrics for retrieval tasks (utils/metrics.py:25-38)."""

    def __init__(self, recall_at_k: list[int], map_at_k: list[int] = [], mrr_at_k: list[int] = [], ndcg_at_k: list[int] = [],
                 device: str = "cuda"):
        self.recall_at_k = recall_at_k
        self.map_at_k = map_at_k
        self.mrr_at_k = mrr_at_k
        self.ndcg_at_k = ndcg_at_k
        self.device = device

    def compute_all_metrics(self, all_ground_truths: list[list[int]], all_results: list[list[int]]) -> dict:
        """Mean metrics over the queries (utils/metrics.py:40-58)."""
        if any(len(g) == 0 for g in all_ground_truths):
            raise ZeroDivisionError("division by zero")        # sum(...) / len(ground_truths), utils/metrics.py:84
        gp, gi = _gold_csr(all_ground_truths, self.device)
        ids, lens = _result_matrix(all_results, self.device)
        vals = ops.rank_metrics(ids, lens, gp, gi, tuple(self.recall_at_k), tuple(self.map_at_k), tuple(self.mrr_at_k),
                                tuple(self.ndcg_at_k)).cpu().tolist()
        names = ops.metric_names(self.recall_at_k, self.map_at_k, self.mrr_at_k, self.ndcg_at_k)
        scores = defaultdict(dict)
        for n, v in zip(names, vals):
            scores[n] = v
        return scores

    def compute_all_metrics_tensors(self, ids: torch.Tensor, lens: torch.Tensor | None, gold_ptr: torch.Tensor,
                                    gold_ids: torch.Tensor) -> dict:
        """Same, from device tensors ([Q, n] int32 ids, -1 padded) without the list round trip."""
        vals = ops.rank_metrics(ids, lens, gold_ptr, gold_ids, tuple(self.recall_at_k), tuple(self.map_at_k),
                                tuple(self.mrr_at_k), tuple(self.ndcg_at_k)).cpu().tolist()
        return dict(zip(ops.metric_names(self.recall_at_k, self.map_at_k, self.mrr_at_k, self.ndcg_at_k), vals))

    # -- per-query forms (utils/metrics.py:73-162): one query through the same kernel
    def _one(self, ground_truths, results, **ks):
        gp, gi = _gold_csr([ground_truths], self.device)
        ids, lens = _result_matrix([results], self.device)
        kw = dict(recall_ks=(), map_ks=(), mrr_ks=(), ndcg_ks=())
        kw.update(ks)
        return ops.rank_metrics(ids, lens, gp, gi, **kw).cpu().tolist()

    def recall(self, ground_truths, results, k=None):
        return self._one(ground_truths, results, recall_ks=(len(results) if k is None else k,))[0]

    def average_precision(self, ground_truths, results, k=None):
        return self._one(ground_truths, results, map_ks=(len(results) if k is None else k,))[0]

    def reciprocal_rank(self, ground_truths, results, k=None):
        return self._one(ground_truths, results, mrr_ks=(len(results) if k is None else k,))[0]

    def ndcg(self, ground_truths, results, k=None):
        return self._one(ground_truths, results, ndcg_ks=(len(results) if k is None else k,))[0]

    def r_precision(self, ground_truths, results, R=None):
        return self._one(ground_truths, results)[-1]

    def precision(self, ground_truths, results, k=None):
        k = len(results) if k is None else k
        return self.recall(ground_truths, results, k) * len(ground_truths) / len(results[:k])

    def fscore(self, ground_truths, results, k=None):
        p, r = self.precision(ground_truths, results, k), self.recall(ground_truths, results, k)
        return (2 * p * r) / (p + r) if (p != 0.0 or r != 0.0) else 0.0
