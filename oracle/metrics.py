"""Oracle (TEST INFRASTRUCTURE): restatement of the reference's retrieval metrics and of the linear-fusion weight sweep.

Follows ``src/utils/metrics.py``:
  * recall@k, precision@k                      Metrics.recall / .precision   :137-162
  * average precision@k (sum of P@i at hits / |gold|)      .average_precision :73-84
  * reciprocal rank@k                          .reciprocal_rank :86-96
  * nDCG@k with the reference's discount (rel_0 + sum_i>=1 rel_i / log2(i+1), idcg over ALL gold docs)   .ndcg :98-111
  * R-precision                                .r_precision :113-124
  * mean over queries with statistics.mean     .compute_mean_score :60-71
and the sweep of ``src/retrievers/hybrid.py:404-426`` (fuse with every weight vector of the grid, evaluate with
``run_evaluation``'s metric set, :24-27).  Pinned against the verbatim classes by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import itertools
from statistics import mean

import numpy as np

from . import fusion as ofusion

RECALL_KS = (5, 10, 20, 50, 100, 200, 500, 1000)      # hybrid.py:27
MAP_KS = MRR_KS = NDCG_KS = (10, 100)


def metric_names(recall_ks=RECALL_KS, map_ks=MAP_KS, mrr_ks=MRR_KS, ndcg_ks=NDCG_KS):
    return ([f"recall@{k}" for k in recall_ks] + [f"map@{k}" for k in map_ks] + [f"mrr@{k}" for k in mrr_ks] +
            [f"ndcg@{k}" for k in ndcg_ks] + ["r-precision"])


def query_metrics(gold: list[int], results: list[int], recall_ks=RECALL_KS, map_ks=MAP_KS, mrr_ks=MRR_KS,
                  ndcg_ks=NDCG_KS) -> list[float]:
    g = list(gold)
    hit = [1 if d in g else 0 for d in results]
    out = []
    for k in recall_ks:
        out.append(sum(hit[:k]) / len(g))
    for k in map_ks:
        p = [(sum(hit[:i + 1]) / (i + 1)) if hit[i] else 0 for i in range(len(results[:k]))]
        out.append(sum(p) / len(g))
    for k in mrr_ks:
        out.append(max([1 / (i + 1) if hit[i] else 0.0 for i in range(len(results[:k]))]))
    for k in ndcg_ks:
        rel = hit[:k]
        dcg = rel[0] + sum(rel[i] / np.log2(i + 1) for i in range(1, len(rel)))
        idcg = 1 + sum(1 / np.log2(i + 1) for i in range(1, len(g)))
        out.append((dcg / idcg) if idcg != 0 else 0)
    R = len(g)
    out.append(sum(hit[:R]) / R)
    return [float(x) for x in out]


def mean_metrics(golds, results, **ks) -> list[float]:
    per_q = [query_metrics(g, r, **ks) for g, r in zip(golds, results)]
    return [mean(col) for col in zip(*per_q)]


def weight_grid(n_sys: int, step: float = 0.05) -> list[tuple[float, ...]]:
    """hybrid.py:405-409: every combination of np.arange(0, 1+step, step) per system that sums to 1 (np.isclose)."""
    return [tuple(float(x) for x in comb) for comb in itertools.product(np.arange(0, 1 + step, step), repeat=n_sys)
            if np.isclose(sum(comb), 1.0)]


def sweep(ids_per_sys, scores_per_sys, golds, weights_list, normalization, distrs=None, **ks):
    """ids_per_sys[s][q], scores_per_sys[s][q]: one ranked list per system and query -> [W][M] mean metrics."""
    n_q = len(golds)
    out = []
    for w in weights_list:
        res = []
        for qi in range(n_q):
            ids, _ = ofusion.fuse_query([x[qi] for x in ids_per_sys], [x[qi] for x in scores_per_sys], "nsf",
                                        normalization, list(w), distrs)
            res.append(ids)
        out.append(mean_metrics(golds, res, **ks))
    return out
