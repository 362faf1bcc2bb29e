"""Oracle (TEST INFRASTRUCTURE): array restatement of the reference rank fusion.

Follows ``src/retrievers/hybrid.py``:
  * fuse loop (per query, per system, dict order)          Aggregator.fuse :170-220
  * duplicate ids keep first position, last score          Aggregator.convert2dict :223-233
  * borda ``(n-idx+1)/n`` / reciprocal rank ``1/(60+idx+1)`` (Python fp64)        :247-252
  * min-max / z-score (unbiased std) / arctan / percentile-rank / NCE in torch fp32   :254-278
  * weight ``score * w`` (np.float32 * python float stays fp32 under numpy>=2)        :283-291
  * union-sum in system order, stable descending sort => ties by first insertion      :294-307
The reference's ``final_results[:return_topk]`` slices the list of QUERIES (hybrid.py:220);
that quirk lives in the list-of-dict adapter, not here.

Pinned against the verbatim ``Aggregator`` by ``oracle/make_golden.py`` / ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

METHODS = ("bcf", "rrf", "nsf")
NORMALIZATIONS = ("none", "min-max", "z-score", "arctan", "percentile-rank", "normal-curve-equivalent")


def _dedup(ids: np.ndarray, scores: np.ndarray):
    """dict semantics: a repeated id keeps its first position and its last score."""
    first: dict[int, int] = {}
    out_ids, out_sc = [], []
    for i, s in zip(ids.tolist(), scores.tolist()):
        p = first.get(i)
        if p is None:
            first[i] = len(out_ids)
            out_ids.append(i)
            out_sc.append(s)
        else:
            out_sc[p] = s
    return out_ids, out_sc


def transform(scores: list[float], transformation: str | None, distr=None):
    """-> list of per-rank values with the reference's dtypes (np.float32 for torch paths, float otherwise)."""
    n = len(scores)
    if transformation == "borda-count":
        return [(n - idx + 1) / n for idx in range(n)]
    if transformation == "reciprocal-rank":
        return [1 / (60 + idx + 1) for idx in range(n)]
    if transformation in ("min-max", "z-score", "arctan", "percentile-rank", "normal-curve-equivalent"):
        t = torch.tensor(scores, dtype=torch.float32)
        if transformation == "min-max":
            lo, hi = torch.min(t), torch.max(t)
            t = (t - lo) / (hi - lo) if lo != hi else torch.ones_like(t)
        elif transformation == "z-score":
            m, sd = torch.mean(t), torch.std(t)
            t = (t - m) / sd if sd != 0 else torch.zeros_like(t)
        elif transformation == "arctan":
            t = (2 / math.pi) * torch.atan(0.1 * t)
        else:
            d = torch.tensor(np.asarray(distr), dtype=torch.float32)
            # argmin_i |d_i - s| (first index on ties) / len(d); evaluated in row blocks to bound memory
            idx = torch.empty(n, dtype=torch.int64)
            for lo in range(0, n, 256):
                idx[lo:lo + 256] = torch.argmin(torch.abs(d[:, None] - t[lo:lo + 256]), dim=0)
            t = idx / d.size(0)
            if transformation == "normal-curve-equivalent":
                t = torch.distributions.Normal(0, 1).icdf(t / 100) * 21.06 + 50
        return list(t.numpy())
    return list(scores)


def fuse_query(ids_per_sys, scores_per_sys, method: str, normalization: str | None = None,
               weights=None, distrs=None, promote_f64: bool = False):
    """One query.  ids_per_sys[s]: int array, scores_per_sys[s]: float64 array (rank order).
    -> (ids list, scores list) of the union, fused-score descending, ties by first insertion."""
    assert method in METHODS
    agg: dict[int, float] = {}
    for s, (ids, sc) in enumerate(zip(ids_per_sys, scores_per_sys)):
        ids, sc = _dedup(np.asarray(ids), np.asarray(sc, dtype=np.float64))
        if method == "bcf":
            vals = transform(sc, "borda-count")
        elif method == "rrf":
            vals = transform(sc, "reciprocal-rank")
        else:
            vals = transform(sc, normalization, None if distrs is None else distrs[s])
            # python float: np.float32 * float stays fp32 under NumPy >= 2 (NEP 50); NumPy 1.x (what the reference pins)
            # promotes to float64 - identical to multiplying by an np.float64 scalar
            w = np.float64(weights[s]) if promote_f64 else float(weights[s])
            vals = [v * w for v in vals]
        for i, v in zip(ids, vals):
            agg[i] = agg.get(i, 0.0) + v
    items = sorted(agg.items(), key=lambda x: x[1], reverse=True)
    return [i for i, _ in items], [v for _, v in items]
