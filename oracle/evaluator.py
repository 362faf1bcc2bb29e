"""Oracle (TEST INFRASTRUCTURE): restatement of ``InformationRetrievalEvaluatorCustom.compute_metrices`` /
``compute_metrics`` (``src/utils/sentence_transformers.py:314-393`` and ``:395-485``).

The reference class derives from sentence-transformers' ``InformationRetrievalEvaluator`` (sentence-transformers==2.2.2,
third-party, not installed here), so the module cannot be imported; the two methods are restated loop for loop:
corpus chunks of ``corpus_chunk_size``, one score call per query, ``torch.topk(sorted=False)``, ``heapq`` merge keeping
``max_k`` (score, corpus_id) pairs (:334-364), then the metric loops of :395-485 with the result lists sorted by score
(:410).  Only ``tests/`` may import this.
"""
from __future__ import annotations

import heapq

import numpy as np
import torch


def _score(a, b, name):
    if "cos" in name:
        a = torch.nn.functional.normalize(a, p=2, dim=1)
        b = torch.nn.functional.normalize(b, p=2, dim=1)
    return torch.mm(a, b.t())


def search(query_embeddings, corpus_embeddings, corpus_ids, max_k, name, corpus_chunk_size=50000):
    """:334-364 - per corpus chunk, per query: score, topk (unsorted), heap merge."""
    res = [[] for _ in range(len(query_embeddings))]
    for start in range(0, len(corpus_embeddings), corpus_chunk_size):
        sub = corpus_embeddings[start:start + corpus_chunk_size]
        for qi in range(len(query_embeddings)):
            scores = _score(query_embeddings[qi:qi + 1], sub, name)
            vals, idx = torch.topk(scores, min(max_k, scores.shape[1]), dim=1, largest=True, sorted=False)
            for sub_id, sc in zip(idx[0].tolist(), vals[0].tolist()):
                cid = corpus_ids[start + sub_id]
                if len(res[qi]) < max_k:
                    heapq.heappush(res[qi], (sc, cid))
                else:
                    heapq.heappushpop(res[qi], (sc, cid))
    return [[{"corpus_id": c, "score": s} for s, c in h] for h in res]


def dcg(relevances, k):
    return sum(r / np.log2(i + 2) for i, r in enumerate(relevances[:k]))


def compute_metrics(queries_result_list, queries_ids, relevant_docs, n_queries, mrr_at_k, ndcg_at_k, accuracy_at_k,
                    precision_recall_at_k, map_at_k):
    """:395-485, loop for loop."""
    num_hits = {k: 0 for k in accuracy_at_k}
    precision = {k: [] for k in precision_recall_at_k}
    recall = {k: [] for k in precision_recall_at_k}
    mrr = {k: 0 for k in mrr_at_k}
    ndcg = {k: [] for k in ndcg_at_k}
    avgp = {k: [] for k in map_at_k}
    rp = []
    for qi in range(len(queries_result_list)):
        top_hits = sorted(queries_result_list[qi], key=lambda x: x["score"], reverse=True)
        rel = relevant_docs[queries_ids[qi]]
        n_rel = len(rel)
        for k in accuracy_at_k:
            for hit in top_hits[:k]:
                if hit["corpus_id"] in rel:
                    num_hits[k] += 1
                    break
        for k in precision_recall_at_k:
            c = sum(1 for hit in top_hits[:k] if hit["corpus_id"] in rel)
            precision[k].append(c / k)
            recall[k].append(c / n_rel)
        for k in mrr_at_k:
            for rank, hit in enumerate(top_hits[:k]):
                if hit["corpus_id"] in rel:
                    mrr[k] += 1.0 / (rank + 1)
                    break
        for k in ndcg_at_k:
            pred = [1 if h["corpus_id"] in rel else 0 for h in top_hits[:k]]
            ndcg[k].append(dcg(pred, k) / dcg([1] * n_rel, k))
        for k in map_at_k:
            c, sp = 0, 0
            for rank, hit in enumerate(top_hits[:k]):
                if hit["corpus_id"] in rel:
                    c += 1
                    sp += c / (rank + 1)
            avgp[k].append(sp / min(k, n_rel))
        rp.append(sum(1 for hit in top_hits[:n_rel] if hit["corpus_id"] in rel) / n_rel)
    return {"accuracy@k": {k: v / n_queries for k, v in num_hits.items()},
            "precision@k": {k: np.mean(v) for k, v in precision.items()},
            "recall@k": {k: np.mean(v) for k, v in recall.items()},
            "ndcg@k": {k: np.mean(v) for k, v in ndcg.items()},
            "mrr@k": {k: v / n_queries for k, v in mrr.items()},
            "map@k": {k: np.mean(v) for k, v in avgp.items()},
            "r-precision": np.mean(rp)}
