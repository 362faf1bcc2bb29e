"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the reference dense score + top-k loops.

  * cos_sim / dot_score      ``src/retrievers/splade/base.py:186-197`` (F.normalize both sides, torch.mm)
  * chunked search + heap    ``src/retrievers/splade/base.py:199-251`` (query chunks of 100, doc chunks of
                             500,000, ``torch.topk(sorted=False)``, ``heapq`` of (score, id), final sort desc)
  * evaluator variant        ``src/utils/sentence_transformers.py:314-393,410`` (corpus chunks of 50,000,
                             one GEMV + topk per query, heap merge, sort at :410)
  * ``sentence_transformers.util.semantic_search`` (sentence-transformers==2.2.2, third-party, absent;
    call site ``src/retrievers/hybrid.py:103``): same algorithm with key ``corpus_id``.

The reference's order inside a group of exactly tied scores is arbitrary (unsorted topk + heap array
order): callers compare tie groups as sets.
"""
from __future__ import annotations

import heapq

import torch


def similarity(q: torch.Tensor, d: torch.Tensor, sim: str = "cos_sim") -> torch.Tensor:
    if sim == "cos_sim":
        q = torch.nn.functional.normalize(q, p=2, dim=-1)
        d = torch.nn.functional.normalize(d, p=2, dim=-1)
    return torch.mm(q, d.t())


def semantic_search(q_embs: torch.Tensor, d_embs: torch.Tensor, top_k: int, sim: str = "cos_sim",
                    query_chunk_size: int = 100, corpus_chunk_size: int = 500000, key: str = "corpus_id"):
    res = [[] for _ in range(len(q_embs))]
    for qs in range(0, len(q_embs), query_chunk_size):
        for ds in range(0, len(d_embs), corpus_chunk_size):
            scores = similarity(q_embs[qs:qs + query_chunk_size], d_embs[ds:ds + corpus_chunk_size], sim)
            vals, idx = torch.topk(scores, min(top_k, scores.shape[1]), dim=1, largest=True, sorted=False)
            vals, idx = vals.tolist(), idx.tolist()
            for qi in range(len(vals)):
                h = res[qs + qi]
                for sub, sc in zip(idx[qi], vals[qi]):
                    if len(h) < top_k:
                        heapq.heappush(h, (sc, ds + sub))
                    else:
                        heapq.heappushpop(h, (sc, ds + sub))
    out = []
    for h in res:
        out.append(sorted(({key: i, "score": s} for s, i in h), key=lambda x: x["score"], reverse=True))
    return out


def topk_tensors(q_embs: torch.Tensor, d_embs: torch.Tensor, top_k: int, sim: str = "cos_sim",
                 chunk: int = 262144):
    """Same result as :func:`semantic_search` as (scores[Q,k] desc, ids[Q,k]); ties by lower id.
    Used for the larger parity cases where Python heaps are too slow."""
    qn = torch.nn.functional.normalize(q_embs, dim=-1) if sim == "cos_sim" else q_embs
    best_s = best_i = None
    for ds in range(0, len(d_embs), chunk):
        d = d_embs[ds:ds + chunk]
        dn = torch.nn.functional.normalize(d, dim=-1) if sim == "cos_sim" else d
        s = qn @ dn.t()
        i = torch.arange(ds, ds + len(d)).expand_as(s)
        if best_s is not None:
            s = torch.cat([best_s, s], 1)
            i = torch.cat([best_i, i], 1)
        k = min(top_k, s.shape[1])
        # stable descending sort => lower id first inside a tie group (ids are ascending in ``i``)
        order = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :k]
        best_s, best_i = torch.gather(s, 1, order), torch.gather(i, 1, order)
        o2 = torch.sort(best_i, dim=1, stable=True).indices      # restore id-ascending for the next merge
        best_s, best_i = torch.gather(best_s, 1, o2), torch.gather(best_i, 1, o2)
    order = torch.sort(best_s, dim=1, descending=True, stable=True).indices
    return torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
