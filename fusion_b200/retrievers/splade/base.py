"""Drop-in mirror of the scoring part of ``src/retrievers/splade/base.py``:
``compute_pairwise_similarity`` (:173-184), ``compute_batchwise_similarity`` (:186-197) and ``search`` (:199-251).

The reference's ``BaseModel`` is an abstract ``torch.nn.Module`` that also tokenises, encodes, trains and saves;
those parts stay stock PyTorch.  This mixin supplies the three scoring methods with the reference's signatures;
a model class lists it before the reference base (or any class providing ``encode`` and ``similarity``).
"""
from __future__ import annotations

import torch

from ... import ops
from ..hybrid import Ranker


class BaseModel:
    similarity: str = "cos_sim"

    def compute_pairwise_similarity(self, q_embs: torch.Tensor, d_embs: torch.Tensor) -> torch.Tensor:
        """[B, d] x [B, d] -> [B] (base.py:173-184): row-wise products, on the GPU library."""
        q32, _ = ops.normalize_rows(q_embs.float().cuda(), normalize=self.similarity == "cos_sim", want_bf16=False)
        d32, _ = ops.normalize_rows(d_embs.float().cuda(), normalize=self.similarity == "cos_sim", want_bf16=False)
        return ops.pairwise_dot(q32, d32)

    def compute_batchwise_similarity(self, q_embs: torch.Tensor, d_embs: torch.Tensor) -> torch.Tensor:
        """[Q, d] x [D, d] -> [Q, D] fp32 (base.py:186-197): optional L2 normalisation, exact fp32 products."""
        q32, _ = ops.normalize_rows(q_embs.float().cuda(), normalize=self.similarity == "cos_sim", want_bf16=False)
        d32, _ = ops.normalize_rows(d_embs.float().cuda(), normalize=self.similarity == "cos_sim", want_bf16=False)
        return ops.dense_scores(q32, d32)

    def search(self, queries: list[str], documents: list[str], batch_size: int = 32, query_chunk_size: int = 100,
               doc_chunk_size: int = 500000, topk: int = 10) -> list[dict[str, float]]:
        """Similarity search between queries and documents (base.py:199-251).

        ``query_chunk_size`` / ``doc_chunk_size`` are accepted for signature compatibility; the device kernels tile
        the corpus themselves and merge on the GPU instead of through Python heaps.  Returns, per query, the top-k
        ``{'doc_id', 'score'}`` sorted by score descending."""
        query_embeddings = self.encode(queries, query_mode=True, batch_size=batch_size)
        doc_embeddings = self.encode(documents, query_mode=False, batch_size=batch_size)
        scores, ids = self.search_tensors(query_embeddings, doc_embeddings, topk)
        return [[{"doc_id": i, "score": s} for i, s in zip(ri, rs)]
                for ri, rs in zip(ids.cpu().tolist(), scores.cpu().tolist())]

    def search_tensors(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, topk: int):
        q, d = query_embeddings.cuda().float(), doc_embeddings.cuda().float()
        sim = "cos_sim" if self.similarity == "cos_sim" else "dot"
        # SPLADE activations are [*, |V|] with a few hundred non-zeros: score them through the inverted index
        if d.shape[1] >= 4096 and float((d[: min(len(d), 256)] != 0).float().mean()) < 0.1:
            return Ranker.sparse_vector_search_tensors(q, d, topk, sim)
        return Ranker.dense_search_tensors(q, d, topk, sim)
