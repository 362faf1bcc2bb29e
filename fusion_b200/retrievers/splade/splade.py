"""Drop-in mirror of the activation head of ``src/retrievers/splade/splade.py``: ``SPLADE.forward`` (:80-99) after the
encoder call and ``SPLADE._prune_activations`` (:295-306).

The masked-language-model transformer stays stock PyTorch (``self.model``: anything returning an object with
``.logits`` of shape [batch, seq, vocab]); pooling, pruning and the conversion to the CSR rows the sparse index consumes
run in the CUDA library (``csrc/activations.cu``).  Signatures, argument meaning and the assertion message follow the
reference class; ``encode_csr`` is the additional index-build entry (the reference keeps the vectors dense).
"""
from __future__ import annotations

import torch

from ... import activations as act_ops
from .base import BaseModel

__all__ = ["SPLADE"]


class SPLADE(BaseModel, torch.nn.Module):
    def __init__(self, model: torch.nn.Module, pooling: str = "max", pruning_topk: int = None, similarity: str = "cos_sim",
                 tokenizer=None, max_query_length: int = 64, max_doc_length: int = 512):
        torch.nn.Module.__init__(self)
        assert pooling in ["max", "sum"], "The sparse vector aggregation strategy should either be 'max' or 'sum'."
        self.model = model
        self.pooling = pooling
        self.pruning_topk = pruning_topk
        self.similarity = similarity
        self.tokenizer = tokenizer
        self.max_query_length, self.max_doc_length = max_query_length, max_doc_length

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, max_query_length: int = 64, max_doc_length: int = 512, **kw):
        """Stock Hugging Face masked-LM + tokenizer (the encoder is not part of the hot path) under the CUDA activation head."""
        from transformers import AutoModelForMaskedLM, AutoTokenizer
        model = AutoModelForMaskedLM.from_pretrained(model_name_or_path)
        tok = AutoTokenizer.from_pretrained(model_name_or_path)
        return cls(model.cuda().eval(), tokenizer=tok, max_query_length=max_query_length, max_doc_length=max_doc_length, **kw)

    @torch.no_grad()
    def encode(self, sentences: list[str], query_mode: bool = False, batch_size: int = 32, convert_to_tensor: bool = True,
               show_progress_bar: bool = False, **_):
        """Texts -> activations [n, vocab] fp32 on the device (the reference's ``encode``, splade/base.py:254-291)."""
        if self.tokenizer is None:
            raise RuntimeError("SPLADE.encode needs a tokenizer (SPLADE.from_pretrained or tokenizer=...)")
        dev = next(self.model.parameters()).device
        max_len = self.max_query_length if query_mode else self.max_doc_length
        out = []
        for lo in range(0, len(sentences), batch_size):
            enc = self.tokenizer(sentences[lo:lo + batch_size], padding=True, truncation=True, max_length=max_len, return_tensors="pt")
            out.append(self.forward(enc["input_ids"].to(dev), enc["attention_mask"].to(dev)))
        acts = torch.cat(out, 0) if out else torch.zeros((0, 0), device=dev)
        return acts if convert_to_tensor else acts.cpu().numpy()

    def forward(self, input_ids: torch.Tensor, input_masks: torch.Tensor) -> torch.Tensor:
        """[batch, seq] ids and masks -> activations [batch, vocab] fp32 (splade.py:80-99)."""
        out = self.model(input_ids=input_ids, attention_mask=input_masks)
        activations = act_ops.splade_pool(out.logits, input_masks, self.pooling)
        if self.pruning_topk is not None:
            activations, _ = self._prune_activations(activations, keep_topk=self.pruning_topk)
        return activations

    def _prune_activations(self, activations: torch.Tensor, keep_topk: int):
        """-> (pruned activations [batch, vocab], top-k indices [batch, keep_topk] by value) (splade.py:295-306)."""
        return act_ops.prune_activations(activations, int(keep_topk))

    def encode_csr(self, input_ids: torch.Tensor, input_masks: torch.Tensor):
        """One batch -> CSR (ptr, term ids ascending, weights): the rows of a ``SparseIndex`` or of a query batch."""
        out = self.model(input_ids=input_ids, attention_mask=input_masks)
        return act_ops.splade_encode_csr(out.logits, input_masks, self.pooling, self.pruning_topk)
