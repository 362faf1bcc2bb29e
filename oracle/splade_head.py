"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the SPLADE activation head.

  * pooling      ``src/retrievers/splade/splade.py:88-94``: ``torch.sum`` / ``torch.amax`` over the sequence of
                 ``log1p(relu(logits * attention_mask.unsqueeze(-1)))``
  * pruning      ``SPLADE._prune_activations`` (``splade.py:295-306``): ``torch.topk(sorted=True)`` then
                 ``zeros_like(...).scatter(1, indices, values)``
  * dense -> CSR not in the reference (it keeps [rows, vocab] dense); the obvious restatement: non-zeros of every row,
                 term ids ascending.

Pinned by ``tests/golden/splade_head_small.npz``: outputs of the verbatim ``SPLADE.forward`` /
``_prune_activations`` run with a stub encoder (``oracle/make_golden.py::golden_splade_head``).
Which of several equal activations survives a pruning cutoff is unspecified in torch.topk: compare kept VALUES
(sorted) and the kept set outside the tie group.
"""
from __future__ import annotations

import numpy as np
import torch


def pool(logits: torch.Tensor, input_masks: torch.Tensor, pooling: str = "max") -> torch.Tensor:
    assert pooling in ["max", "sum"]
    attention_mask = input_masks.unsqueeze(-1)
    x = torch.log1p(torch.relu(logits * attention_mask))
    return torch.sum(x, dim=1) if pooling == "sum" else torch.amax(x, dim=1)


def prune(activations: torch.Tensor, keep_topk: int):
    vals, idx = torch.topk(activations, k=int(keep_topk), dim=1, largest=True, sorted=True)
    return torch.zeros_like(activations).scatter(dim=1, index=idx, src=vals), idx


def to_csr(activations: torch.Tensor):
    a = activations.numpy()
    rows, cols = np.nonzero(a)
    ptr = np.zeros(a.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=a.shape[0]), out=ptr[1:])
    return ptr, cols.astype(np.int32), a[rows, cols].astype(np.float32)
