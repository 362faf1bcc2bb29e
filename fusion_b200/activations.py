"""SPLADE activation head on the device: MLM logits -> pooled term weights -> pruning -> CSR.

Mirrors ``SPLADE.forward`` after the encoder call (src/retrievers/splade/splade.py:88-99) and
``SPLADE._prune_activations`` (:295-306).  The transformer that produces the logits stays stock PyTorch; everything
after it runs in the kernels of ``csrc/activations.cu``.  The reference keeps the [rows, vocab] activations dense all
the way into ``torch.mm``; here they become the CSR rows the sparse index (documents) and the scorer (queries) consume.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import FusionB200Error, check
from .ops import _ptr, _stream

POOLINGS = {"max": 0, "sum": 1}       # FZ_POOL_MAX / FZ_POOL_SUM
MAX_ROWS_PER_CALL = 65535


def splade_pool(logits: torch.Tensor, input_masks: torch.Tensor, pooling: str = "max") -> torch.Tensor:
    """logits [B, L, V] (fp32 or bf16, CUDA) and attention mask [B, L] -> activations fp32 [B, V]."""
    if pooling not in POOLINGS:
        raise AssertionError("The sparse vector aggregation strategy should either be 'max' or 'sum'.")   # splade.py:74
    if not logits.is_cuda:
        raise FusionB200Error("logits must be a CUDA tensor (fusion_b200 has no CPU path)")
    if logits.dim() != 3 or input_masks.shape != logits.shape[:2]:
        raise FusionB200Error(f"expected logits [B, L, V] and masks [B, L], got {tuple(logits.shape)} / {tuple(input_masks.shape)}")
    if logits.dtype not in (torch.float32, torch.bfloat16):
        logits = logits.float()
    logits = logits.contiguous()
    mask = input_masks.to(device=logits.device, dtype=torch.int32).contiguous()
    b, l, v = logits.shape
    out = torch.empty((b, v), dtype=torch.float32, device=logits.device)
    lib = _lib.load()
    for r0 in range(0, b, MAX_ROWS_PER_CALL):
        r1 = min(b, r0 + MAX_ROWS_PER_CALL)
        check(lib.fz_splade_pool(_ptr(logits[r0:r1]), int(logits.dtype == torch.bfloat16), _ptr(mask[r0:r1]), r1 - r0, l, v,
                                 POOLINGS[pooling], _ptr(out[r0:r1]), _stream(out)), "fz_splade_pool")
    return out


def prune_activations(activations: torch.Tensor, keep_topk: int, want_indices: bool = True):
    """``SPLADE._prune_activations``: -> (pruned [B, V], topk_indices int64 [B, keep_topk] by value, best first).
    Ties at the cutoff keep the lower term id (torch.topk leaves that choice unspecified)."""
    act = ops._req(activations, torch.float32, "activations")
    b, v = act.shape
    keep_topk = int(keep_topk)
    out = torch.empty_like(act)
    check(_lib.load().fz_prune_topk(_ptr(act), b, v, keep_topk, _ptr(out), _stream(out)), "fz_prune_topk")
    idx = None
    if want_indices:
        _, idx = ops.rank_rows(act, keep_topk, 0)
        idx = idx.long()
    return out, idx


def activations_to_csr(activations: torch.Tensor):
    """Dense [B, V] fp32 -> CSR (ptr int64 [B + 1], term int32 ascending per row, weight fp32), zeros dropped."""
    act = ops._req(activations, torch.float32, "activations")
    b, v = act.shape
    lib = _lib.load()
    nnz = torch.empty(b, dtype=torch.int32, device=act.device)
    check(lib.fz_csr_count(_ptr(act), b, v, _ptr(nnz), _stream(act)), "fz_csr_count")
    ptr = torch.zeros(b + 1, dtype=torch.int64, device=act.device)
    ptr[1:] = torch.cumsum(nnz, 0)
    total = int(ptr[-1])
    term = torch.empty(total, dtype=torch.int32, device=act.device)
    weight = torch.empty(total, dtype=torch.float32, device=act.device)
    if total:
        check(lib.fz_csr_fill(_ptr(act), b, v, _ptr(ptr), _ptr(term), _ptr(weight), _stream(act)), "fz_csr_fill")
    return ptr, term, weight


def splade_encode_csr(logits: torch.Tensor, input_masks: torch.Tensor, pooling: str = "max", pruning_topk: int | None = None):
    """logits -> pooled -> (pruned) -> CSR in one call; what an index build or a query batch feeds to K2."""
    act = splade_pool(logits, input_masks, pooling)
    if pruning_topk is not None:
        act, _ = prune_activations(act, pruning_topk, want_indices=False)
    return activations_to_csr(act)
