"""Multi-GPU partitioning: the corpus is cut into contiguous doc-id ranges, one per rank; queries are replicated.

Every rank scores its shard and keeps a local top-k; ONE collective per retriever moves only the k candidates per
query over NVLink (NCCL through torch.distributed), then the k-way merge kernel (K5) produces the global list:

  * ``gather_merge_topk``    all-gather of [Q, k] (score, id) pairs, every rank merges every query
  * ``exchange_merge_topk``  all-to-all so that rank g receives and merges only its query slice (G x less receive
                             traffic, and the per-query tail - ColBERT rescoring, fusion - is then query-sharded)

BM25 statistics (N, df, sum of doc lengths) are corpus-global: ``allreduce_lexical_stats`` sums them once at index
time so that sharded scores are bit-identical to unsharded ones (bm25.py:133-147 computes them over the whole corpus).
The reference has no equivalent (single process); this is the north-star's only exchange step.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_docs: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous doc range [lo, hi) of ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_docs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def query_slice(n_queries: int, world: int, rank: int) -> tuple[int, int]:
    """Query range owned by ``rank`` for the query-sharded tail; every slice has ceil(Q / world) slots."""
    per = (n_queries + world - 1) // world
    return min(n_queries, rank * per), min(n_queries, (rank + 1) * per)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_lexical_stats(n_local: int, df_local: np.ndarray, sum_dl_local: int, device, group=None):
    """-> (N, df[V], sum_dl) over all shards."""
    world, _ = _world(group)
    if world == 1:
        return int(n_local), np.asarray(df_local), int(sum_dl_local)
    t = torch.cat([torch.tensor([n_local, sum_dl_local], dtype=torch.int64), torch.as_tensor(df_local, dtype=torch.int64)]).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu()
    return int(t[0]), t[2:].numpy(), int(t[1])


def allreduce_min(t: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise minimum over the ranks (in place): the cross-shard score floors (``ops.ShardSync``, and the rescoring
    floor of the exact dense mode)."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def allreduce_max_ints(values: list[int], device, group=None) -> list[int]:
    """Element-wise maximum of a few host integers over the ranks (the largest shard size per index: every shard
    follows the round schedule of the largest one, so all ranks issue the same collectives)."""
    world, _ = _world(group)
    if world == 1:
        return [int(v) for v in values]
    t = torch.tensor(values, dtype=torch.int64).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [int(v) for v in t.cpu().tolist()]


def gather_topk(scores: torch.Tensor, ids: torch.Tensor, group=None):
    """All-gather local top-k lists: [Q, k] -> [G, Q, k] (one collective per tensor)."""
    world, _ = _world(group)
    if world == 1:
        return scores[None], ids[None]
    q, k = scores.shape
    gs = torch.empty((world * q, k), dtype=scores.dtype, device=scores.device)     # concatenated form: every backend takes it
    gi = torch.empty((world * q, k), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    return gs.view(world, q, k), gi.view(world, q, k)


def exchange_topk(scores: torch.Tensor, ids: torch.Tensor, group=None):
    """All-to-all by query slice: [Q, k] local lists -> [G, Qs, k] lists of THIS rank's query slice
    (Qs = ceil(Q / G); slots past Q are padded with (-inf, -1))."""
    world, rank = _world(group)
    if world == 1:
        return scores[None], ids[None]
    q, k = scores.shape
    per = (q + world - 1) // world
    pad = per * world - q
    if pad:
        scores = torch.cat([scores, torch.full((pad, k), float("-inf"), dtype=scores.dtype, device=scores.device)])
        ids = torch.cat([ids, torch.full((pad, k), -1, dtype=ids.dtype, device=ids.device)])
    rs, ri = torch.empty_like(scores), torch.empty_like(ids)
    dist.all_to_all_single(rs, scores.contiguous(), group=group)
    dist.all_to_all_single(ri, ids.contiguous(), group=group)
    return rs.view(world, per, k), ri.view(world, per, k)


def gather_merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int, group=None, merge=None):
    """Global top-k on every rank.  ``merge`` defaults to the CUDA k-way merge (fz_merge_topk)."""
    if merge is None:
        from . import ops
        merge = ops.merge_topk
    gs, gi = gather_topk(scores, ids, group)
    return merge(gs, gi, k)


def exchange_merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int, group=None, merge=None):
    """Global top-k of this rank's query slice ([Qs, k]); see ``query_slice`` for the slice bounds."""
    if merge is None:
        from . import ops
        merge = ops.merge_topk
    gs, gi = exchange_topk(scores, ids, group)
    return merge(gs, gi, k)


def reduce_scatter_scores(partial: torch.Tensor, group=None):
    """Sum the per-shard partial score matrices [Q, C] (each (q, c) is non-zero on exactly one shard) and return
    this rank's query slice [Qs, C]."""
    world, rank = _world(group)
    if world == 1:
        return partial
    q, c = partial.shape
    per = (q + world - 1) // world
    pad = per * world - q
    if pad:
        partial = torch.cat([partial, torch.zeros((pad, c), dtype=partial.dtype, device=partial.device)])
    out = torch.empty((per, c), dtype=partial.dtype, device=partial.device)
    dist.reduce_scatter_tensor(out, partial.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


def allgather_rows(x: torch.Tensor, group=None):
    """Concatenate per-rank row blocks of equal shape: [Qs, ...] -> [G * Qs, ...]."""
    world, _ = _world(group)
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out
