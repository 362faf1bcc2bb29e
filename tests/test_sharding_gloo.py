"""Multi-rank host logic on CPU: world_size-2 gloo group, corpus shards, the top-k exchange and global BM25 stats.
The device merge kernel is replaced by a numpy merge here; GPU tests cover fz_merge_topk itself."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fusion_b200 import sharding


def _np_merge(gs, gi, k):
    g, q, kin = gs.shape
    out_s = torch.full((q, k), float("-inf"), dtype=gs.dtype)
    out_i = torch.full((q, k), -1, dtype=gi.dtype)
    for qi in range(q):
        s, i = gs[:, qi].reshape(-1).numpy(), gi[:, qi].reshape(-1).numpy()
        keep = i >= 0
        s, i = s[keep], i[keep]
        order = np.lexsort((i, -s))[:k]
        out_s[qi, :len(order)] = torch.from_numpy(s[order])
        out_i[qi, :len(order)] = torch.from_numpy(i[order])
    return out_s, out_i


def _worker(rank, world, port, n_docs, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(3))
        scores = rng.normal(size=(nq, n_docs)).round(1)          # every rank sees the same "corpus scores"
        lo, hi = sharding.shard_bounds(n_docs, world, rank)
        loc = scores[:, lo:hi]
        order = np.stack([np.lexsort((np.arange(hi - lo), -loc[q]))[:k] for q in range(nq)])
        ls = torch.from_numpy(np.take_along_axis(loc, order, 1))
        li = torch.from_numpy((order + lo).astype(np.int32))
        # global stats
        n, df, sdl = sharding.allreduce_lexical_stats(hi - lo, np.arange(5) + rank, 100 * (rank + 1), "cpu")
        assert n == n_docs and df.tolist() == [sum(j + r for r in range(world)) for j in range(5)] and sdl == 100 * world * (world + 1) // 2
        # all-gather + merge: every rank holds the full result
        ms, mi = sharding.gather_merge_topk(ls, li, k, merge=_np_merge)
        ref = np.stack([np.lexsort((np.arange(n_docs), -scores[q]))[:k] for q in range(nq)])
        assert np.array_equal(mi.numpy(), ref)
        # all-to-all + merge: this rank holds its query slice
        xs, xi = sharding.exchange_merge_topk(ls, li, k, merge=_np_merge)
        qlo, qhi = sharding.query_slice(nq, world, rank)
        assert np.array_equal(xi.numpy()[: qhi - qlo], ref[qlo:qhi])
        assert (xi.numpy()[qhi - qlo:] == -1).all()
        # partial score matrices -> reduce-scatter by query slice, all-gather back
        part = torch.zeros(nq, 6)
        part[:, rank::world] = torch.arange(nq, dtype=torch.float32)[:, None] + rank
        mine = sharding.reduce_scatter_scores(part)
        full = sharding.allgather_rows(mine)[:nq]
        exp = torch.zeros(nq, 6)
        for r in range(world):
            exp[:, r::world] = torch.arange(nq, dtype=torch.float32)[:, None] + r
        assert torch.equal(full, exp)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_topk_exchange():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29641, 101, 7, 5, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_shard_bounds_cover_the_corpus():
    for n, w in ((8841823, 8), (10, 3), (5, 8)):
        b = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert sharding.query_slice(10, 4, 3) == (9, 10) and sharding.query_slice(10, 4, 0) == (0, 3)
