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


def _floor_protocol(rank, world, scores, k):
    """Host restatement of the threshold exchange of ``fz_shard_sync_t`` (rounds over the shard, local ceil(k/G)-th best,
    all-reduce MIN, keep everything at or above the floor, ties included): the merged shard lists are the global top-k,
    every rank issues the same number of collectives although the shards differ in size."""
    nq, n_docs = scores.shape
    lo, hi = sharding.shard_bounds(n_docs, world, rank)
    sched = sharding.allreduce_max_ints([hi - lo], "cpu")[0]
    rank_floor = -(-k // world)
    cand = [[] for _ in range(nq)]                       # (score, global id)
    tau = np.full(nq, -np.inf)
    r_lo, r_hi, n_coll = 0, min(sched, 8), 0
    while True:
        for q in range(nq):
            for d in range(min(r_lo, hi - lo), min(r_hi, hi - lo)):
                s = scores[q, lo + d]
                if s > tau[q]:
                    cand[q].append((s, lo + d))
        last = r_hi >= sched
        if not last:
            loc = torch.tensor([sorted((c[0] for c in cand[q]), reverse=True)[rank_floor - 1] if len(cand[q]) >= rank_floor
                                else -np.inf for q in range(nq)], dtype=torch.float64)
            floor = sharding.allreduce_min(loc).numpy()
            n_coll += 1
        for q in range(nq):
            best = sorted(cand[q], key=lambda c: (-c[0], c[1]))
            if len(best) >= k:
                best = best[:k]
                tau[q] = max(tau[q], best[-1][0])        # strict: a later local doc that ties loses (higher id)
            if not last and np.isfinite(floor[q]):
                best = [c for c in best if c[0] >= floor[q]]
                tau[q] = max(tau[q], np.nextafter(floor[q], -np.inf))   # ties with the floor stay
            cand[q] = best
        if last:
            break
        r_lo, r_hi = r_hi, min(sched, r_hi * 2)
    counts = sharding.allreduce_max_ints([n_coll, -n_coll], "cpu")
    assert counts[0] == n_coll and -counts[1] == n_coll           # same number of collectives on every rank
    ls = torch.full((nq, k), float("-inf"), dtype=torch.float64)
    li = torch.full((nq, k), -1, dtype=torch.int32)
    for q in range(nq):
        for j, (s, d) in enumerate(cand[q][:k]):
            ls[q, j], li[q, j] = s, d
    ms, mi = sharding.gather_merge_topk(ls, li, k, merge=_np_merge)
    ref = np.stack([np.lexsort((np.arange(n_docs), -scores[q]))[:k] for q in range(nq)])
    assert np.array_equal(mi.numpy(), ref)


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
        # the helpers of the cross-shard threshold exchange
        assert sharding.allreduce_max_ints([hi - lo, 7 + rank], "cpu") == [max(h - l for l, h in
                                                                               (sharding.shard_bounds(n_docs, world, r) for r in range(world))),
                                                                           7 + world - 1]
        t = torch.tensor([float(rank), 5.0 - rank, float("-inf") if rank == 0 else 3.0], dtype=torch.float64)
        assert sharding.allreduce_min(t).tolist() == [0.0, 5.0 - (world - 1), float("-inf")]
        _floor_protocol(rank, world, scores, k)
        # sharded query upload: every rank copies 1/G of the large query tensors and all-gathers them (HybridQueries.to_sharded)
        from fusion_b200.hybrid_engine import HybridQueries
        g = torch.Generator().manual_seed(5)
        hq = HybridQueries(lex_ptr=torch.arange(nq + 1, dtype=torch.int32), lex_term=torch.arange(nq, dtype=torch.int32),
                           dense=torch.randn((nq, 8), generator=g), colbert=torch.randn((nq, 3, 4), generator=g).bfloat16())
        dq, h2d = hq.to_sharded("cpu")
        assert torch.equal(dq.dense, hq.dense) and torch.equal(dq.colbert, hq.colbert) and torch.equal(dq.lex_term, hq.lex_term)
        per = -(-nq // world)
        rows = max(0, min(nq, (rank + 1) * per) - min(nq, rank * per))
        assert h2d == (nq + 1) * 4 + nq * 4 + rows * (8 * 4 + 12 * 2)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_topk_exchange():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29641, 101, 7, 5, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_three_rank_uneven_shards():
    """100 docs over 3 ranks (34 / 33 / 33): the small shards run an empty last round so the collectives line up."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(3, 29651, 100, 5, 6, ret), nprocs=3, join=True)
    assert all(ret.get(r) for r in range(3))


def test_shard_bounds_cover_the_corpus():
    for n, w in ((8841823, 8), (10, 3), (5, 8)):
        b = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert sharding.query_slice(10, 4, 3) == (9, 10) and sharding.query_slice(10, 4, 0) == (0, 3)
