"""Multi-GPU parity (run under torchrun on G GPUs of one box): the corpus-sharded hybrid pipeline over NCCL must return,
for every rank's query slice, exactly what ONE unsharded index returns.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_parity.py

Every rank builds the same small synthetic corpus (fixed seeds), keeps its contiguous doc range as the shard, runs
HybridSearcher over the process group, and also computes the single-index answer locally for its query slice.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_b200 import sharding, synth  # noqa: E402
from fusion_b200.hybrid_engine import HybridQueries, HybridSearcher  # noqa: E402
from fusion_b200.index import DenseIndex, LexicalIndex, SparseIndex, TokenStore, sparse_queries  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_docs, nq, k, vocab, sv, dim = 40000, 96, 200, 3000, 1500, 128
    (dptr, dtok), (qptr, qtok) = synth.c3_lexical(n_docs, nq, vocab)
    sp_d = synth.splade_vectors(n_docs, sv, 40, 4, 120, seed=311)
    sp_q = synth.splade_vectors(nq, sv, 12, 2, 40, seed=312)
    emb = torch.from_numpy(synth.dense_embeddings(n_docs, dim, seed=201))
    qemb = torch.from_numpy(synth.dense_embeddings(nq, dim, seed=202))
    tptr, temb = synth.colbert_tokens(n_docs, 128, 30, 4, 90, seed=401)
    qtok_c = torch.from_numpy(synth.colbert_queries(nq, 32, 128, seed=402)).to(dev).bfloat16()

    def build(lo, hi, full_lex=None):
        p = dptr[lo:hi + 1] - dptr[lo]
        t = dtok[dptr[lo]:dptr[hi]]
        if full_lex is None:
            lex = LexicalIndex(p, t, vocab, "bm25", 0.9, 0.4, device=dev, doc_base=lo, tile_docs=1024,
                               stats_reduce=(lambda n, df, s: sharding.allreduce_lexical_stats(n, df, s, dev)) if hi - lo < n_docs else None)
        else:
            lex = full_lex
        sp_ptr = sp_d[0][lo:hi + 1] - sp_d[0][lo]
        sl = slice(sp_d[0][lo], sp_d[0][hi])
        spx = SparseIndex(sp_ptr, sp_d[1][sl], sp_d[2][sl], sv, "cos_sim", device=dev, doc_base=lo, tile_docs=1024)
        den = DenseIndex.build(emb[lo:hi].to(dev), "cos_sim", doc_base=lo)
        tp = torch.from_numpy(tptr[lo:hi + 1] - tptr[lo]).to(dev)
        te = torch.from_numpy(temb[tptr[lo]:tptr[hi]]).to(dev).bfloat16()
        return lex, spx, den, TokenStore(tp, te, lo)

    q = HybridQueries()
    q.lex_ptr = torch.from_numpy(qptr.astype(np.int32)).to(dev)
    q.lex_term = torch.from_numpy(np.where(qtok < vocab, qtok, -1).astype(np.int32)).to(dev)
    q.sp_ptr, q.sp_term, q.sp_weight = sparse_queries(sp_q[0], sp_q[1], sp_q[2], "cos_sim", dev)
    q.dense = qemb.to(dev)
    q.colbert = qtok_c

    lo, hi = sharding.shard_bounds(n_docs, world, rank)
    shard = HybridSearcher(*build(lo, hi), k=k, fusion="nsf", normalization="z-score")
    lists_s = shard.retrieve(q)
    fused_s = shard.fuse(lists_s)
    torch.cuda.synchronize()

    # single-index answer, computed locally without collectives
    saved = dist.group.WORLD
    full = HybridSearcher(*build(0, n_docs), k=k, fusion="nsf", normalization="z-score", shard_sync=False)
    full.world, full.rank = 1, 0
    lists_f = full.retrieve(q)
    fused_f = full.fuse(lists_f)
    qlo, qhi = sharding.query_slice(nq, world, rank)
    ok = True
    for name in lists_f:
        fs, fi = lists_f[name][0][qlo:qhi], lists_f[name][1][qlo:qhi]
        ss, si = lists_s[name][0][: qhi - qlo], lists_s[name][1][: qhi - qlo]
        if name == "bm25":
            good = torch.equal(fi, si) and torch.equal(fs, ss)
        else:
            good = torch.allclose(fs, ss, rtol=1e-5, atol=1e-5) and float((fi == si).float().mean()) > 0.995
        print(f"rank {rank} {name}: {'OK' if good else 'MISMATCH'} ids equal {float((fi == si).float().mean()):.4f}", flush=True)
        ok &= good
    fi, fsn = fused_f[0][qlo:qhi], fused_f[1][qlo:qhi]
    si, ssn = fused_s[0][: qhi - qlo], fused_s[1][: qhi - qlo]
    agree = float((fi[:, :50] == si[:, :50]).float().mean())
    print(f"rank {rank} fused top-50 ids equal {agree:.4f}", flush=True)
    ok &= agree > 0.98
    # exhaustive ColBERT search_all over the SHARDED token store: shards' top-k merged over NCCL == one store
    from fusion_b200.retrievers.hybrid import Ranker
    qs = qtok_c[:16]
    es, ei = Ranker.maxsim_search_tensors(qs, shard.tokens, 50, group=dist.group.WORLD, chunk_pairs=16 * 3000)
    fs, fi = Ranker.maxsim_search_tensors(qs, full.tokens, 50, group=False, chunk_pairs=16 * 7000)
    good = torch.allclose(es, fs, rtol=1e-5, atol=1e-4) and float((ei == fi).float().mean()) > 0.995
    print(f"rank {rank} exhaustive colbert: {'OK' if good else 'MISMATCH'} ids equal {float((ei == fi).float().mean()):.4f}", flush=True)
    if not good:
        print(f"rank {rank} sharded {es[0, :6].tolist()} {ei[0, :6].tolist()}\nrank {rank} full    {fs[0, :6].tolist()} {fi[0, :6].tolist()}", flush=True)
    ok &= good
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("MULTI_GPU_PARITY", "PASS" if int(t) == 0 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 0 else 1)


if __name__ == "__main__":
    main()
