"""GPU parity: K1 (tcgen05 filter GEMM + fp32 rescoring) and K3 (tcgen05 MaxSim) against the oracle / goldens."""
import os

import numpy as np
import pytest
import torch

from fusion_b200 import synth
from oracle import dense as odense
from oracle import maxsim as omaxsim

pytestmark = pytest.mark.gpu


def _check_topk(sc, ids, esc, eids, tol):
    torch.testing.assert_close(sc.cpu(), esc, rtol=tol, atol=tol)
    for qi in range(sc.shape[0]):
        cut = float(esc[qi, -1])
        a = {int(i) for i, s in zip(ids[qi].cpu(), sc[qi].cpu()) if s > cut + tol}
        b = {int(i) for i, s in zip(eids[qi], esc[qi]) if s > cut + tol}
        assert a == b, qi


@pytest.mark.parametrize("sim", ["cos_sim", "dot"])
def test_dense_golden_small(golden_dir, sim):
    """Verbatim BaseModel.search output (splade/base.py:199-251) on 7 x 3000 x 64, k=50, exact mode."""
    from fusion_b200.retrievers.hybrid import Ranker
    g = np.load(os.path.join(golden_dir, "dense_small.npz"))
    q, d = torch.from_numpy(g["q"]).cuda(), torch.from_numpy(g["d"]).cuda()
    sc, ids = Ranker.dense_search_tensors(q, d, 50, sim)
    exp_s = torch.from_numpy(g[f"{sim}_scores"]).float()
    tol = 1e-5 * max(1.0, float(exp_s.abs().max()))
    torch.testing.assert_close(sc.cpu(), exp_s, rtol=1e-5, atol=tol)
    for qi in range(7):
        cut = float(exp_s[qi, -1])
        a = {int(i) for i, s in zip(ids[qi].cpu(), sc[qi].cpu()) if s > cut + tol}
        b = {int(i) for i, s in zip(g[f"{sim}_ids"][qi], exp_s[qi]) if s > cut + tol}
        assert a == b


def test_dense_scores_and_full_ranking():
    """Exact fp32 score matrix (CUDA cores) and the rank-every-document mode (hybrid.py:103 with top_k = N)."""
    from fusion_b200 import ops
    from fusion_b200.retrievers.hybrid import Ranker
    q = torch.from_numpy(synth.dense_embeddings(9, 96, seed=41)).cuda()
    d = torch.from_numpy(synth.dense_embeddings(3001, 96, seed=42)).cuda()
    ref = odense.similarity(q.cpu(), d.cpu(), "cos_sim")
    q32, _ = ops.normalize_rows(q)
    d32, _ = ops.normalize_rows(d)
    torch.testing.assert_close(ops.dense_scores(q32, d32).cpu(), ref, rtol=1e-5, atol=1e-6)
    sc, ids = Ranker.dense_search_tensors(q, d, 3001, "cos_sim")
    order = torch.sort(ref, dim=1, descending=True, stable=True)
    torch.testing.assert_close(sc.cpu(), order.values, rtol=1e-5, atol=1e-6)
    assert (ids.cpu().long() == order.indices).float().mean() > 0.999


@pytest.mark.parametrize("nq,n_docs,dim,k", [(130, 20000, 768, 100), (5, 9000, 128, 1000), (300, 70000, 64, 10)])
def test_dense_topk_exact_vs_oracle(nq, n_docs, dim, k):
    """tcgen05 filter + fp32 rescoring == fp32 oracle top-k within 1e-5 (several rounds, ragged last tiles)."""
    from fusion_b200.retrievers.hybrid import Ranker
    q = torch.from_numpy(synth.dense_embeddings(nq, dim, seed=51))
    d = torch.from_numpy(synth.dense_embeddings(n_docs, dim, seed=52))
    sc, ids = Ranker.dense_search_tensors(q.cuda(), d.cuda(), k, "cos_sim")
    esc, eids = odense.topk_tensors(q, d, k, "cos_sim")
    _check_topk(sc, ids, esc, eids, 1e-5)


def test_dense_topk_bf16_mode_overlap():
    """bf16 throughput mode (north-star): scores within 2e-3, top-k overlap >= 99.9 % excluding ties at the cutoff.  A tie
    in this mode is a reference score within the stated bf16 score tolerance (2e-3 relative) of the k-th one: two documents
    closer than the arithmetic's resolution cannot be ordered by it."""
    from fusion_b200.retrievers.hybrid import Ranker
    nq, n_docs, dim, k = 64, 50000, 768, 100
    q = torch.from_numpy(synth.dense_embeddings(nq, dim, seed=61))
    d = torch.from_numpy(synth.dense_embeddings(n_docs, dim, seed=62))
    sc, ids = Ranker.dense_search_tensors(q.cuda(), d.cuda(), k, "cos_sim", exact=False)
    esc, eids = odense.topk_tensors(q, d, k, "cos_sim")
    assert float((sc.cpu() - esc).abs().max()) < 2e-3
    hit = tot = 0
    for qi in range(nq):
        cut = float(esc[qi, -1])
        b = {int(i) for i, s in zip(eids[qi], esc[qi]) if s > cut + 2e-3 * abs(cut)}
        hit += len(b & set(ids[qi].cpu().tolist()))
        tot += len(b)
    assert hit / tot >= 0.999


def test_dense_exact_shards_with_cross_shard_floor():
    """Staged exact mode with the threshold exchange between the filter rounds (emulated all-reduce, uneven shards: the
    small shard runs empty rounds so that every shard issues the same collectives): merged == unsharded."""
    from fusion_b200 import ops
    nq, n, d, k, margin, cap = 24, 30000, 128, 100, 0.008, 512
    q = torch.from_numpy(synth.dense_embeddings(nq, d, seed=11)).cuda()
    docs = torch.from_numpy(synth.dense_embeddings(n, d, seed=12)).cuda()
    q32, q16 = ops.normalize_rows(q)
    d32, d16 = ops.normalize_rows(docs)
    sc, ids = ops.dense_topk(q16, d16, q32, d32, k, margin=margin, cap=cap)
    floor = sc[:, -1] - margin                       # every true top-k doc has a bf16 score above this
    counts, parts = [], []
    bounds = ((0, 13000), (13000, 24000), (24000, n))
    for lo, hi in bounds:
        calls = []

        def fake_min(t):
            calls.append(1)
            return torch.minimum(t, floor, out=t)
        parts.append(ops.dense_topk(q16, d16[lo:hi].contiguous(), q32, d32[lo:hi].contiguous(), k, margin=margin, doc_base=lo,
                                    cap=cap, tau_reduce=fake_min, n_shards=3, sched_docs=13000))
        counts.append(len(calls))
    assert len(set(counts)) == 1 and counts[0] >= 3
    ms, mi = ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    assert torch.equal(mi, ids) and torch.equal(ms, sc)


@pytest.mark.parametrize("lq,n_cand,max_len", [(64, 40, 300), (32, 33, 300), (128, 7, 300), (64, 70, 900), (17, 100, 60)])
def test_maxsim_vs_oracle(lq, n_cand, max_len):
    """Doc lengths 1..max_len (passages longer than one MMA group are cut into 2-4 pieces), candidates outside the
    shard, an empty document, exact multiples of the group capacity."""
    from fusion_b200 import ops
    nq, n_docs = 5, 300
    rng = np.random.Generator(np.random.PCG64(71))
    lens = rng.integers(1, max_len + 1, n_docs)
    lens[3] = 0
    lens[5], lens[6], lens[7] = 248, 496, 249
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    emb = rng.standard_normal((int(ptr[-1]), 128), dtype=np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    q = synth.colbert_queries(nq, lq, 128, seed=72)
    cand = rng.integers(0, n_docs, (nq, n_cand)).astype(np.int32)
    cand[0, 0] = 3
    cand[1, 1] = -1
    ptr_t, emb_t, q_t, cand_t = torch.from_numpy(ptr), torch.from_numpy(emb), torch.from_numpy(q), torch.from_numpy(cand)
    out = ops.maxsim(q_t.cuda().bfloat16(), ptr_t.cuda(), emb_t.cuda().bfloat16(), cand_t.cuda())
    ref = omaxsim.maxsim_scores(q_t, ptr_t, emb_t, cand_t)
    ref[ref == float("-inf")] = 0.0          # candidates outside the shard are skipped (score 0)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-4)


def test_maxsim_search_ranking():
    from fusion_b200.index import TokenStore
    from fusion_b200.retrievers.hybrid import Ranker
    ptr, emb = synth.colbert_tokens(500, 128, 70, 8, 180, seed=401)
    q = synth.colbert_queries(6, 64, 128, seed=402)
    store = TokenStore(torch.from_numpy(ptr).cuda(), torch.from_numpy(emb).cuda().bfloat16())
    sc, ids = Ranker.maxsim_search_tensors(torch.from_numpy(q).cuda(), store, 50)
    cand = torch.arange(500, dtype=torch.int32).expand(6, -1)
    ref = omaxsim.maxsim_scores(torch.from_numpy(q), torch.from_numpy(ptr), torch.from_numpy(emb), cand)
    order = torch.sort(ref, dim=1, descending=True, stable=True)
    torch.testing.assert_close(sc.cpu(), order.values[:, :50], rtol=1e-5, atol=1e-4)
    assert (ids.cpu().long() == order.indices[:, :50]).float().mean() > 0.98
