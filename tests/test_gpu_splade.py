"""GPU parity: K2c, the SPLADE head-GEMM + tail-bound + exact-rescore pipeline (fz_splade_topk), against the dense oracle
(the reference scores SPLADE as a dense [., V] cosine, src/retrievers/hybrid.py:101-103 / splade/base.py:186-251) and
against the general inverted-index path."""
import numpy as np
import pytest
import torch

from fusion_b200 import synth

pytestmark = pytest.mark.gpu


def _sets_equal_above_cut(ids_a, sc_a, ids_b, sc_b, tol):
    for qi in range(ids_a.shape[0]):
        cut = float(sc_b[qi, -1])
        a = {int(i) for i, s in zip(ids_a[qi], sc_a[qi]) if s > cut + tol}
        b = {int(i) for i, s in zip(ids_b[qi], sc_b[qi]) if s > cut + tol}
        assert a == b, qi


@pytest.mark.parametrize("head_dim", [64, 192])
def test_splade_pipeline_vs_dense_oracle(head_dim):
    """scores 1e-5, top-k set equal (ties at the cut excluded), several rounds (cap forces a first round of 256 docs)."""
    from fusion_b200 import ops
    from fusion_b200.index import SparseIndex, sparse_queries
    from oracle import dense as odense
    vocab, n_docs, nq, k = 2000, 6000, 12, 100
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 60, 8, 200, seed=311)
    qp, qt, qw = synth.splade_vectors(nq, vocab, 12, 2, 40, seed=312)
    ix = SparseIndex(dp, dt, dw, vocab, "cos_sim", tile_docs=1024, tiled_min=64, head_dim=head_dim, tail_tile_docs=1024)
    assert ix.head is not None
    q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "cos_sim", ix.device)
    sc, ids = ops.splade_topk(ix, q_ptr, q_term, q_w, k, cap=512)
    dd, qd = torch.from_numpy(synth.densify(dp, dt, dw, vocab)), torch.from_numpy(synth.densify(qp, qt, qw, vocab))
    esc, eids = odense.topk_tensors(qd, dd, k, "cos_sim")
    torch.testing.assert_close(sc.cpu(), esc, rtol=1e-5, atol=1e-5)
    _sets_equal_above_cut(ids.cpu().numpy(), sc.cpu().numpy(), eids.numpy(), esc.numpy(), 1e-5)


@pytest.mark.parametrize("boot_docs", [0, 4096])
def test_splade_pipeline_equals_inverted_index_path(boot_docs):
    """Mid-size corpus, k = 1000, default cap: the fast path returns what the general inverted-index kernel returns
    (exact fp32 scores in both; the summation order differs, hence 1e-6 absolute)."""
    from fusion_b200 import ops
    from fusion_b200.index import SparseIndex, sparse_queries
    vocab, n_docs, nq, k = 8000, 120_000, 64, 1000
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 100, 8, 400, seed=311)
    qp, qt, qw = synth.splade_vectors(nq, vocab, 20, 2, 64, seed=312)
    ix = SparseIndex(dp, dt, dw, vocab, "cos_sim", head_dim=128, boot_docs=boot_docs)
    assert (ix.head.boot is not None) == (boot_docs > 0)
    q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "cos_sim", ix.device)
    sc, ids = ops.splade_topk(ix, q_ptr, q_term, q_w, k)
    sc2, ids2 = ops.sparse_topk(ix.view(), q_ptr, q_term, q_w, k)
    torch.testing.assert_close(sc, sc2, rtol=1e-5, atol=1e-6)
    _sets_equal_above_cut(ids.cpu().numpy(), sc.cpu().numpy(), ids2.cpu().numpy(), sc2.cpu().numpy(), 2e-6)
    assert float((ids == ids2).float().mean()) > 0.99


def test_splade_pipeline_fallback_few_matches_and_negative_weights():
    """Queries with fewer than k matching docs (zero-score docs fill up in doc-id order) and queries with a negative weight
    are handed back to the general path: the result must equal it exactly."""
    from fusion_b200 import ops
    from fusion_b200.index import SparseIndex, sparse_queries
    vocab, n_docs, nq, k = 3000, 3000, 8, 200
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 10, 2, 30, seed=5)
    qp, qt, qw = synth.splade_vectors(nq, vocab, 3, 1, 6, seed=6)
    qw = qw.copy()
    qw[qp[2]] = -qw[qp[2]]                    # one query with a negative weight
    qt = qt.copy()
    qt[qp[3]:qp[4]] = np.arange(vocab - (qp[4] - qp[3]), vocab)     # rare terms only: fewer than k matches
    ix = SparseIndex(dp, dt, dw, vocab, "cos_sim", tile_docs=512, tiled_min=16, head_dim=64, tail_tile_docs=512)
    q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "dot", ix.device)
    sc, ids = ops.splade_topk(ix, q_ptr, q_term, q_w, k)
    sc2, ids2 = ops.sparse_topk(ix.view(), q_ptr, q_term, q_w, k)
    torch.testing.assert_close(sc, sc2, rtol=1e-5, atol=1e-6)
    assert torch.equal(ids[2], ids2[2]) and torch.equal(ids[3], ids2[3])
    assert float((ids == ids2).float().mean()) > 0.97


def test_splade_pipeline_shards_with_cross_shard_floor():
    from fusion_b200 import ops
    from fusion_b200.index import SparseIndex, sparse_queries
    vocab, n_docs, nq, k, cap = 2000, 18000, 16, 100, 512
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 60, 8, 200, seed=311)
    qp, qt, qw = synth.splade_vectors(nq, vocab, 12, 2, 40, seed=312)
    full = SparseIndex(dp, dt, dw, vocab, "cos_sim", head_dim=64, tail_tile_docs=1024)
    q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "cos_sim", full.device)
    sc, ids = ops.splade_topk(full, q_ptr, q_term, q_w, k, cap=cap)
    kth = sc[:, -1].clone()
    parts = []
    for lo, hi in ((0, 6000), (6000, 12000), (12000, n_docs)):
        ix = SparseIndex(dp[lo:hi + 1] - dp[lo], dt[dp[lo]:dp[hi]], dw[dp[lo]:dp[hi]], vocab, "cos_sim", doc_base=lo,
                         head_dim=64, tail_tile_docs=1024)
        parts.append(ops.splade_topk(ix, q_ptr, q_term, q_w, k, doc_base=lo, cap=cap,
                                     sync=ops.ShardSync(lambda t: torch.minimum(t, kth, out=t), 3, 6000)))
    ms, mi = ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    torch.testing.assert_close(ms, sc, rtol=1e-5, atol=1e-6)
    assert float((mi == ids).float().mean()) > 0.98
