"""GPU parity on the edges: empty / out-of-vocabulary queries, k larger than the corpus, corpora smaller than one tile,
odd query-token counts, single-entry lists, the drop-in list-of-dict adapters, and size-independent properties at the
benchmark's tile sizes."""
import numpy as np
import pytest
import torch

from fusion_b200 import synth
from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import maxsim as omaxsim

pytestmark = pytest.mark.gpu


def test_bm25_adapter_matches_reference_shapes_on_edges():
    """search / search_all keep the reference's list-of-dict shape; empty and all-OOV queries rank every doc at 0.0 in
    index order (bm25.py:100-106), top_k > N returns N entries."""
    from fusion_b200.retrievers.bm25 import BM25
    docs = ["a b c", "b c d d", "e", "a a a e f", "c"]
    r = BM25(docs, k1=0.9, b=0.4)
    for q in ("", "zzz qqq"):
        res = r.search(q, top_k=3)
        assert [x["corpus_id"] for x in res] == [0, 1, 2] and all(x["score"] == 0.0 for x in res)
    res = r.search_all(["a e", "d", ""], top_k=50)
    assert [len(x) for x in res] == [5, 5, 5]
    o = obm25.LexicalOracle(*_csr(docs), 6, "bm25", 0.9, 0.4)
    ids, sc = o.search_ids(np.array([0, 4]), 5)            # "a e"
    assert [x["corpus_id"] for x in res[0]] == ids.tolist()
    assert [x["score"] for x in res[0]] == sc.tolist()
    assert isinstance(res[0][0]["score"], float) and isinstance(res[0][0]["corpus_id"], int)


def _csr(docs):
    vocab, ptr, toks = {}, [0], []
    for d in docs:
        for w in d.split():
            toks.append(vocab.setdefault(w, len(vocab)))
        ptr.append(len(toks))
    return np.asarray(ptr, dtype=np.int64), np.asarray(toks, dtype=np.int32)


@pytest.mark.parametrize("n_docs", [1, 7, 255, 1025])
def test_bm25_tiny_and_ragged_corpora_vs_oracle(n_docs):
    """Corpora smaller than a tile, one doc past a tile boundary, k > N."""
    from fusion_b200 import ops
    from fusion_b200.index import LexicalIndex
    vocab = 50
    (dptr, dtok), (qptr, qtok) = synth.c3_lexical(n_docs, 12, vocab)
    ix = LexicalIndex(dptr, dtok, vocab, "bm25", 0.9, 0.4, tile_docs=1024, tiled_min=4, dense_frac=0.3)
    q_ptr = torch.from_numpy(qptr.astype(np.int32)).cuda()
    q_term = torch.from_numpy(np.where(qtok < vocab, qtok, -1).astype(np.int32)).cuda()
    k = 40
    sc, ids = ops.sparse_topk(ix.view(), q_ptr, q_term, None, k)
    o = obm25.LexicalOracle(dptr, dtok, vocab, "bm25", 0.9, 0.4)
    for qi in range(12):
        t = qtok[qptr[qi]:qptr[qi + 1]]
        eids, esc = o.search_ids(np.where(t < vocab, t, -1), k)
        assert np.array_equal(ids[qi].cpu().numpy(), eids), (n_docs, qi)
        assert np.array_equal(sc[qi].cpu().numpy(), esc), (n_docs, qi)


def test_queries_longer_than_the_kernel_term_table_are_scored_exactly():
    """The reference has no query-length limit (bm25.py:149-156 loops over query.split()).  The kernels hold 128 terms per
    CTA: longer queries are flagged (never truncated) and scored in chunks whose sums continue in query order - ids and
    fp64 scores bit-equal to the oracle, for the top-k entry, the full-ranking entry and a batch mixing long and short queries."""
    from fusion_b200 import ops
    from fusion_b200.index import LexicalIndex
    vocab = 60
    (dptr, dtok), _ = synth.c3_lexical(700, 1, vocab)
    ix = LexicalIndex(dptr, dtok, vocab, "bm25", 0.9, 0.4, tile_docs=256, tiled_min=8)
    rng = np.random.default_rng(9)
    lens = [300, 5, 129, 1, 128, 257]
    toks = [rng.integers(0, vocab + 3, n) for n in lens]                 # ids >= vocab are out of vocabulary
    q_ptr = np.zeros(len(lens) + 1, dtype=np.int32)
    np.cumsum(lens, out=q_ptr[1:])
    q_term = np.concatenate(toks).astype(np.int32)
    q_term = np.where(q_term >= vocab, -1, q_term).astype(np.int32)
    o = obm25.LexicalOracle(dptr, dtok, vocab, "bm25", 0.9, 0.4)
    sc, ids = ops.sparse_topk(ix.view(), torch.from_numpy(q_ptr).cuda(), torch.from_numpy(q_term).cuda(), None, 50)
    full = ops.sparse_scores(ix.view(), torch.from_numpy(q_ptr).cuda(), torch.from_numpy(q_term).cuda()).cpu().numpy()
    for qi in range(len(lens)):
        eids, esc = o.search_ids(q_term[q_ptr[qi]:q_ptr[qi + 1]], 50)
        assert np.array_equal(ids[qi].cpu().numpy(), eids), qi
        assert np.array_equal(sc[qi].cpu().numpy(), esc), qi
        assert np.array_equal(np.sort(full[qi])[::-1][:50], esc), qi


def test_splade_queries_longer_than_the_term_table():
    from fusion_b200 import ops
    from fusion_b200.index import SparseIndex, sparse_queries
    from oracle import dense as odense
    vocab, n_docs, k = 900, 3000, 40
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 60, 8, 200, seed=311)
    qp, qt, qw = synth.splade_vectors(4, vocab, 200, 150, 300, seed=312)         # 150 - 300 terms per query
    qp2, qt2, qw2 = synth.splade_vectors(3, vocab, 12, 2, 40, seed=313)
    qp = np.concatenate([qp, qp[-1] + qp2[1:]]); qt = np.concatenate([qt, qt2]); qw = np.concatenate([qw, qw2])
    ix = SparseIndex(dp, dt, dw, vocab, "cos_sim", tile_docs=512, tiled_min=16, head_dim=64)
    q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "cos_sim", ix.device)
    sc, ids = ix.topk(q_ptr, q_term, q_w, k)
    dd, qd = torch.from_numpy(synth.densify(dp, dt, dw, vocab)), torch.from_numpy(synth.densify(qp, qt, qw, vocab))
    esc, eids = odense.topk_tensors(qd, dd, k, "cos_sim")
    torch.testing.assert_close(sc.cpu(), esc, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("nq,n_docs,k", [(1, 5, 10), (3, 255, 255), (129, 257, 7)])
def test_dense_small_shapes_vs_oracle(nq, n_docs, k):
    """Fewer docs than one UMMA tile, one past it, one query, an odd number of query tiles (CTA pair padding), k > N."""
    from fusion_b200.retrievers.hybrid import Ranker
    q = torch.from_numpy(synth.dense_embeddings(nq, 64, seed=81))
    d = torch.from_numpy(synth.dense_embeddings(n_docs, 64, seed=82))
    sc, ids = Ranker.dense_search_tensors(q.cuda(), d.cuda(), k, "dot")
    esc, eids = odense.topk_tensors(q, d, min(k, n_docs), "dot")
    assert sc.shape == (nq, min(k, n_docs))
    torch.testing.assert_close(sc.cpu(), esc, rtol=1e-5, atol=1e-4)
    assert (ids.cpu().long() == eids.long()).float().mean() > 0.99


@pytest.mark.parametrize("lq", [1, 33, 100])
def test_maxsim_odd_query_lengths(lq):
    """Query token counts that are not multiples of 8 / 32 (partial epilogue warps, TMA boxes past the last query)."""
    from fusion_b200 import ops
    ptr, emb = synth.colbert_tokens(200, 128, 40, 1, 140, seed=91)
    q = synth.colbert_queries(3, lq, 128, seed=92)
    cand = np.random.default_rng(3).integers(0, 200, (3, 65)).astype(np.int32)
    out = ops.maxsim(torch.from_numpy(q).cuda().bfloat16(), torch.from_numpy(ptr).cuda(), torch.from_numpy(emb).cuda().bfloat16(),
                     torch.from_numpy(cand).cuda())
    ref = omaxsim.maxsim_scores(torch.from_numpy(q), torch.from_numpy(ptr), torch.from_numpy(emb), torch.from_numpy(cand))
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-4)


def test_fuse_single_entry_and_empty_lists():
    """A one-element list has std = NaN under torch.std (hybrid.py:262); an empty list contributes nothing."""
    from fusion_b200 import ops
    from oracle import fusion as ofusion
    ids = [np.array([[5]], dtype=np.int32), np.array([[7, 5, 9]], dtype=np.int32)]
    sc = [np.array([[2.0]]), np.array([[3.0, 1.0, 0.5]])]
    lists = [(torch.from_numpy(i).cuda(), torch.from_numpy(s).cuda(), None) for i, s in zip(ids, sc)]
    for method, norm in (("rrf", None), ("bcf", None), ("nsf", "min-max"), ("nsf", "none")):
        w = [0.5, 0.5] if method == "nsf" else None
        oi, os_, on = ops.fuse(lists, method, norm, w)
        eids, esc = ofusion.fuse_query([i[0] for i in ids], [s[0] for s in sc], method, norm, w)
        assert oi[0, :int(on[0])].cpu().tolist() == eids
        np.testing.assert_allclose(os_[0, :int(on[0])].cpu().numpy(), np.asarray(esc, dtype=np.float64), rtol=1e-6)
    lens = torch.tensor([0], dtype=torch.int32).cuda()
    oi, os_, on = ops.fuse([(lists[0][0], lists[0][1], lens), lists[1]], "rrf")
    assert oi[0, :int(on[0])].cpu().tolist() == [7, 5, 9]


def test_full_size_tile_properties():
    """Size-independent properties at the benchmark's tile sizes on a corpus spanning many tile groups:
    linearity (scoring a query twice = 2 x scores for TF-IDF), idempotence of the ranking, sortedness and uniqueness."""
    from fusion_b200 import ops
    from fusion_b200.index import LexicalIndex
    n_docs, vocab, nq, k = 300_000, 30_000, 64, 1000
    (dptr, dtok), (qptr, qtok) = synth.c3_lexical(n_docs, nq, vocab)
    ix = LexicalIndex(dptr, dtok, vocab, "bm25", 0.9, 0.4)
    q_ptr = torch.from_numpy(qptr.astype(np.int32)).cuda()
    q_term = torch.from_numpy(np.where(qtok < vocab, qtok, -1).astype(np.int32)).cuda()
    sc, ids = ops.sparse_topk(ix.view(), q_ptr, q_term, None, k)
    sc2, ids2 = ops.sparse_topk(ix.view(), q_ptr, q_term, None, k)
    assert torch.equal(sc, sc2) and torch.equal(ids, ids2)                         # deterministic
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())                                   # sorted
    tie = sc[:, 1:] == sc[:, :-1]
    assert bool((ids[:, 1:][tie] > ids[:, :-1][tie]).all())                        # ties by lower doc id
    assert all(len(set(r.tolist())) == k for r in ids.cpu())                       # no document twice
    # the top-k of a query equals the top-k of the same query against every score materialised (two code paths)
    full = ops.sparse_scores(ix.view(), q_ptr[:5].contiguous(), q_term[: int(q_ptr[4])].contiguous())
    rs, ri = ops.rank_rows(full, k)
    assert torch.equal(ri, ids[:4]) and torch.equal(rs, sc[:4])
