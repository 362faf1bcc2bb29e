"""GPU parity of the DROP-IN ENTRY POINTS (8b): every mirror of a reference function on the hot
path is executed with the reference's own argument shapes and compared with the verbatim-reference goldens or the oracle.

  Ranker.bm25_search / single_vector_search / multi_vector_search    src/retrievers/hybrid.py:50-137
  BM25.update_params / save_indexes (+ load_indexes)                  src/retrievers/bm25.py:117-126,158-161
  BaseModel.search / compute_batchwise_similarity / compute_pairwise_similarity   src/retrievers/splade/base.py:173-251
  InformationRetrievalEvaluatorCustom.compute_metrices                src/utils/sentence_transformers.py:314-393
  CustomSearcher.search_all                                           src/utils/colbert_ir.py:245-255
"""
import os
import pickle

import numpy as np
import pytest
import torch

from fusion_b200 import synth

pytestmark = pytest.mark.gpu


class _DenseEncoder:
    """Stand-in for a stock sentence-transformers model: texts are 'q<i>' / 'd<i>' and map to injected embeddings."""

    def __init__(self, q, d):
        self.q, self.d = torch.as_tensor(q), torch.as_tensor(d)

    def encode(self, sentences, batch_size=32, convert_to_tensor=True, show_progress_bar=False, **kw):
        rows = [(self.q if s[0] == "q" else self.d)[int(s[1:])] for s in sentences]
        return torch.stack(rows)


class SPLADEStub(_DenseEncoder):
    """Class name starts with SPLADE: ``single_vector_search`` then passes ``query_mode`` like the reference (hybrid.py:99-100)."""

    def encode(self, sentences, query_mode=False, **kw):
        assert all((s[0] == "q") == query_mode for s in sentences)
        return super().encode(sentences)


# ---------------------------------------------------------------------------------------------------- a4: Ranker.bm25_search
@pytest.mark.parametrize("return_topk", [None, 25])
def test_ranker_bm25_search_dict_corpus_and_id_mapping(golden_dir, return_topk):
    from fusion_b200.retrievers.hybrid import Ranker
    g = np.load(os.path.join(golden_dir, "lexical_small.npz"))
    docs, queries = [str(x) for x in g["docs"]], [str(x) for x in g["queries"]]
    pids = [1000 + 7 * i for i in range(len(docs))]                       # corpus ids are not the row numbers
    res = Ranker.bm25_search(queries, dict(zip(pids, docs)), do_preprocessing=False, k1=2.5, b=0.2, return_topk=return_topk)
    k = return_topk or len(docs)
    assert len(res) == len(queries)
    for qi in range(len(queries)):
        assert [x["corpus_id"] for x in res[qi]] == [pids[i] for i in g["bm25_ids"][qi, :k]]
        assert np.array_equal(np.array([x["score"] for x in res[qi]]), g["bm25_scores"][qi, :k])
        assert isinstance(res[qi][0]["score"], float)


def test_ranker_bm25_search_preprocessing_hook():
    from fusion_b200.retrievers.hybrid import Ranker

    class Lower:
        def preprocess(self, texts, lemmatize=True):
            assert lemmatize
            return [t.lower() for t in texts]

    corpus = {5: "Le Chat dort", 9: "le chien court", 11: "un oiseau"}
    with pytest.raises(ImportError):
        Ranker.bm25_search(["chat"], corpus, do_preprocessing=True, k1=1.2, b=0.75)
    res = Ranker.bm25_search(["CHAT"], corpus, do_preprocessing=True, k1=1.2, b=0.75, preprocessor=Lower())
    assert res[0][0]["corpus_id"] == 5 and res[0][0]["score"] > 0 and res[0][1]["score"] == 0.0


# ---------------------------------------------------------------------------------------------------- a3: update_params, save/load
def test_bm25_update_params_equals_fresh_index(golden_dir):
    from fusion_b200.retrievers.bm25 import BM25
    g = np.load(os.path.join(golden_dir, "lexical_small.npz"))
    docs, queries = [str(x) for x in g["docs"]], [str(x) for x in g["queries"]]
    r = BM25(docs, k1=0.9, b=0.4)
    sc0, id0 = r.search_all_tensors(queries, top_k=len(docs))
    assert np.array_equal(id0.cpu().numpy(), g["bm25_mm_ids"])
    r.update_params(k1=2.5, b=0.2)                                        # bm25.py:158-161
    assert (r.k1, r.b) == (2.5, 0.2)
    sc, ids = r.search_all_tensors(queries, top_k=len(docs))
    assert np.array_equal(ids.cpu().numpy(), g["bm25_ids"]) and np.array_equal(sc.cpu().numpy(), g["bm25_scores"])
    assert r.score(queries[3], 17) == float(g["bm25_scores"][3][list(g["bm25_ids"][3]).index(17)])


def test_bm25_save_indexes_reference_pickles_and_load(tmp_path, golden_dir):
    from fusion_b200.retrievers.bm25 import BM25, TFIDF
    g = np.load(os.path.join(golden_dir, "lexical_small.npz"))
    docs, queries = [str(x) for x in g["docs"]], [str(x) for x in g["queries"]]
    r = BM25(docs, k1=2.5, b=0.2)
    r.save_indexes(str(tmp_path), "lleqa")
    # the reference's four pickles (bm25.py:117-126): a set, a dict of dicts, a Counter, a dict
    objs = {n: pickle.load(open(tmp_path / f"bm25_{n}_lleqa.pkl", "rb")) for n in ("vocab", "tf", "df", "idf")}
    words = set(w for d in docs for w in d.split())
    assert objs["vocab"] == words
    w0 = docs[0].split()[0]
    assert objs["tf"][w0][0] == docs[0].split().count(w0)
    assert objs["df"][w0] == sum(1 for d in docs if w0 in d.split())
    assert set(objs["idf"]) == words
    r2 = BM25.load_indexes(str(tmp_path), "lleqa", corpus=docs)
    assert (r2.k1, r2.b) == (2.5, 0.2)
    sc, ids = r2.search_all_tensors(queries, top_k=len(docs))
    assert np.array_equal(ids.cpu().numpy(), g["bm25_ids"]) and np.array_equal(sc.cpu().numpy(), g["bm25_scores"])
    t = TFIDF(docs)
    t.save_indexes(str(tmp_path), "x", reference_pickles=False)
    sc, ids = TFIDF.load_indexes(str(tmp_path), "x").search_all_tensors(queries, top_k=len(docs))
    assert np.array_equal(ids.cpu().numpy(), g["tfidf_ids"]) and np.array_equal(sc.cpu().numpy(), g["tfidf_scores"])
    with pytest.raises(Exception, match="device_tokenizer"):
        BM25(docs[:50], k1=0.9, b=0.4, device_tokenizer=True).save_indexes(str(tmp_path), "y")


# ---------------------------------------------------------------------------------------------------- a5: single_vector_search
def _check_dense_lists(res, g, sim, pids, k):
    for qi in range(len(res)):
        got_s = np.array([x["score"] for x in res[qi]])
        np.testing.assert_allclose(got_s, g[f"{sim}_scores"][qi, :k], rtol=1e-5, atol=1e-5)
        cut = g[f"{sim}_scores"][qi, k - 1]
        a = {x["corpus_id"] for x in res[qi] if x["score"] > cut + 1e-5}
        b = {pids[i] for i, s in zip(g[f"{sim}_ids"][qi, :k], g[f"{sim}_scores"][qi, :k]) if s > cut + 1e-5}
        assert a == b, qi
        assert all(res[qi][j]["score"] >= res[qi][j + 1]["score"] for j in range(len(res[qi]) - 1))


def test_ranker_single_vector_search_dense_encoder(golden_dir):
    """verbatim ``BaseModel.search`` golden (= semantic_search's algorithm) through the entry point, with an encoder
    object and a non-trivial id mapping; return_topk=50 (filter GEMM + exact rescoring needs d % 64 == 0: d = 64)."""
    from fusion_b200.retrievers.hybrid import Ranker
    g = np.load(os.path.join(golden_dir, "dense_small.npz"))
    nd, nq = len(g["d"]), len(g["q"])
    pids = [5 * i + 3 for i in range(nd)]
    corpus = {pid: f"d{i}" for i, pid in enumerate(pids)}
    res = Ranker.single_vector_search([f"q{i}" for i in range(nq)], corpus, _DenseEncoder(g["q"], g["d"]), return_topk=50)
    _check_dense_lists(res, g, "cos_sim", pids, 50)
    # return_topk=None ranks every document (hybrid.py:103: top_k = len(documents))
    full = Ranker.single_vector_search([f"q{i}" for i in range(2)], corpus, _DenseEncoder(g["q"], g["d"]))
    assert len(full[0]) == nd
    _check_dense_lists([r[:50] for r in full], g, "cos_sim", pids, 50)


def test_ranker_single_vector_search_splade_encoder():
    """A SPLADE-class encoder returns [*, V] activations; the entry point scores them through the sparse pipeline.  Oracle:
    the reference's dense cosine over the V-dim vectors (hybrid.py:101-103)."""
    from fusion_b200.retrievers.hybrid import Ranker
    from oracle import dense as odense
    vocab, n_docs, nq, k = 5000, 2500, 6, 100
    dp, dt, dw = synth.splade_vectors(n_docs, vocab, 60, 8, 200, seed=311)
    qp, qt, qw = synth.splade_vectors(nq, vocab, 12, 2, 40, seed=312)
    dd, qd = synth.densify(dp, dt, dw, vocab), synth.densify(qp, qt, qw, vocab)
    corpus = {100 + i: f"d{i}" for i in range(n_docs)}
    res = Ranker.single_vector_search([f"q{i}" for i in range(nq)], corpus, SPLADEStub(qd, dd), return_topk=k)
    esc, eids = odense.topk_tensors(torch.from_numpy(qd), torch.from_numpy(dd), k, "cos_sim")
    for qi in range(nq):
        np.testing.assert_allclose([x["score"] for x in res[qi]], esc[qi].numpy(), rtol=1e-5, atol=1e-5)
        cut = float(esc[qi, -1])
        assert {x["corpus_id"] for x in res[qi] if x["score"] > cut + 1e-5} == \
               {100 + int(i) for i, s in zip(eids[qi], esc[qi]) if s > cut + 1e-5}


# ---------------------------------------------------------------------------------------------------- a6 / a7: BaseModel
class _Model:
    def __init__(self, q, d, similarity):
        self.q, self.d, self.similarity = torch.as_tensor(q), torch.as_tensor(d), similarity

    def encode(self, texts, query_mode, batch_size=32):
        return self.q if query_mode else self.d


@pytest.mark.parametrize("sim", ["cos_sim", "dot"])
def test_basemodel_search_golden(golden_dir, sim):
    """``BaseModel.search`` (base.py:199-251) against its own verbatim golden: 'doc_id' keys, sorted by score descending."""
    from fusion_b200.retrievers.splade.base import BaseModel
    g = np.load(os.path.join(golden_dir, "dense_small.npz"))
    M = type("M", (_Model, BaseModel), {})
    res = M(g["q"], g["d"], sim).search(["q"] * len(g["q"]), ["d"] * len(g["d"]), batch_size=8, query_chunk_size=3,
                                        doc_chunk_size=1100, topk=50)
    assert set(res[0][0].keys()) == {"doc_id", "score"}
    ren = [[{"corpus_id": x["doc_id"], "score": x["score"]} for x in r] for r in res]
    _check_dense_lists(ren, g, sim, list(range(len(g["d"]))), 50)


@pytest.mark.parametrize("sim", ["cos_sim", "dot"])
def test_basemodel_similarity_functions(golden_dir, sim):
    """compute_batchwise_similarity (base.py:186-197) and compute_pairwise_similarity (:173-184) vs the oracle."""
    from fusion_b200.retrievers.splade.base import BaseModel
    from oracle import dense as odense
    g = np.load(os.path.join(golden_dir, "dense_small.npz"))
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"][:40])
    m = type("M", (_Model, BaseModel), {})(q, d, sim)
    ref = odense.similarity(q, d, sim)
    torch.testing.assert_close(m.compute_batchwise_similarity(q, d).cpu(), ref, rtol=1e-5, atol=1e-5)
    pair = m.compute_pairwise_similarity(q, d[: len(q)]).cpu()
    assert pair.shape == (len(q),)
    torch.testing.assert_close(pair, torch.diagonal(ref[:, : len(q)]), rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------- a8: compute_metrices
def test_evaluator_compute_metrices_vs_restated_reference(golden_dir):
    from fusion_b200.utils.sentence_transformers import InformationRetrievalEvaluatorCustom, cos_sim, dot_score
    from oracle import evaluator as oev
    g = np.load(os.path.join(golden_dir, "dense_small.npz"))
    nq, nd = len(g["q"]), len(g["d"])
    rng = np.random.default_rng(3)
    queries = {f"Q{i}": f"q{i}" for i in range(nq)}
    corpus = {f"D{i}": f"d{i}" for i in range(nd)}
    rel = {f"Q{i}": {f"D{int(j)}" for j in np.concatenate([g["cos_sim_ids"][i, rng.choice(50, 3, replace=False)],
                                                           rng.choice(nd, 2)])} for i in range(nq)}
    rel["Q6"] = set()                                                     # a query without relevant docs is dropped (:268-270)
    ks = dict(mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1, 3, 5, 10], precision_recall_at_k=[1, 3, 5, 10], map_at_k=[40])
    ev = InformationRetrievalEvaluatorCustom(queries, corpus, rel, corpus_chunk_size=700, **ks,
                                             score_functions={"cos_sim": cos_sim, "dot_score": dot_score})
    assert ev.queries_ids == [f"Q{i}" for i in range(nq - 1)]
    enc = _DenseEncoder(g["q"], g["d"])
    got = ev.compute_metrices(enc, corpus_embeddings=torch.from_numpy(g["d"]))
    assert set(got) == {"cos_sim", "dot_score"}
    qe = torch.from_numpy(g["q"][: nq - 1])
    for name in ("cos_sim", "dot_score"):
        lists = oev.search(qe, torch.from_numpy(g["d"]), ev.corpus_ids, 40, name, corpus_chunk_size=700)
        want = oev.compute_metrics(lists, ev.queries_ids, rel, nq - 1, **ks)
        for metric, val in want.items():
            if isinstance(val, dict):
                for k, v in val.items():
                    assert got[name][metric][k] == pytest.approx(v, abs=1e-12), (name, metric, k)
            else:
                assert got[name][metric] == pytest.approx(val, abs=1e-12)
    # corpus encoded on the fly in chunks (corpus_embeddings=None) gives the same numbers; __call__ returns map@max_k
    got2 = ev.compute_metrices(enc)
    assert got2["cos_sim"]["map@k"][40] == got["cos_sim"]["map@k"][40]
    assert ev(enc) == max(got[n]["map@k"][40] for n in got)
    # the score functions are callable like sentence_transformers.util.cos_sim
    torch.testing.assert_close(cos_sim(g["q"][:2], g["d"][:5]).cpu(),
                               torch.nn.functional.normalize(torch.from_numpy(g["q"][:2]), dim=1) @
                               torch.nn.functional.normalize(torch.from_numpy(g["d"][:5]), dim=1).t(), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------- a10: ColBERT entry points
class _TokEncoder:
    def __init__(self, g):
        self.g = g

    def encode_queries(self, queries):
        return torch.from_numpy(self.g["q"][[int(s[1:]) for s in queries]])

    def encode_docs(self, docs):
        rows = [int(s[1:]) for s in docs]
        ptr = self.g["tok_ptr"]
        lens = np.array([ptr[r + 1] - ptr[r] for r in rows])
        out_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=out_ptr[1:])
        emb = np.concatenate([self.g["tok_emb"][ptr[r]:ptr[r + 1]] for r in rows]) if rows else np.zeros((0, 128), np.float32)
        return torch.from_numpy(out_ptr), torch.from_numpy(emb)


def test_maxsim_kernel_vs_independent_fixture(golden_dir):
    """K3 against the committed padded-batch float64 fixture (oracle/make_golden_maxsim.py), not only against oracle/maxsim.py."""
    from fusion_b200 import ops
    g = np.load(os.path.join(golden_dir, "maxsim_small.npz"))
    sc = ops.maxsim(torch.from_numpy(g["q"]).cuda().to(torch.bfloat16), torch.from_numpy(g["tok_ptr"]).cuda(),
                    torch.from_numpy(g["tok_emb"]).cuda().to(torch.bfloat16), torch.from_numpy(g["cand"]).cuda())
    np.testing.assert_allclose(sc.cpu().numpy(), g["scores"], rtol=1e-5, atol=1e-4)


def test_ranker_multi_vector_search_and_custom_searcher(golden_dir):
    """Exhaustive MaxSim ranking through ``Ranker.multi_vector_search`` (hybrid.py:109-137) and ``CustomSearcher.search_all``
    (colbert_ir.py:245-255): every passage scored, result shapes of the reference ({'corpus_id','score'} / (pid, rank, score))."""
    from fusion_b200.index import TokenStore
    from fusion_b200.retrievers.hybrid import Ranker
    from fusion_b200.utils.colbert_ir import CustomSearcher
    g = np.load(os.path.join(golden_dir, "maxsim_small.npz"))
    n_docs, nq = len(g["tok_ptr"]) - 1, len(g["q"])
    keep = [d for d in range(n_docs) if g["tok_ptr"][d + 1] > g["tok_ptr"][d]]       # (an empty passage cannot be indexed)
    enc = _TokEncoder(g)
    pids = {d: 900 + d for d in keep}
    res = Ranker.multi_vector_search([f"q{i}" for i in range(nq)], {pids[d]: f"d{d}" for d in keep}, enc, return_topk=10)
    # expected: the fixture's float64 formulation over ALL kept passages
    from oracle.make_golden_maxsim import colbert_score_padded
    for qi in range(nq):
        want = colbert_score_padded(g["q"][qi], g["tok_ptr"], g["tok_emb"], np.array(keep))
        order = np.argsort(-want, kind="stable")[:10]
        assert [x["corpus_id"] for x in res[qi]] == [pids[keep[j]] for j in order]
        np.testing.assert_allclose([x["score"] for x in res[qi]], want[order], rtol=1e-5, atol=1e-4)
    # the exhaustive search walks the store in chunks of passages and merges the chunks' lists: same result
    ptr, emb = enc.encode_docs([f"d{d}" for d in keep])
    st = TokenStore(ptr.cuda(), emb.cuda().to(torch.bfloat16))
    q16 = torch.from_numpy(g["q"]).cuda()
    s1, i1 = Ranker.maxsim_search_tensors(q16, st, 10)
    s2, i2 = Ranker.maxsim_search_tensors(q16, st, 10, chunk_pairs=nq * 17)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    searcher = CustomSearcher(TokenStore(ptr.cuda(), emb.cuda().to(torch.bfloat16)), encoder=enc)
    ranking = searcher.search_all({f"qid{i}": f"q{i}" for i in range(nq)}, k=5)
    assert list(ranking) == [f"qid{i}" for i in range(nq)]
    for qi in range(nq):
        want = colbert_score_padded(g["q"][qi], g["tok_ptr"], g["tok_emb"], np.array(keep))
        order = np.argsort(-want, kind="stable")[:5]
        assert [(p, r) for p, r, _ in ranking[f"qid{qi}"]] == [(int(j), r + 1) for r, j in enumerate(order)]


# ---------------------------------------------------------------------------------------------------- persistence of the four containers
def test_index_containers_save_and_load(tmp_path):
    """SURVEY section 5 (checkpoint / resume hook), bm25.py:117-126: every device index can be written to disk and restored
    without re-tokenising / re-sorting / re-normalising; the restored index returns identical results."""
    from fusion_b200 import ops
    from fusion_b200.index import DenseIndex, LexicalIndex, SparseIndex, TokenStore, sparse_queries
    (dptr, dtok), (qptr, qtok) = synth.c3_lexical(3000, 8, 500)
    lex = LexicalIndex(dptr, dtok, 500, "bm25", 0.9, 0.4, tile_docs=512, tiled_min=16, doc_base=7)
    lex.save(str(tmp_path / "lex"))
    lex2 = LexicalIndex.load(str(tmp_path / "lex"), tiled_min=16)
    qp = torch.from_numpy(qptr.astype(np.int32)).cuda()
    qt = torch.from_numpy(np.where(qtok < 500, qtok, -1).astype(np.int32)).cuda()
    a, b = ops.sparse_topk(lex.view(), qp, qt, None, 50, 7), ops.sparse_topk(lex2.view(), qp, qt, None, 50, 7)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and lex2.doc_base == 7 and lex2.avgdl == lex.avgdl

    dp, dt, dw = synth.splade_vectors(4000, 1500, 40, 4, 120, seed=311)
    sq = synth.splade_vectors(6, 1500, 12, 2, 40, seed=312)
    sp = SparseIndex(dp, dt, dw, 1500, "cos_sim", head_dim=64, doc_base=3)
    sp.save(str(tmp_path / "sp"))
    sp2 = SparseIndex.load(str(tmp_path / "sp"))
    q3 = sparse_queries(sq[0], sq[1], sq[2], "cos_sim", sp.device)
    a, b = sp.topk(*q3, 40), sp2.topk(*q3, 40)
    assert sp2.similarity == "cos_sim" and sp2.head is not None and sp2.head.head_dim == 64
    torch.testing.assert_close(a[0], b[0], rtol=1e-6, atol=1e-7)
    assert float((a[1] == b[1]).float().mean()) > 0.99

    emb = torch.from_numpy(synth.dense_embeddings(2000, 64, seed=201)).cuda()
    qe = torch.from_numpy(synth.dense_embeddings(5, 64, seed=202)).cuda()
    de = DenseIndex.build(emb, "cos_sim", doc_base=11)
    de.save(str(tmp_path / "de"))
    de2 = DenseIndex.load(str(tmp_path / "de"))
    assert torch.equal(de.d_f32, de2.d_f32) and torch.equal(de.d_bf16, de2.d_bf16) and de2.doc_base == 11
    q32, q16 = de.prepare_queries(qe)
    a = ops.dense_topk(q16, de.d_bf16, q32, de.d_f32, 30, margin=0.008, doc_base=11)
    b = ops.dense_topk(q16, de2.d_bf16, q32, de2.d_f32, 30, margin=0.008, doc_base=11)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])

    tptr, temb = synth.colbert_tokens(300, 128, 30, 4, 90, seed=401)
    ts = TokenStore(torch.from_numpy(tptr).cuda(), torch.from_numpy(temb).cuda().bfloat16(), 5)
    ts.save(str(tmp_path / "ts"))
    ts2 = TokenStore.load(str(tmp_path / "ts"))
    qtk = torch.from_numpy(synth.colbert_queries(3, 32, 128, seed=402)).cuda().bfloat16()
    cand = (torch.arange(40, dtype=torch.int32).cuda() + 5).expand(3, -1).contiguous()
    assert torch.equal(ops.maxsim(qtk, ts.tok_ptr, None, cand, 5, packed=ts.packed()),
                       ops.maxsim(qtk, ts2.tok_ptr, None, cand, 5, packed=ts2.packed()))
