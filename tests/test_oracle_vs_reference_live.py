"""Differential tests of the oracle against the UNMODIFIED reference on fresh random inputs (beyond the committed
goldens).  They need /root/reference, which only the build container has: skipped everywhere else (the GPU box runs
`-m gpu` only and never reads the reference)."""
import contextlib
import copy
import io

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def _random_lists(rng, n_q, n, pool, kind):
    out = []
    for _ in range(n_q):
        m = int(rng.integers(1, n + 1))
        ids = rng.choice(pool, size=m, replace=False)
        if kind == "ties":
            sc = np.sort(rng.integers(0, 4, m).astype(np.float64))[::-1]
        elif kind == "f32":
            sc = np.sort(rng.random(m).astype(np.float32))[::-1].astype(np.float64)
        else:
            sc = np.sort(rng.normal(0, 3, m))[::-1]
        if kind == "dup" and m > 3:
            ids[m // 2] = ids[0]
        out.append([{"corpus_id": int(i), "score": float(s)} for i, s in zip(ids, sc)])
    return out


@pytest.mark.parametrize("seed", range(6))
def test_fusion_oracle_equals_reference_on_random_lists(seed):
    from oracle import fusion as ofusion
    mod = ref_loader.load_hybrid()
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    n_q, pool = 4, 90
    kinds = ["plain", "ties", "f32", "dup"]
    systems = [f"s{j}" for j in range(int(rng.integers(1, 5)))]
    lists = {s: _random_lists(rng, n_q, 40, pool, kinds[int(rng.integers(0, 4))]) for s in systems}
    w = rng.random(len(systems)) + 0.05
    weights = {s: float(x) for s, x in zip(systems, w / w.sum())}
    distrs = {s: np.quantile(rng.normal(0, 3, 2000), np.linspace(0, 1, 51)) for s in systems}
    cases = [("bcf", None), ("rrf", None)] + [("nsf", nm) for nm in
             ("none", "min-max", "z-score", "arctan", "percentile-rank", "normal-curve-equivalent")]
    for method, norm in cases:
        res = _quiet(mod.Aggregator.fuse, copy.deepcopy(lists), method=method, normalization=norm, linear_weights=weights,
                     percentile_distributions=distrs)
        for qi in range(n_q):
            ids, sc = ofusion.fuse_query([np.array([x["corpus_id"] for x in lists[s][qi]]) for s in systems],
                                         [np.array([x["score"] for x in lists[s][qi]], dtype=np.float64) for s in systems],
                                         method, norm, [weights[s] for s in systems], [distrs[s] for s in systems])
            assert ids == [x["corpus_id"] for x in res[qi]], (seed, method, norm, qi)
            got = np.array([float(v) for v in sc])
            want = np.array([float(x["score"]) for x in res[qi]])
            if np.isnan(want).any():            # z-score of a single-entry list: torch.std -> nan, like the reference
                assert np.array_equal(np.isnan(got), np.isnan(want))
            else:
                assert np.array_equal(got, want), (seed, method, norm, qi)


@pytest.mark.parametrize("seed,variant,k1,b", [(1, "bm25", 0.9, 0.4), (2, "bm25", 2.5, 0.2), (3, "tfidf", 0, 0), (4, "atire", 1.2, 0.75)])
def test_lexical_oracle_equals_reference_on_random_corpora(seed, variant, k1, b):
    from oracle import bm25 as obm25
    mod = ref_loader.load_bm25()
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    vocab = [f"w{j}" for j in range(40)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(1, 30)))) for _ in range(300)]
    docs[7] = docs[3]                                                        # duplicate document: an exact tie
    queries = [" ".join(rng.choice(vocab + ["zzz"], size=int(rng.integers(1, 7)))) for _ in range(10)] + ["", "zzz zzz"]
    cls = {"bm25": mod.BM25, "tfidf": mod.TFIDF, "atire": mod.AtireBM25}[variant]
    r = _quiet(cls, docs) if variant == "tfidf" else _quiet(cls, docs, k1=k1, b=b)
    res = _quiet(r.search_all, queries, top_k=len(docs))
    o = obm25.LexicalOracle.from_strings(docs, variant, k1, b)
    for qi, q in enumerate(queries):
        ids, sc = o.search_ids(o.query_ids(q), len(docs))
        assert ids.tolist() == [x["corpus_id"] for x in res[qi]], (variant, qi)
        assert sc.tolist() == [x["score"] for x in res[qi]], (variant, qi)      # bit-exact fp64


@pytest.mark.parametrize("pooling,keep", [("max", None), ("sum", 16), ("max", 1)])
def test_splade_head_oracle_equals_reference_on_random_logits(pooling, keep):
    from oracle import splade_head as oh
    g = torch.Generator().manual_seed(7)
    logits = torch.randn((5, 11, 257), generator=g) * 3
    mask = (torch.rand((5, 11), generator=g) > 0.3).long()
    m = ref_loader.make_splade_head(logits, pooling, keep)
    want = m.forward(None, mask)
    got = oh.pool(logits, mask, pooling)
    if keep is not None:
        got, _ = oh.prune(got, keep)
    assert torch.equal(got, want)


@pytest.mark.parametrize("seed", range(3))
def test_metrics_oracle_equals_reference_on_random_rankings(seed):
    import importlib
    from oracle import metrics as ometrics
    ref_loader.load_hybrid()
    metrics_mod = importlib.import_module("src.utils.metrics")
    rng = np.random.Generator(np.random.PCG64(3000 + seed))
    n_q, pool = 25, 300
    results = [rng.permutation(pool)[: int(rng.integers(1, 200))].tolist() for _ in range(n_q)]
    golds = [rng.choice(pool + 20, size=int(rng.integers(1, 8)), replace=False).tolist() for _ in range(n_q)]
    ks = dict(recall_ks=(5, 10, 100, 1000), map_ks=(10, 100), mrr_ks=(10, 100), ndcg_ks=(10, 100))
    ev = metrics_mod.Metrics(recall_at_k=list(ks["recall_ks"]), map_at_k=list(ks["map_ks"]), mrr_at_k=list(ks["mrr_ks"]),
                             ndcg_at_k=list(ks["ndcg_ks"]))
    want = _quiet(ev.compute_all_metrics, all_ground_truths=golds, all_results=results)
    got = dict(zip(ometrics.metric_names(**ks), ometrics.mean_metrics(golds, results, **ks)))
    assert set(got) == set(want)
    for name in got:
        assert got[name] == pytest.approx(float(want[name]), abs=1e-12), name


@pytest.mark.parametrize("sim", ["cos_sim", "dot"])
def test_dense_oracle_equals_reference_search(sim):
    """oracle/dense.py::semantic_search against the verbatim BaseModel.search (splade/base.py:199-251) with injected
    embeddings, several query / doc chunks, an exact duplicate row."""
    from oracle import dense as odense
    g = torch.Generator().manual_seed(13)
    q, d = torch.randn((9, 48), generator=g), torch.randn((2500, 48), generator=g)
    d[100] = d[40]
    m = ref_loader.make_injected_searcher(sim, q, d)
    res = _quiet(m.search, ["x"] * 9, ["y"] * 2500, query_chunk_size=4, doc_chunk_size=700, topk=30)
    mine = odense.semantic_search(q, d, 30, sim, query_chunk_size=4, corpus_chunk_size=700, key="doc_id")
    for qi in range(9):
        want_s = [x["score"] for x in res[qi]]
        got_s = [x["score"] for x in mine[qi]]
        assert got_s == want_s
        # ids: equal outside exact-tie groups (the reference's order inside a tie group is arbitrary)
        want_i, got_i = [x["doc_id"] for x in res[qi]], [x["doc_id"] for x in mine[qi]]
        for s in set(want_s):
            assert {i for i, v in zip(want_i, want_s) if v == s} == {i for i, v in zip(got_i, got_s) if v == s}
