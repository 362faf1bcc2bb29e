"""The oracle restatements against fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from fusion_b200 import synth
from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import fusion as ofusion


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("tag,variant,k1,b", [("tfidf", "tfidf", 0, 0), ("bm25", "bm25", 2.5, 0.2),
                                              ("bm25_mm", "bm25", 0.9, 0.4), ("atire", "atire", 0.9, 0.4)])
def test_lexical_small_bit_exact(golden_dir, tag, variant, k1, b):
    g = _load(golden_dir, "lexical_small.npz")
    docs, queries = [str(x) for x in g["docs"]], [str(x) for x in g["queries"]]
    o = obm25.LexicalOracle.from_strings(docs, variant, k1, b)
    for qi, q in enumerate(queries):
        ids, sc = o.search_ids(o.query_ids(q), len(docs))
        assert np.array_equal(ids, g[f"{tag}_ids"][qi]), (tag, qi)
        assert np.array_equal(sc, g[f"{tag}_scores"][qi]), (tag, qi)      # bit-exact fp64


def test_lexical_c1_slice_bit_exact(golden_dir):
    g = _load(golden_dir, "lexical_c1_slice.npz")
    (dptr, dtok), (qptr, qtok) = synth.c1_lexical(n_docs=int(g["n_docs"]), n_queries=int(g["n_queries"]))
    docs, queries = synth.ids_to_strings(dptr, dtok), synth.ids_to_strings(qptr, qtok)
    o = obm25.LexicalOracle.from_strings(docs, "bm25", 2.5, 0.2)
    for qi, q in enumerate(queries):
        ids, sc = o.search_ids(o.query_ids(q), int(g["top_k"]))
        assert np.array_equal(ids, g["ids"][qi])
        assert np.array_equal(sc, g["scores"][qi])


def test_lexical_token_id_path_matches_string_path():
    (dptr, dtok), (qptr, qtok) = synth.c1_lexical(n_docs=500, n_queries=8, vocab=300)
    docs = synth.ids_to_strings(dptr, dtok)
    a = obm25.LexicalOracle.from_strings(docs, "bm25", 0.9, 0.4)
    b = obm25.LexicalOracle(dptr, dtok, 300, "bm25", 0.9, 0.4)
    for qi in range(8):
        toks = qtok[qptr[qi]:qptr[qi + 1]]
        q = " ".join(f"t{t}" for t in toks)
        ia, sa = a.search_ids(a.query_ids(q), 500)
        ib, sb = b.search_ids(np.where(toks < 300, toks, -1), 500)
        assert np.array_equal(ia, ib) and np.array_equal(sa, sb)


FUSION_CASES = [("bcf", None), ("rrf", None)] + [("nsf", n) for n in ofusion.NORMALIZATIONS]


@pytest.mark.parametrize("method,norm", FUSION_CASES)
def test_fusion_small(golden_dir, method, norm):
    g = _load(golden_dir, "fusion_small.npz")
    systems = [str(s) for s in g["systems"]]
    tag = method if norm is None else f"{method}_{norm}"
    exp_ids, exp_sc = g[f"out_ids_{tag}"], g[f"out_scores_{tag}"]
    for qi in range(exp_ids.shape[0]):
        ids, sc = ofusion.fuse_query([g[f"in_ids_{s}"][qi] for s in systems], [g[f"in_scores_{s}"][qi] for s in systems],
                                     method, norm, list(g["weights"]), [g[f"distr_{s}"] for s in systems])
        n = int((exp_ids[qi] >= 0).sum())
        assert ids == exp_ids[qi, :n].tolist(), (tag, qi)
        np.testing.assert_allclose(np.asarray(sc, dtype=np.float64), exp_sc[qi, :n], rtol=0, atol=0, equal_nan=True)


@pytest.mark.parametrize("sim", ["cos_sim", "dot"])
def test_dense_small(golden_dir, sim):
    g = _load(golden_dir, "dense_small.npz")
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"])
    res = odense.semantic_search(q, d, 50, sim=sim, query_chunk_size=3, corpus_chunk_size=1000, key="doc_id")
    sc, ids = odense.topk_tensors(q, d, 50, sim=sim, chunk=700)
    for qi in range(7):
        got_ids = [x["doc_id"] for x in res[qi]]
        got_sc = np.array([x["score"] for x in res[qi]])
        np.testing.assert_array_equal(got_sc, g[f"{sim}_scores"][qi])
        assert sorted(got_ids) == sorted(g[f"{sim}_ids"][qi].tolist())
        # tensor form: same scores to fp32 rounding, same id set
        np.testing.assert_allclose(sc[qi].numpy(), g[f"{sim}_scores"][qi], rtol=2e-6, atol=2e-6)
        assert sorted(ids[qi].tolist()) == sorted(g[f"{sim}_ids"][qi].tolist())


def test_metrics_and_sweep_oracle_match_verbatim_reference(golden_dir):
    """oracle/metrics.py against Aggregator.fuse + Metrics executed verbatim for every weight vector (hybrid.py:404-426)."""
    import os
    import numpy as np
    from oracle import metrics as om
    g = np.load(os.path.join(golden_dir, "sweep_small.npz"))
    systems = [str(s) for s in g["systems"]]
    nq = g[f"in_ids_{systems[0]}"].shape[0]
    ids = [[g[f"in_ids_{k}"][q] for q in range(nq)] for k in systems]
    sc = [[g[f"in_scores_{k}"][q] for q in range(nq)] for k in systems]
    gp, gi = g["gold_ptr"], g["gold_ids"]
    golds = [gi[gp[i]:gp[i + 1]].tolist() for i in range(nq)]
    assert [str(x) for x in g["metric_names"]] == om.metric_names()
    assert [tuple(w) for w in g["weights"]] == om.weight_grid(len(systems), 0.25)
    for norm in ("min-max", "z-score", "none"):
        got = np.array(om.sweep(ids, sc, golds, [tuple(w) for w in g["weights"]], norm))
        assert np.array_equal(got, g[f"metrics_{norm}"]), norm


def test_distribution_oracle_matches_pandas_golden(golden_dir):
    import os
    import numpy as np
    from oracle import distributions as od
    g = np.load(os.path.join(golden_dir, "distribution_small.npz"))
    for s in ("bm25", "dpr"):
        for n in (10, 1000):
            assert np.array_equal(od.percentile_distribution(g[f"scores_{s}"], n), g[f"distr_{s}_{n}"])


@pytest.mark.parametrize("pooling", ["max", "sum"])
def test_splade_head_matches_reference(golden_dir, pooling):
    """oracle/splade_head.py against the verbatim SPLADE.forward / _prune_activations outputs (stub encoder)."""
    from oracle import splade_head as oh
    g = _load(golden_dir, "splade_head_small.npz")
    logits, mask = torch.from_numpy(g["logits"]), torch.from_numpy(g["mask"])
    act = oh.pool(logits, mask, pooling)
    assert np.array_equal(act.numpy(), g[f"act_{pooling}"])
    for k in (1, 32, 517):
        pruned, idx = oh.prune(act, k)
        assert np.array_equal(pruned.numpy(), g[f"pruned_{pooling}_{k}"])
        assert np.array_equal(idx.numpy(), g[f"topk_{pooling}_{k}"])
    ptr, term, w = oh.to_csr(act)
    dense = np.zeros_like(g[f"act_{pooling}"])
    dense[np.repeat(np.arange(len(ptr) - 1), np.diff(ptr)), term] = w
    assert np.array_equal(dense, g[f"act_{pooling}"]) and ptr[-1] == np.count_nonzero(dense)
    assert all(np.all(np.diff(term[ptr[i]:ptr[i + 1]]) > 0) for i in range(len(ptr) - 1))


def test_maxsim_oracle_matches_independent_padded_fixture(golden_dir):
    """colbert-ai is third-party and absent: ``oracle/maxsim.py`` (ragged per-pair loop) is pinned against the committed
    fixture of ``oracle/make_golden_maxsim.py`` - colbert_score's padded-batch formulation (D_padded @ Q^T, padding rows
    -> -9999, max over doc tokens, sum over query tokens) in float64, incl. an empty and a 1-token passage."""
    import torch
    from oracle import maxsim
    g = np.load(os.path.join(golden_dir, "maxsim_small.npz"))
    got = maxsim.maxsim_scores(torch.from_numpy(g["q"]), torch.from_numpy(g["tok_ptr"]), torch.from_numpy(g["tok_emb"]),
                               torch.from_numpy(g["cand"])).numpy()
    np.testing.assert_allclose(got, g["scores"], rtol=1e-6, atol=2e-6)


@pytest.mark.parametrize("norm", ["min-max", "z-score", "arctan", "percentile-rank"])
def test_fusion_oracle_numpy1_promotion(golden_dir, norm):
    """NumPy 1.x promotion (fp32 normalised score * float weight -> float64, summed in float64): the oracle's
    ``promote_f64`` mode against the verbatim reference run with np.float64 weights (oracle/make_golden_promotion.py)."""
    g, leg = _load(golden_dir, "fusion_small.npz"), _load(golden_dir, "fusion_legacy.npz")
    systems = [str(s) for s in g["systems"]]
    for qi in range(leg[f"out_ids_{norm}"].shape[0]):
        ids, sc = ofusion.fuse_query([g[f"in_ids_{s}"][qi] for s in systems], [g[f"in_scores_{s}"][qi] for s in systems], "nsf",
                                     norm, list(g["weights"]), [g[f"distr_{s}"] for s in systems], promote_f64=True)
        n = int((leg[f"out_ids_{norm}"][qi] >= 0).sum())
        assert ids == leg[f"out_ids_{norm}"][qi, :n].tolist()
        assert np.array_equal(np.array(sc, dtype=np.float64), leg[f"out_scores_{norm}"][qi, :n])
