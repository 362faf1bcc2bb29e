"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Each fixture stores its inputs next to the reference's outputs, so the GPU box (which has no
/root/reference) can replay them.  Reference entry points executed verbatim:
  * ``src/retrievers/bm25.py``  TFIDF / BM25 / AtireBM25 ``.search_all``          (bm25.py:33-173)
  * ``src/retrievers/hybrid.py`` ``Aggregator.fuse``                              (hybrid.py:166-307)
  * ``src/retrievers/splade/base.py`` ``BaseModel.search`` with injected embeddings (base.py:199-251)
"""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np
import torch

from fusion_b200 import synth
from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        return fn(*a, **k)


def golden_lexical():
    mod = ref_loader.load_bm25()
    # --- small adversarial corpus: negative idf (df > N/2), OOV, repeated tokens, < k matches
    rng = np.random.Generator(np.random.PCG64(7))
    n_docs, vocab = 600, 800
    dptr, dtok = synth.lexical_corpus(n_docs, vocab, 1.1, 3.0, 0.6, 3, 120, seed=11)
    # force term 0 into 80% of the docs (negative BM25 idf) and make two docs identical (tie on every query)
    first = dptr[:-1][rng.random(n_docs) < 0.8]
    dtok[first] = 0
    l0, l1 = dptr[5], dptr[6]
    n_copy = min(dptr[6] - dptr[5], dptr[8] - dptr[7])
    qptr, qtok = synth.lexical_queries(24, vocab, 1.1, 4.0, seed=12, oov_every=5)
    docs = synth.ids_to_strings(dptr, dtok)
    docs[7] = docs[5]                       # exact duplicate document => exact score ties
    queries = synth.ids_to_strings(qptr, qtok)
    queries[1] = "t0"                       # only the negative-idf term
    queries[2] = "t3 t3 t9 t3"              # repeated tokens
    queries[3] = "zzz yyy"                  # nothing matches: all scores 0 -> doc order
    queries[4] = "t799 t0"                  # rare + negative
    out = {"docs": np.array(docs), "queries": np.array(queries)}
    for name, cls, kw in (("tfidf", mod.TFIDF, {}), ("bm25", mod.BM25, dict(k1=2.5, b=0.2)),
                          ("bm25_mm", mod.BM25, dict(k1=0.9, b=0.4)), ("atire", mod.AtireBM25, dict(k1=0.9, b=0.4))):
        r = _quiet(cls, docs, **kw)
        res = _quiet(r.search_all, queries, top_k=n_docs)
        out[f"{name}_ids"] = np.array([[x["corpus_id"] for x in q] for q in res], dtype=np.int32)
        out[f"{name}_scores"] = np.array([[x["score"] for x in q] for q in res], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "lexical_small.npz"), **out)

    # --- C1-shaped slice: 4,000 docs of the LLeQA-shaped generator, 16 queries, top-200, k1=2.5 b=0.2
    (dptr, dtok), (qptr, qtok) = synth.c1_lexical(n_docs=4000, n_queries=16)
    docs, queries = synth.ids_to_strings(dptr, dtok), synth.ids_to_strings(qptr, qtok)
    r = _quiet(mod.BM25, docs, k1=2.5, b=0.2)
    res = _quiet(r.search_all, queries, top_k=200)
    np.savez_compressed(
        os.path.join(OUT, "lexical_c1_slice.npz"), n_docs=4000, n_queries=16, top_k=200,
        ids=np.array([[x["corpus_id"] for x in q] for q in res], dtype=np.int32),
        scores=np.array([[x["score"] for x in q] for q in res], dtype=np.float64))


def _ranked_lists(rng, n_q, n, pool, dup=False, const=False):
    out = []
    for _ in range(n_q):
        ids = rng.choice(pool, size=n, replace=False)
        sc = np.sort(rng.normal(0, 3, n))[::-1]
        if const:
            sc = np.full(n, 1.25)
        if dup:
            ids[n // 2] = ids[1]          # repeated id inside one list
        out.append([{"corpus_id": int(i), "score": float(s)} for i, s in zip(ids, sc)])
    return out


def golden_fusion():
    mod = ref_loader.load_hybrid()
    rng = np.random.Generator(np.random.PCG64(21))
    n_q, n, pool = 6, 60, 150
    lists = {"bm25": _ranked_lists(rng, n_q, n, pool, dup=True),
             "dpr": _ranked_lists(rng, n_q, n - 7, pool),
             "splade": _ranked_lists(rng, n_q, n, pool, const=True)}
    # fp32-exact dense-like scores for one system (as produced by .tolist() of an fp32 tensor)
    for q in lists["dpr"]:
        for x in q:
            x["score"] = float(np.float32(x["score"] * 0.01))
    weights = {"bm25": 0.5, "dpr": 0.3, "splade": 0.2}
    distrs = {k: np.quantile(rng.normal(0, 3, 5000), np.linspace(0, 1, 101)) for k in lists}
    out = {"systems": np.array(list(lists)), "weights": np.array([weights[k] for k in lists])}
    for k, v in lists.items():
        L = max(len(q) for q in v)
        out[f"in_ids_{k}"] = np.array([[x["corpus_id"] for x in q] for q in v], dtype=np.int32)
        out[f"in_scores_{k}"] = np.array([[x["score"] for x in q] for q in v], dtype=np.float64)
        out[f"distr_{k}"] = distrs[k]
    import copy
    cases = [("bcf", None), ("rrf", None)] + [("nsf", nm) for nm in
             ("none", "min-max", "z-score", "arctan", "percentile-rank", "normal-curve-equivalent")]
    for method, norm in cases:
        res = mod.Aggregator.fuse(copy.deepcopy(lists), method=method, normalization=norm,
                                  linear_weights=weights, percentile_distributions=distrs)
        tag = method if norm is None else f"{method}_{norm}"
        L = max(len(q) for q in res)
        ids = np.full((n_q, L), -1, dtype=np.int32)
        sc = np.full((n_q, L), np.nan, dtype=np.float64)
        for qi, q in enumerate(res):
            ids[qi, :len(q)] = [x["corpus_id"] for x in q]
            sc[qi, :len(q)] = [float(x["score"]) for x in q]
        out[f"out_ids_{tag}"], out[f"out_scores_{tag}"] = ids, sc
        out[f"out_dtype_{tag}"] = np.array(type(res[0][0]["score"]).__name__)
    np.savez_compressed(os.path.join(OUT, "fusion_small.npz"), **out)


def golden_sweep():
    """Weight sweep + metrics, verbatim: Aggregator.fuse per weight vector (hybrid.py:404-426) and Metrics with
    run_evaluation's configuration (hybrid.py:24-27, utils/metrics.py)."""
    import copy
    import importlib
    import itertools
    mod = ref_loader.load_hybrid()
    metrics_mod = importlib.import_module("src.utils.metrics")
    rng = np.random.Generator(np.random.PCG64(23))
    n_q, n, pool = 8, 40, 90
    lists = {"bm25": _ranked_lists(rng, n_q, n, pool, dup=True), "dpr": _ranked_lists(rng, n_q, n - 5, pool),
             "splade": _ranked_lists(rng, n_q, n, pool)}
    for q in lists["dpr"]:
        for x in q:
            x["score"] = float(np.float32(x["score"] * 0.01))
    golds = [sorted(rng.choice(pool, size=int(rng.integers(1, 6)), replace=False).tolist()) for _ in range(n_q)]
    golds[2] = [pool + 5]                  # a query whose only relevant doc is never retrieved
    step = 0.25
    combos = [comb for comb in itertools.product(np.arange(0, 1 + step, step), repeat=len(lists)) if np.isclose(sum(comb), 1.0)]
    evaluator = metrics_mod.Metrics(recall_at_k=[5, 10, 20, 50, 100, 200, 500, 1000], map_at_k=[10, 100], mrr_at_k=[10, 100],
                                    ndcg_at_k=[10, 100])
    out = {"systems": np.array(list(lists)), "weights": np.array(combos, dtype=np.float64),
           "gold_ptr": np.cumsum([0] + [len(g) for g in golds]).astype(np.int32),
           "gold_ids": np.concatenate(golds).astype(np.int32)}
    for k, v in lists.items():
        out[f"in_ids_{k}"] = np.array([[x["corpus_id"] for x in q] for q in v], dtype=np.int32)
        out[f"in_scores_{k}"] = np.array([[x["score"] for x in q] for q in v], dtype=np.float64)
    names = None
    for norm in ("min-max", "z-score", "none"):
        rows = []
        for comb in combos:
            weights = {name: w for name, w in zip(lists, comb)}
            ranked = mod.Aggregator.fuse(copy.deepcopy(lists), method="nsf", normalization=norm, linear_weights=weights,
                                         percentile_distributions={})
            sc = evaluator.compute_all_metrics(all_ground_truths=golds, all_results=[[x["corpus_id"] for x in r] for r in ranked])
            names = list(sc)
            rows.append([float(sc[m]) for m in names])
        out[f"metrics_{norm}"] = np.array(rows, dtype=np.float64)
    out["metric_names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "sweep_small.npz"), **out)


def golden_distribution():
    """Percentile-based score distributions, the pandas expression of hybrid.py:391-398 executed as written."""
    import pandas as pd
    rng = np.random.Generator(np.random.PCG64(29))
    rows = []
    for system, (mu, sd, n) in {"bm25": (4.0, 2.5, 4000), "dpr": (0.3, 0.1, 3000)}.items():
        sc = rng.normal(mu, sd, n)
        sc[rng.random(n) < 0.2] = 0.0                      # unmatched documents
        sc[:50] = sc.min()                                 # a repeated smallest value
        sc = np.float32(sc).astype(np.float64) if system == "dpr" else sc
        rows += [{"system": system, "score": float(v)} for v in sc]
    all_scores_df = pd.DataFrame(rows, columns=["system", "score"])
    out = {"systems": np.array(["bm25", "dpr"])}
    for s in ("bm25", "dpr"):
        out[f"scores_{s}"] = all_scores_df[all_scores_df["system"] == s]["score"].to_numpy()
    # hybrid.py:391-398 chains two groupby().apply() calls; under the pandas installed here (>= 2.2) apply() no longer
    # carries the grouping column through reset_index(drop=True), so the same two steps run per group explicitly with the
    # same pandas calls (filter: drop zeros and the two smallest distinct scores; Series.quantile(np.linspace(0, 1, N+1))).
    for N in (10, 1000):
        for s, group in all_scores_df.groupby('system'):
            kept = group[(group['score'] != 0.0) & (~group['score'].isin(group['score'].drop_duplicates().nsmallest(2)))]
            out[f"distr_{s}_{N}"] = pd.Series(kept['score'].quantile(np.linspace(0, 1, N+1))).to_numpy()
    np.savez_compressed(os.path.join(OUT, "distribution_small.npz"), **out)


def golden_dense():
    q = torch.from_numpy(synth.dense_embeddings(7, 64, seed=31))
    d = torch.from_numpy(synth.dense_embeddings(3000, 64, seed=32))
    d[11] = d[10]                          # exact tie
    out = {"q": q.numpy(), "d": d.numpy()}
    for sim in ("cos_sim", "dot"):
        m = ref_loader.make_injected_searcher(sim, q, d)
        res = m.search(["x"] * 7, ["y"] * 3000, query_chunk_size=3, doc_chunk_size=1000, topk=50)
        out[f"{sim}_ids"] = np.array([[x["doc_id"] for x in r] for r in res], dtype=np.int32)
        out[f"{sim}_scores"] = np.array([[x["score"] for x in r] for r in res], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "dense_small.npz"), **out)


def splade_head_inputs():
    """Seeded logits [6, 9, 517] with the cases the head must get right: padded tails, a fully padded row, one
    unmasked position, all-negative logits, exact duplicates across the pruning cutoff."""
    g = torch.Generator().manual_seed(77)
    logits = torch.randn((6, 9, 517), generator=g) * 2.0
    mask = torch.ones((6, 9), dtype=torch.int64)
    mask[0, 5:] = 0
    mask[1, :] = 0
    mask[2, 1:] = 0
    logits[3] = -logits[3].abs()
    logits[4, :, 100:140] = logits[4, :, 60:100]          # duplicated columns: equal activations
    return logits, mask


def golden_splade_head():
    logits, mask = splade_head_inputs()
    out = {"logits": logits.numpy(), "mask": mask.numpy()}
    for pooling in ("max", "sum"):
        m = ref_loader.make_splade_head(logits, pooling, None)
        act = m.forward(None, mask)
        out[f"act_{pooling}"] = act.numpy()
        for k in (1, 32, 517):
            pruned, idx = m._prune_activations(act, keep_topk=k)
            out[f"pruned_{pooling}_{k}"] = pruned.numpy()
            out[f"topk_{pooling}_{k}"] = idx.numpy()
        m2 = ref_loader.make_splade_head(logits, pooling, 32)
        assert torch.equal(m2.forward(None, mask), torch.from_numpy(out[f"pruned_{pooling}_32"]))
    np.savez_compressed(os.path.join(OUT, "splade_head_small.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    golden_splade_head()
    golden_lexical()
    golden_fusion()
    golden_sweep()
    golden_distribution()
    golden_dense()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
