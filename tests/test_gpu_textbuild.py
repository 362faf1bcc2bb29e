"""GPU parity: device tokenisation (str.split() rules) and percentile distributions (8f-3)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from text_cases import DOCS, QUERIES  # noqa: E402  (strings with every str.isspace() character)


def test_tokenizer_matches_str_split():
    from fusion_b200 import text
    vocab, ptr, tok = text.tokenize_corpus(DOCS, "cuda")
    ptr, tok = ptr.cpu().numpy(), tok.cpu().numpy()
    expect = [d.split() for d in DOCS]
    assert np.diff(ptr).tolist() == [len(e) for e in expect]
    word_of, id_of = {}, {}
    for di, words in enumerate(expect):
        for w, t in zip(words, tok[ptr[di]:ptr[di + 1]].tolist()):
            assert word_of.setdefault(t, w) == w            # one id -> one word
            assert id_of.setdefault(w, t) == t              # one word -> one id
    assert len(vocab) == len(id_of)
    queries = QUERIES
    qptr, qtok = text.tokenize_queries(queries, vocab, "cuda")
    qptr, qtok = qptr.cpu().numpy(), qtok.cpu().numpy()
    for qi, q in enumerate(queries):
        got = qtok[qptr[qi]:qptr[qi + 1]].tolist()
        assert got == [id_of.get(w, -1) for w in q.split()]


@pytest.mark.parametrize("cls_name,kw", [("BM25", dict(k1=2.5, b=0.2)), ("TFIDF", {})])
def test_device_tokenizer_gives_reference_rankings(golden_dir, cls_name, kw):
    """Same golden as the host-tokenised path (verbatim reference output): ids and fp64 scores bit-exact."""
    from fusion_b200.retrievers import bm25 as mod
    g = np.load(os.path.join(golden_dir, "lexical_small.npz"))
    docs, queries = [str(x) for x in g["docs"]], [str(x) for x in g["queries"]]
    tag = "bm25" if cls_name == "BM25" else "tfidf"
    r = getattr(mod, cls_name)(docs, device_tokenizer=True, tile_docs=256, tiled_min=8, dense_frac=0.3, **kw)
    sc, ids = r.search_all_tensors(queries, top_k=100)
    assert np.array_equal(ids.cpu().numpy(), g[f"{tag}_ids"][:, :100])
    assert np.array_equal(sc.cpu().numpy(), g[f"{tag}_scores"][:, :100])


def test_large_random_corpus_token_counts():
    from fusion_b200 import synth, text
    (dptr, dtok), _ = synth.c3_lexical(20000, 4, 5000)
    docs = synth.ids_to_strings(dptr, dtok)
    vocab, ptr, tok = text.tokenize_corpus(docs, "cuda")
    assert torch.equal(ptr.cpu(), torch.from_numpy(dptr))
    # same partition of the token occurrences as the generator's term ids
    a, b = tok.cpu().numpy().astype(np.int64), dtok.astype(np.int64)
    assert len(np.unique(a)) == len(np.unique(b)) == len(np.unique(a * 10 ** 6 + b))


def test_percentile_distribution_matches_pandas(golden_dir):
    """hybrid.py:391-398 (pandas, executed when the golden was made) vs sort + fz_quantiles_f64."""
    from fusion_b200 import text
    from oracle import distributions as od
    g = np.load(os.path.join(golden_dir, "distribution_small.npz"))
    for s in ("bm25", "dpr"):
        sc = torch.from_numpy(g[f"scores_{s}"]).cuda()
        for n in (10, 1000):
            got = text.percentile_distribution(sc, n).cpu().numpy()
            np.testing.assert_allclose(got, g[f"distr_{s}_{n}"], rtol=1e-12, atol=1e-12)
            assert np.array_equal(od.percentile_distribution(g[f"scores_{s}"], n), g[f"distr_{s}_{n}"])
    # the distribution feeds percentile-rank fusion: ascending, as fz_fuse requires
    d = text.percentile_distribution(torch.from_numpy(g["scores_bm25"]).cuda(), 100)
    assert bool((d[1:] >= d[:-1]).all())
