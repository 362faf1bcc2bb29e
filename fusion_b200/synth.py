"""Seeded synthetic corpora, queries and embeddings for the retrieval hot path.

The shapes and seeds follow SURVEY.md section 8(d).  Every generator uses
``numpy.random.Generator(PCG64(seed))`` so the oracle, the parity tests and the
benchmark see identical inputs.  Nothing here touches a GPU; the full-size
(8.8M passage) device-side generators live in ``bench.py``.

Reference scale constants: ``src/data/mmarco.py:1-9`` (8.8M passages, 6,980 dev
queries), ``src/retrievers/hybrid.py:346`` (LLeQA BM25 k1=2.5, b=0.2),
``scripts/run_bm25.sh:22-28`` (mMARCO k1=0.9, b=0.4).
"""
from __future__ import annotations

import numpy as np

# seeds are part of the measurement contract (SURVEY.md 8d / BASELINE.md)
SEED_C1_DOCS, SEED_C1_QUERIES, SEED_C1_DEMB, SEED_C1_QEMB = 101, 102, 103, 104
SEED_C2_DEMB, SEED_C2_QEMB = 201, 202
SEED_C3_DOCS, SEED_C3_QUERIES, SEED_C3_SPL_D, SEED_C3_SPL_Q = 301, 302, 311, 312
SEED_C4_TOK, SEED_C4_Q = 401, 402


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def zipf_ids(rng: np.random.Generator, n: int, vocab: int, s: float) -> np.ndarray:
    """Draw ``n`` term ids from a Zipf(s) law truncated to ``[0, vocab)`` (inverse-CDF sampling)."""
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-s))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(n), side="left").astype(np.int32)


def lexical_corpus(n_docs: int, vocab: int, zipf_s: float, len_mu: float, len_sigma: float,
                   len_min: int, len_max: int, seed: int):
    """Token-id corpus: returns (doc_ptr[int64, n_docs+1], tokens[int32])."""
    rng = _rng(seed)
    lens = np.clip(rng.lognormal(len_mu, len_sigma, n_docs), len_min, len_max).astype(np.int64)
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    toks = zipf_ids(rng, int(ptr[-1]), vocab, zipf_s)
    return ptr, toks


def lexical_queries(n_queries: int, vocab: int, zipf_s: float, mean_extra: float, seed: int,
                    oov_every: int = 0):
    """Token-id queries of length 1+Poisson(mean_extra).  ``oov_every`` > 0 replaces the first token
    of every ``oov_every``-th query by an out-of-vocabulary id (== ``vocab``)."""
    rng = _rng(seed)
    lens = 1 + rng.poisson(mean_extra, n_queries).astype(np.int64)
    ptr = np.zeros(n_queries + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    toks = zipf_ids(rng, int(ptr[-1]), vocab, zipf_s)
    if oov_every:
        toks[ptr[:-1][::oov_every]] = vocab
    return ptr, toks


def ids_to_strings(ptr: np.ndarray, toks: np.ndarray) -> list[str]:
    """Render token ids as the whitespace-separated strings the reference's BM25 consumes."""
    words = np.char.add("t", toks.astype(np.int64).astype(str))
    return [" ".join(words[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


def c1_lexical(n_docs: int = 27942, n_queries: int = 200, vocab: int = 50000):
    """C1 (LLeQA-shaped): Zipf(1.1), doc length clip(lognormal(4.5, 0.8), 5, 2000), query 1+Poisson(7)."""
    dptr, dtok = lexical_corpus(n_docs, vocab, 1.1, 4.5, 0.8, 5, 2000, SEED_C1_DOCS)
    qptr, qtok = lexical_queries(n_queries, vocab, 1.1, 7.0, SEED_C1_QUERIES, oov_every=17)
    return (dptr, dtok), (qptr, qtok)


def c3_lexical(n_docs: int, n_queries: int, vocab: int = 500000):
    """C3 (mMARCO-shaped BM25): Zipf(1.07), doc length clip(lognormal(3.3, 0.5), 3, 256), query 1+Poisson(4)."""
    dptr, dtok = lexical_corpus(n_docs, vocab, 1.07, 3.3, 0.5, 3, 256, SEED_C3_DOCS)
    qptr, qtok = lexical_queries(n_queries, vocab, 1.07, 4.0, SEED_C3_QUERIES, oov_every=29)
    return (dptr, dtok), (qptr, qtok)


def dense_embeddings(n: int, d: int, seed: int, normalize: bool = False) -> np.ndarray:
    x = _rng(seed).standard_normal((n, d), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def splade_vectors(n: int, vocab: int, mean_nnz: float, nnz_min: int, nnz_max: int, seed: int,
                   zipf_s: float = 1.05):
    """Sparse term-weight vectors shaped like SPLADE output (``splade.py:88-99``):
    weights ``log1p(relu(N(0.5, 0.7)))`` with zeros dropped, term ids Zipf(1.05) de-duplicated per vector.
    Returns CSR (ptr[int64], term[int32] ascending per row, weight[float32])."""
    rng = _rng(seed)
    want = np.clip(rng.poisson(mean_nnz, n), nnz_min, nnz_max).astype(np.int64)
    over = (want * 2 + 8)  # oversample, then unique: Zipf draws collide often
    optr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(over, out=optr[1:])
    draws = zipf_ids(rng, int(optr[-1]), vocab, zipf_s)
    row = np.repeat(np.arange(n, dtype=np.int64), over)
    key = np.unique(row * vocab + draws)            # sorted (row, term), duplicates removed
    urow, uterm = key // vocab, (key % vocab).astype(np.int32)
    cnt = np.bincount(urow, minlength=n)
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    # keep a random subset of size min(want, cnt) per row
    prio = rng.random(len(key))
    order = np.lexsort((prio, urow))
    rank = np.arange(len(key)) - start[urow[order]]
    keep_sorted = rank < want[urow[order]]
    keep = np.zeros(len(key), dtype=bool)
    keep[order] = keep_sorted
    w = np.log1p(np.maximum(rng.normal(0.5, 0.7, len(key)), 0.0)).astype(np.float32)
    keep &= w > 0
    urow, uterm, w = urow[keep], uterm[keep], w[keep]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(urow, minlength=n), out=ptr[1:])
    return ptr, uterm, w


def densify(ptr: np.ndarray, term: np.ndarray, w: np.ndarray, vocab: int) -> np.ndarray:
    out = np.zeros((len(ptr) - 1, vocab), dtype=np.float32)
    row = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    out[row, term] = w
    return out


def colbert_tokens(n_docs: int, dim: int, mean_len: float, len_min: int, len_max: int, seed: int):
    """Token-embedding store: returns (tok_ptr[int64, n_docs+1], tok_emb[float32, T, dim] unit rows)."""
    rng = _rng(seed)
    lens = np.clip(rng.poisson(mean_len, n_docs), len_min, len_max).astype(np.int64)
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    emb = rng.standard_normal((int(ptr[-1]), dim), dtype=np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    return ptr, emb


def colbert_queries(n_queries: int, lq: int, dim: int, seed: int) -> np.ndarray:
    q = _rng(seed).standard_normal((n_queries, lq, dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q
