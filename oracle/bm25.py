"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the reference lexical scorers.

Follows ``src/retrievers/bm25.py``:
  * index build            TFIDF.__init__/_build_vocab/_build_tf_index/_build_df_index   :37-83
  * idf                    TFIDF._compute_idf :85-87, BM25._compute_idf :145-147, AtireBM25 :171-173
  * doc lengths / avgdl    BM25.__init__/_build_dl_index :133-143 (``statistics.mean``)
  * scoring                TFIDF.score :108-115, BM25.score :149-156 (fp64, query-token order, duplicates kept)
  * ranking                TFIDF.search :100-106 (every doc scored, stable sort descending => ties by lower index)

Pinned bit-for-bit against the verbatim classes by ``oracle/make_golden.py`` / ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

VARIANTS = ("tfidf", "bm25", "atire")


def tokenize_corpus(corpus: list[str]):
    """``doc.split()`` tokenisation (bm25.py:55,62,71,143) -> (word->id dict, doc_ptr, token ids)."""
    vocab: dict[str, int] = {}
    ptr = np.zeros(len(corpus) + 1, dtype=np.int64)
    toks: list[int] = []
    for i, doc in enumerate(corpus):
        for w in doc.split():
            t = vocab.get(w)
            if t is None:
                t = len(vocab)
                vocab[w] = t
            toks.append(t)
        ptr[i + 1] = len(toks)
    return vocab, ptr, np.asarray(toks, dtype=np.int32)


def idf_table(df: np.ndarray, n_docs: int, variant: str) -> np.ndarray:
    """Per-term idf with ``math.log10`` (the reference's libm call), evaluated once per distinct df."""
    uniq, inv = np.unique(df, return_inverse=True)
    if variant == "bm25":      # bm25.py:147
        vals = [math.log10((n_docs - int(d) + 0.5) / (int(d) + 0.5)) for d in uniq]
    else:                      # bm25.py:87 (TF-IDF) and :173 (ATIRE)
        vals = [math.log10((n_docs + 1) / (int(d) + 1)) for d in uniq]
    return np.asarray(vals, dtype=np.float64)[inv]


class LexicalOracle:
    """Exhaustive fp64 scorer over a term->postings CSR built from token ids."""

    def __init__(self, doc_ptr: np.ndarray, doc_tok: np.ndarray, vocab_size: int,
                 variant: str = "bm25", k1: float = 0.9, b: float = 0.4):
        assert variant in VARIANTS
        self.variant, self.k1, self.b = variant, k1, b
        self.n_docs = len(doc_ptr) - 1
        self.vocab_size = vocab_size
        doc_of_tok = np.repeat(np.arange(self.n_docs, dtype=np.int64), np.diff(doc_ptr))
        key = doc_tok.astype(np.int64) * self.n_docs + doc_of_tok       # (term, doc) sorted
        ukey, tf = np.unique(key, return_counts=True)
        self.post_term = (ukey // self.n_docs).astype(np.int32)
        self.post_doc = (ukey % self.n_docs).astype(np.int32)
        self.post_tf = tf.astype(np.int32)
        self.df = np.bincount(self.post_term, minlength=vocab_size).astype(np.int64)
        self.term_ptr = np.zeros(vocab_size + 1, dtype=np.int64)
        np.cumsum(self.df, out=self.term_ptr[1:])
        self.idf = idf_table(self.df, self.n_docs, variant)
        self.doc_len = np.diff(doc_ptr).astype(np.int64)
        # statistics.mean of ints == correctly rounded exact rational (bm25.py:138)
        self.avgdl = float(Fraction(int(self.doc_len.sum()), self.n_docs)) if self.n_docs else 0.0

    @classmethod
    def from_strings(cls, corpus: list[str], variant="bm25", k1=0.9, b=0.4):
        vocab, ptr, toks = tokenize_corpus(corpus)
        o = cls(ptr, toks, len(vocab), variant, k1, b)
        o.vocab = vocab
        return o

    def query_ids(self, query: str) -> np.ndarray:
        """Token ids of ``query.split()``; out-of-vocabulary words map to -1 (idf 0, tf 0 => contribute 0)."""
        return np.asarray([self.vocab.get(w, -1) for w in query.split()], dtype=np.int64)

    def scores(self, q_tokens: np.ndarray) -> np.ndarray:
        """fp64 score of every document, contributions added in query-token order (bm25.py:152-155)."""
        s = np.zeros(self.n_docs, dtype=np.float64)
        k1, b = self.k1, self.b
        if self.variant != "tfidf" and self.n_docs:
            kd_min = k1 * (1 - b + b * self.doc_len.astype(np.float64) / self.avgdl)
            if np.any(kd_min == 0.0):
                # reference: tf == 0 and k1*(...) == 0 -> 0.0/0.0 -> ZeroDivisionError
                raise ZeroDivisionError("float division by zero")
        for t in q_tokens:
            if t < 0 or t >= self.vocab_size:
                continue
            lo, hi = self.term_ptr[t], self.term_ptr[t + 1]
            docs = self.post_doc[lo:hi]
            tf = self.post_tf[lo:hi].astype(np.float64)
            idf = self.idf[t]
            if self.variant == "tfidf":
                contrib = tf * idf                                       # bm25.py:114
            else:
                dl = self.doc_len[docs].astype(np.float64)
                contrib = idf * (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * dl / self.avgdl))   # bm25.py:155
            s[docs] += contrib
        return s

    def search_ids(self, q_tokens: np.ndarray, top_k: int):
        s = self.scores(q_tokens)
        order = np.argsort(-s, kind="stable")[:top_k]                    # bm25.py:105
        return order.astype(np.int64), s[order]

    def search(self, query: str, top_k: int):
        ids, sc = self.search_ids(self.query_ids(query), top_k)
        return [{"corpus_id": int(i), "score": float(v)} for i, v in zip(ids, sc)]

    def search_all(self, queries: list[str], top_k: int):
        return [self.search(q, top_k) for q in queries]
