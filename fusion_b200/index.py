"""Device-resident index containers for the four retrievers.

  * LexicalIndex   term-major CSR postings with fp64 impacts (BM25 / TF-IDF / ATIRE)   <- src/retrievers/bm25.py:37-87,133-147
  * SparseIndex    term-major CSR postings with L2-normalised fp32 weights (SPLADE)     <- src/retrievers/splade/splade.py:88-99 output
  * DenseIndex     [N, d] fp32 + bf16 copies, normalised for cos_sim                    <- src/retrievers/hybrid.py:101-103
  * TokenStore     ColBERT token embeddings [T, 128] bf16 + per-doc offsets             <- src/utils/colbert_ir.py:175-205 (index)

Construction uses torch device ops for the data movement (sort / unique / cumsum); the arithmetic that must
match the reference bit for bit (idf with math.log10, fp64 impacts) is done on the host / in the CUDA library.
Every index covers a contiguous shard [doc_base, doc_base + n_docs) of the corpus; corpus-global statistics
(N, df, sum of doc lengths) are passed in so that sharded BM25 scores equal the unsharded ones.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from fractions import Fraction

import numpy as np
import torch

from . import ops
from ._lib import LEX_BM25, LEX_TFIDF, FusionB200Error

VARIANTS = ("tfidf", "bm25", "atire")
DEFAULT_TILE_DOCS = int(os.environ.get("FZ_TILE_DOCS", 2048))      # docs per shared-memory accumulator tile
LONG_LIST_MIN = 512


def idf_table(df: np.ndarray, n_docs: int, variant: str) -> np.ndarray:
    """idf per term with the reference's libm call (math.log10), once per distinct df.
    BM25: log10((N - df + 0.5) / (df + 0.5)) (bm25.py:147); TF-IDF / ATIRE: log10((N + 1) / (df + 1)) (:87, :173)."""
    uniq, inv = np.unique(df, return_inverse=True)
    if variant == "bm25":
        vals = [math.log10((n_docs - int(d) + 0.5) / (int(d) + 0.5)) for d in uniq]
    else:
        vals = [math.log10((n_docs + 1) / (int(d) + 1)) for d in uniq]
    return np.asarray(vals, dtype=np.float64)[inv]


def _term_major_csr(row_of_entry: torch.Tensor, term_of_entry: torch.Tensor, n_rows: int, n_terms: int):
    """Sort (term, doc) pairs term-major / doc-ascending.  -> (order, term_ptr int64 [V+1])."""
    key = term_of_entry.to(torch.int64) * n_rows + row_of_entry.to(torch.int64)
    order = torch.argsort(key)
    counts = torch.bincount(term_of_entry.to(torch.int64), minlength=n_terms)
    term_ptr = torch.zeros(n_terms + 1, dtype=torch.int64, device=key.device)
    term_ptr[1:] = torch.cumsum(counts, 0)
    return order, term_ptr, counts


def _tile_table(term_ptr, post_doc, df, n_docs, tile_docs, long_min):
    n_terms = df.numel()
    n_tiles = (n_docs + tile_docs - 1) // tile_docs
    long_terms = torch.nonzero(df >= long_min).flatten().to(torch.int32)
    long_row = torch.full((n_terms,), -1, dtype=torch.int32, device=df.device)
    long_row[long_terms.long()] = torch.arange(long_terms.numel(), dtype=torch.int32, device=df.device)
    off = ops.long_tile_offsets(term_ptr, post_doc, long_terms, tile_docs, n_tiles)
    return long_row, off


class LexicalIndex:
    """Inverted index for TF-IDF / BM25 / ATIRE-BM25 over one corpus shard."""

    def __init__(self, doc_ptr, doc_tok, vocab_size: int, variant: str = "bm25", k1: float = 0.9, b: float = 0.4,
                 device="cuda", doc_base: int = 0, tile_docs: int = DEFAULT_TILE_DOCS, long_min: int = LONG_LIST_MIN,
                 global_n_docs: int | None = None, global_df: np.ndarray | None = None,
                 global_sum_dl: int | None = None, stats_reduce=None):
        if variant not in VARIANTS:
            raise FusionB200Error(f"unknown lexical variant {variant!r}")
        self.variant, self.k1, self.b = variant, k1, b
        self.device = torch.device(device)
        self.vocab_size, self.doc_base, self.tile_docs = int(vocab_size), int(doc_base), int(tile_docs)
        doc_ptr = torch.as_tensor(doc_ptr, dtype=torch.int64, device=self.device)
        doc_tok = torch.as_tensor(doc_tok, device=self.device).to(torch.int64)
        self.n_docs = doc_ptr.numel() - 1
        lens = doc_ptr[1:] - doc_ptr[:-1]
        self.doc_len = lens.to(torch.int32)
        doc_of_tok = torch.repeat_interleave(torch.arange(self.n_docs, device=self.device), lens)
        ukey, tf = torch.unique(doc_tok * self.n_docs + doc_of_tok, return_counts=True)   # sorted (term, doc)
        self.post_doc = (ukey % self.n_docs).to(torch.int32)
        post_term = ukey // self.n_docs
        self.post_tf = tf.to(torch.int32)
        df_local = torch.bincount(post_term, minlength=self.vocab_size)
        self.term_ptr = torch.zeros(self.vocab_size + 1, dtype=torch.int64, device=self.device)
        self.term_ptr[1:] = torch.cumsum(df_local, 0)
        del ukey, tf, post_term, doc_of_tok
        # corpus-global statistics (bm25.py:133-147): N, df, avgdl = statistics.mean(doc_len)
        if stats_reduce is not None:     # sharded build: (n_local, df_local, sum_dl_local) -> corpus-global values
            global_n_docs, global_df, global_sum_dl = stats_reduce(self.n_docs, df_local.cpu().numpy(), int(lens.sum()))
        self.global_n_docs = int(global_n_docs if global_n_docs is not None else self.n_docs)
        df = df_local.cpu().numpy() if global_df is None else np.asarray(global_df)
        sum_dl = int(lens.sum()) if global_sum_dl is None else int(global_sum_dl)
        self.df = df
        self.avgdl = float(Fraction(sum_dl, self.global_n_docs)) if self.global_n_docs else 0.0
        self.idf = torch.from_numpy(idf_table(df, self.global_n_docs, variant)).to(self.device)
        self.long_row, self.long_tile_off = _tile_table(self.term_ptr, self.post_doc, df_local, self.n_docs,
                                                        self.tile_docs, long_min)
        self._dl_values = torch.unique(self.doc_len).cpu().numpy().astype(np.float64)
        self.update_params(k1, b)

    def update_params(self, k1: float, b: float) -> None:
        """Recompute the per-posting fp64 impacts for new (k1, b) (bm25.py:158-161)."""
        self.k1, self.b = k1, b
        variant = LEX_TFIDF if self.variant == "tfidf" else LEX_BM25
        self.impact = ops.lexical_impacts(self.term_ptr, self.post_doc, self.post_tf, self.doc_len, self.idf,
                                          self.avgdl, k1, b, variant)

    def check_params(self) -> None:
        """The reference divides 0.0 by ``tf + k1*(1 - b + b*dl/avgdl)`` for every doc that lacks a query term
        (bm25.py:155): when that is 0.0 (e.g. k1 == 0) Python raises ZeroDivisionError.  Same here."""
        if self.variant == "tfidf" or self.n_docs == 0:
            return
        kd = self.k1 * (1 - self.b + self.b * self._dl_values / self.avgdl)
        if np.any(kd == 0.0):
            raise ZeroDivisionError("float division by zero")

    def view(self) -> ops.PostingsView:
        return ops.PostingsView(self.term_ptr, self.post_doc, self.impact, self.long_row, self.long_tile_off,
                                self.n_docs, self.tile_docs)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.term_ptr, self.post_doc, self.post_tf, self.impact, self.long_row, self.long_tile_off, self.doc_len))


class SparseIndex:
    """Inverted index over sparse term-weight vectors (SPLADE): doc weights are L2-normalised at build time for
    cos_sim, so a query is scored as sum_t (q_t/|q|) * (d_t/|d|) over the terms they share."""

    def __init__(self, doc_ptr, doc_term, doc_weight, vocab_size: int, similarity: str = "cos_sim", device="cuda",
                 doc_base: int = 0, tile_docs: int = DEFAULT_TILE_DOCS, long_min: int = LONG_LIST_MIN):
        if similarity not in ("cos_sim", "dot"):
            raise FusionB200Error(f"unknown similarity {similarity!r}")
        self.similarity, self.vocab_size, self.doc_base, self.tile_docs = similarity, int(vocab_size), int(doc_base), int(tile_docs)
        self.device = torch.device(device)
        doc_ptr = torch.as_tensor(doc_ptr, dtype=torch.int64, device=self.device)
        term = torch.as_tensor(doc_term, device=self.device).to(torch.int64)
        w = torch.as_tensor(doc_weight, dtype=torch.float32, device=self.device)
        self.n_docs = doc_ptr.numel() - 1
        lens = doc_ptr[1:] - doc_ptr[:-1]
        row = torch.repeat_interleave(torch.arange(self.n_docs, device=self.device), lens)
        if similarity == "cos_sim":
            sq = torch.zeros(self.n_docs, dtype=torch.float32, device=self.device).index_add_(0, row, w * w)
            w = w / torch.clamp(torch.sqrt(sq), min=1e-12)[row]
        order, self.term_ptr, df = _term_major_csr(row, term, self.n_docs, self.vocab_size)
        self.post_doc = row[order].to(torch.int32)
        self.post_w = w[order].contiguous()
        self.long_row, self.long_tile_off = _tile_table(self.term_ptr, self.post_doc, df, self.n_docs, self.tile_docs, long_min)

    def view(self) -> ops.PostingsView:
        return ops.PostingsView(self.term_ptr, self.post_doc, self.post_w, self.long_row, self.long_tile_off,
                                self.n_docs, self.tile_docs)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.term_ptr, self.post_doc, self.post_w, self.long_row, self.long_tile_off))


def sparse_queries(q_ptr, q_term, q_weight, similarity: str, device):
    """Query-side CSR for SparseIndex: int32 ptr/terms, fp32 weights (L2-normalised for cos_sim)."""
    q_ptr = torch.as_tensor(q_ptr, dtype=torch.int64, device=device)
    t = torch.as_tensor(q_term, device=device).to(torch.int32)
    w = torch.as_tensor(q_weight, dtype=torch.float32, device=device)
    if similarity == "cos_sim":
        lens = q_ptr[1:] - q_ptr[:-1]
        row = torch.repeat_interleave(torch.arange(q_ptr.numel() - 1, device=device), lens)
        sq = torch.zeros(q_ptr.numel() - 1, dtype=torch.float32, device=device).index_add_(0, row, w * w)
        w = w / torch.clamp(torch.sqrt(sq), min=1e-12)[row]
    return q_ptr.to(torch.int32), t.contiguous(), w.contiguous()


@dataclass
class DenseIndex:
    """Dense embeddings of one corpus shard: fp32 (exact rescoring) and bf16 (tensor-core operand) copies,
    L2-normalised when the similarity is cos_sim (splade/base.py:194-196)."""
    d_f32: torch.Tensor | None
    d_bf16: torch.Tensor
    similarity: str
    doc_base: int = 0

    @classmethod
    def build(cls, embeddings: torch.Tensor, similarity: str = "cos_sim", keep_f32: bool = True, doc_base: int = 0):
        if similarity not in ("cos_sim", "dot"):
            raise FusionB200Error(f"unknown similarity {similarity!r}")
        d32, d16 = ops.normalize_rows(embeddings, normalize=similarity == "cos_sim", want_f32=keep_f32)
        return cls(d32, d16, similarity, doc_base)

    @property
    def n_docs(self) -> int:
        return self.d_bf16.shape[0]

    def prepare_queries(self, q: torch.Tensor):
        return ops.normalize_rows(q, normalize=self.similarity == "cos_sim")


@dataclass
class TokenStore:
    """ColBERT token embeddings of one corpus shard (rows of doc d are tok_ptr[d]:tok_ptr[d+1])."""
    tok_ptr: torch.Tensor    # int64 [N+1]
    tok_emb: torch.Tensor    # bf16 [T, 128]
    doc_base: int = 0

    @property
    def n_docs(self) -> int:
        return self.tok_ptr.numel() - 1
