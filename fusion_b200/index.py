"""Device-resident index containers for the four retrievers.

  * LexicalIndex   term-major CSR postings with fp64 impacts (BM25 / TF-IDF / ATIRE)   <- src/retrievers/bm25.py:37-87,133-147
  * SparseIndex    term-major CSR postings with L2-normalised fp32 weights (SPLADE)     <- src/retrievers/splade/splade.py:88-99 output
  * DenseIndex     [N, d] fp32 + bf16 copies, normalised for cos_sim                    <- src/retrievers/hybrid.py:101-103
  * TokenStore     ColBERT token embeddings [T, 128] bf16 + per-doc offsets             <- src/utils/colbert_ir.py:175-205 (index)

Construction goes through the library's ``fz_build_*`` entry points (csrc/build.cu: sort, CSR transpose, the three-form
postings layout, SPLADE head matrix), so a non-Python host can build the same index; the arithmetic that must match the
reference bit for bit (idf with math.log10, fp64 impacts) is done on the host / in the CUDA library.
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
LEX_TILE_DOCS = int(os.environ.get("FZ_TILE_DOCS_LEX", 1024))   # fp64 accumulators (BM25 / TF-IDF)
SP_TILE_DOCS = int(os.environ.get("FZ_TILE_DOCS_SP", DEFAULT_TILE_DOCS))     # fp32 accumulators (SPLADE)
DENSE_FRAC = float(os.environ.get("FZ_DENSE_FRAC", 0.25))   # df / N from which a term is stored as a dense row


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
    """Sort (term, doc) pairs term-major / doc-ascending (``fz_build_term_major``).  -> (order, term_ptr int64 [V+1])."""
    order, term_ptr = ops.build_term_major(row_of_entry.to(torch.int32), term_of_entry.to(torch.int32), n_rows, n_terms)
    return order, term_ptr, term_ptr[1:] - term_ptr[:-1]


def build_postings(term_ptr: torch.Tensor, post_doc: torch.Tensor, post_val: torch.Tensor, n_docs: int, tile_docs: int,
                   tiled_min: int | None = None, dense_frac: float = DENSE_FRAC) -> ops.PostingsView:
    """Term-major CSR (doc-ascending inside a term) -> the three storage forms of ``fz_postings_t``, built on the device by
    ``fz_build_postings_plan`` / ``fz_build_postings_fill`` (csrc/build.cu).

    short  df < tiled_min: kept as (doc, value) pairs, plus coarse marks (postings of the term below tile 16*c).
    tiled  per (term, tile) segments of (uint16 tile-relative offset, value), padded to 4 with (tile_docs, 0) and ordered
           for conflict-free shared-memory scatter: the postings of a segment are dealt round-robin over the 32 banks
           (doc % 32), and that sequence is laid out so that lane l of a warp reads element l of a run of 32 with its
           j-th accumulate (a thread owns the postings of one 16-byte value vector: 4 fp32 or 2 fp64).
    dense  df >= dense_frac * n_docs: one value per document, zero where the term is absent.
    """
    if tile_docs % 4 or not (4 <= tile_docs <= 32768):      # (the K2 kernels take <= 8192; the SPLADE tail kernel 32768)
        raise FusionB200Error(f"tile_docs={tile_docs} must be a multiple of 4 in [4, 32768]")
    if tiled_min is None:
        tiled_min = int(os.environ.get("FZ_TILED_MIN", 512))
    if tiled_min > 65535:
        raise FusionB200Error("tiled_min must be <= 65535 (short-list offsets are 16 bits)")
    dense_min = max(tiled_min, int(math.ceil(dense_frac * n_docs))) if dense_frac > 0 else (1 << 62)
    return ops.build_postings(term_ptr, post_doc, post_val, n_docs, tile_docs, tiled_min, dense_min)


class LexicalIndex:
    """Inverted index for TF-IDF / BM25 / ATIRE-BM25 over one corpus shard."""

    def __init__(self, doc_ptr, doc_tok, vocab_size: int, variant: str = "bm25", k1: float = 0.9, b: float = 0.4,
                 device="cuda", doc_base: int = 0, tile_docs: int = LEX_TILE_DOCS, tiled_min: int | None = None,
                 dense_frac: float = DENSE_FRAC, global_n_docs: int | None = None, global_df: np.ndarray | None = None,
                 global_sum_dl: int | None = None, stats_reduce=None):
        if variant not in VARIANTS:
            raise FusionB200Error(f"unknown lexical variant {variant!r}")
        self.variant, self.k1, self.b = variant, k1, b
        self.device = torch.device(device)
        self.vocab_size, self.doc_base, self.tile_docs = int(vocab_size), int(doc_base), int(tile_docs)
        self.tiled_min, self.dense_frac = tiled_min, dense_frac
        doc_ptr = torch.as_tensor(doc_ptr, dtype=torch.int64, device=self.device).contiguous()
        doc_tok = torch.as_tensor(doc_tok, device=self.device).to(torch.int32).contiguous()
        self.n_docs = doc_ptr.numel() - 1
        lens = doc_ptr[1:] - doc_ptr[:-1]
        # (term, doc, tf) postings, term-major / doc-ascending: fz_build_lexical_* (device sort + run lengths)
        self.term_ptr, self.post_doc, self.post_tf, self.doc_len = ops.build_lexical_postings(doc_ptr, doc_tok, self.vocab_size)
        df_local = self.term_ptr[1:] - self.term_ptr[:-1]
        # corpus-global statistics (bm25.py:133-147): N, df, avgdl = statistics.mean(doc_len)
        if stats_reduce is not None:     # sharded build: (n_local, df_local, sum_dl_local) -> corpus-global values
            global_n_docs, global_df, global_sum_dl = stats_reduce(self.n_docs, df_local.cpu().numpy(), int(lens.sum()))
        self.global_n_docs = int(global_n_docs if global_n_docs is not None else self.n_docs)
        df = df_local.cpu().numpy() if global_df is None else np.asarray(global_df)
        sum_dl = int(lens.sum()) if global_sum_dl is None else int(global_sum_dl)
        self.df = df
        self.avgdl = float(Fraction(sum_dl, self.global_n_docs)) if self.global_n_docs else 0.0
        self.idf = torch.from_numpy(idf_table(df, self.global_n_docs, variant)).to(self.device)
        self._dl_values = torch.unique(self.doc_len).cpu().numpy().astype(np.float64)
        self.update_params(k1, b)

    @classmethod
    def from_postings(cls, term_ptr, post_doc, post_tf, doc_len, vocab_size: int, variant: str = "bm25", k1: float = 0.9,
                      b: float = 0.4, device="cuda", doc_base: int = 0, tile_docs: int = LEX_TILE_DOCS,
                      tiled_min: int | None = None, dense_frac: float = DENSE_FRAC, global_n_docs: int | None = None,
                      global_df=None, avgdl: float | None = None):
        """Rebuild the index from saved term-major CSR arrays (``BM25.load_indexes``): no tokenisation, no sort."""
        if variant not in VARIANTS:
            raise FusionB200Error(f"unknown lexical variant {variant!r}")
        self = cls.__new__(cls)
        self.variant, self.k1, self.b = variant, k1, b
        self.device = torch.device(device)
        self.vocab_size, self.doc_base, self.tile_docs = int(vocab_size), int(doc_base), int(tile_docs)
        self.tiled_min, self.dense_frac = tiled_min, dense_frac
        self.term_ptr = torch.as_tensor(np.asarray(term_ptr), dtype=torch.int64, device=self.device)
        self.post_doc = torch.as_tensor(np.asarray(post_doc), dtype=torch.int32, device=self.device)
        self.post_tf = torch.as_tensor(np.asarray(post_tf), dtype=torch.int32, device=self.device)
        self.doc_len = torch.as_tensor(np.asarray(doc_len), dtype=torch.int32, device=self.device)
        self.n_docs = self.doc_len.numel()
        # corpus-global statistics of a sharded index are stored with the shard (N, df, avgdl)
        self.global_n_docs = int(global_n_docs) if global_n_docs is not None else self.n_docs
        self.df = np.asarray(global_df) if global_df is not None else (self.term_ptr[1:] - self.term_ptr[:-1]).cpu().numpy()
        sum_dl = int(self.doc_len.long().sum())
        self.avgdl = float(avgdl) if avgdl is not None else (float(Fraction(sum_dl, self.global_n_docs)) if self.global_n_docs else 0.0)
        self.idf = torch.from_numpy(idf_table(self.df, self.global_n_docs, variant)).to(self.device)
        self._dl_values = torch.unique(self.doc_len).cpu().numpy().astype(np.float64)
        self.update_params(k1, b)
        return self

    def save(self, path: str) -> None:
        """Persist the shard's term-major CSR and statistics (``.npz``); ``load`` rebuilds the device index from it without
        tokenising or sorting.  (The reference pickles its dicts, bm25.py:117-126, and has no loader.)"""
        np.savez(path, term_ptr=self.term_ptr.cpu().numpy(), post_doc=self.post_doc.cpu().numpy(),
                 post_tf=self.post_tf.cpu().numpy(), doc_len=self.doc_len.cpu().numpy(), df=np.asarray(self.df),
                 meta=np.array([self.vocab_size, self.doc_base, self.global_n_docs, self.tile_docs], dtype=np.int64),
                 params=np.array([self.k1, self.b, self.avgdl, self.dense_frac], dtype=np.float64), variant=np.array(self.variant))

    @classmethod
    def load(cls, path: str, device="cuda", tiled_min: int | None = None):
        g = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
        vocab, doc_base, n_global, tile_docs = (int(x) for x in g["meta"])
        k1, b, avgdl, dense_frac = (float(x) for x in g["params"])
        self = cls.from_postings(g["term_ptr"], g["post_doc"], g["post_tf"], g["doc_len"], vocab, str(g["variant"]), k1, b,
                                 device=device, doc_base=doc_base, tile_docs=tile_docs, tiled_min=tiled_min, dense_frac=dense_frac,
                                 global_n_docs=n_global, global_df=g["df"], avgdl=avgdl)
        return self

    def update_params(self, k1: float, b: float) -> None:
        """Recompute the per-posting fp64 impacts for new (k1, b) (bm25.py:158-161)."""
        self.k1, self.b = k1, b
        variant = LEX_TFIDF if self.variant == "tfidf" else LEX_BM25
        impact = ops.lexical_impacts(self.term_ptr, self.post_doc, self.post_tf, self.doc_len, self.idf,
                                     self.avgdl, k1, b, variant)
        self._view = build_postings(self.term_ptr, self.post_doc, impact, self.n_docs, self.tile_docs, self.tiled_min,
                                    self.dense_frac)

    def check_params(self) -> None:
        """The reference divides 0.0 by ``tf + k1*(1 - b + b*dl/avgdl)`` for every doc that lacks a query term
        (bm25.py:155): when that is 0.0 (e.g. k1 == 0) Python raises ZeroDivisionError.  Same here."""
        if self.variant == "tfidf" or self.n_docs == 0:
            return
        kd = self.k1 * (1 - self.b + self.b * self._dl_values / self.avgdl)
        if np.any(kd == 0.0):
            raise ZeroDivisionError("float division by zero")

    def view(self) -> ops.PostingsView:
        return self._view

    def nbytes(self) -> int:
        return self._view.nbytes() + sum(t.numel() * t.element_size() for t in
                                         (self.term_ptr, self.post_doc, self.post_tf, self.doc_len))


SP_HEAD_DIM = int(os.environ.get("FZ_SPLADE_HEAD", 256))           # head terms scored on the tensor cores (multiple of 64, <= 256)
SP_TAIL_TILE_DOCS = int(os.environ.get("FZ_TILE_DOCS_TAIL", 8192))   # docs per fixed-point accumulator tile of the tail kernel
SP_BOOT_DOCS = int(os.environ.get("FZ_SPLADE_BOOT", 262144))         # docs of the threshold bootstrap (0 = none)


class SparseIndex:
    """Index over sparse term-weight vectors (SPLADE): doc weights are L2-normalised at build time for cos_sim, so a
    query is scored as sum_t (q_t/|q|) * (d_t/|d|) over the terms they share (hybrid.py:101-103 on dense [., V] vectors).

    Two device forms are kept:
      * the HEAD / TAIL split of ``fz_splade_topk`` (non-negative weights only): the ``head_dim`` most frequent terms as a
        dense doc-major bf16 matrix (tensor-core operand), the other terms as an inverted index for the tail bound, and
        the doc-major (term, weight) copy the survivors are rescored from exactly;
      * the general three-form inverted index of ``fz_sparse_topk_f32`` / ``fz_sparse_scores_f32`` over ALL terms, built
        lazily on first use (full-ranking mode, queries the fast path hands back, negative weights).
    """

    def __init__(self, doc_ptr, doc_term, doc_weight, vocab_size: int, similarity: str = "cos_sim", device="cuda",
                 doc_base: int = 0, tile_docs: int = SP_TILE_DOCS, tiled_min: int | None = None,
                 dense_frac: float = DENSE_FRAC, head_dim: int | None = None, tail_tile_docs: int | None = None,
                 boot_docs: int | None = None):
        if similarity not in ("cos_sim", "dot"):
            raise FusionB200Error(f"unknown similarity {similarity!r}")
        self.similarity, self.vocab_size, self.doc_base, self.tile_docs = similarity, int(vocab_size), int(doc_base), int(tile_docs)
        self.tiled_min, self.dense_frac = tiled_min, dense_frac
        self.device = torch.device(device)
        self.doc_ptr = torch.as_tensor(doc_ptr, dtype=torch.int64, device=self.device).contiguous()
        term = torch.as_tensor(doc_term, device=self.device).to(torch.int32)
        w = torch.as_tensor(doc_weight, dtype=torch.float32, device=self.device)
        self.n_docs = self.doc_ptr.numel() - 1
        lens = self.doc_ptr[1:] - self.doc_ptr[:-1]
        row = torch.repeat_interleave(torch.arange(self.n_docs, device=self.device, dtype=torch.int32), lens)
        term = term.contiguous()
        if similarity == "cos_sim":
            w = ops.build_csr_normalize(self.doc_ptr, w.contiguous())
        # doc-major copy: (int32 term, float weight) pairs, 8 bytes per posting
        self.doc_post = torch.stack([term, w.view(torch.int32)], dim=1).contiguous()
        self._df, self._term_max, flags = ops.build_term_stats(term, w.contiguous(), self.vocab_size)
        if flags & 2:
            raise FusionB200Error("a term id lies outside [0, vocab_size)")
        self.nonneg = not (flags & 1)
        self._view = None
        self.head = None
        head_dim = SP_HEAD_DIM if head_dim is None else int(head_dim)
        if self.nonneg and head_dim > 0 and self.n_docs > 0:
            self._build_head_tail(row, term, w, head_dim, int(tail_tile_docs or SP_TAIL_TILE_DOCS))
            # threshold bootstrap (ops.splade_topk): a general index over the first boot_docs docs, when they are a small part
            # of the shard.  Sharded runs must pass the same boot_docs on every rank (the round schedule starts there).
            explicit = boot_docs is not None
            boot_docs = SP_BOOT_DOCS if boot_docs is None else int(boot_docs)
            boot_docs = boot_docs // 256 * 256
            if boot_docs >= 256 and (self.n_docs >= 8 * boot_docs if not explicit else self.n_docs >= 2 * boot_docs):
                self.head.boot = self._general_view(boot_docs)

    def _build_head_tail(self, row, term, w, head_dim, tail_tile_docs):
        dev, n, v = self.device, self.n_docs, self.vocab_size
        if head_dim % 64 or not (64 <= head_dim <= 256):
            raise FusionB200Error(f"head_dim={head_dim} must be 64, 128, 192 or 256")
        df = self._df
        n_head = min(head_dim, int((df > 0).sum()))
        head_terms = torch.topk(df, n_head).indices if n_head else torch.zeros(0, dtype=torch.int64, device=dev)
        term_head = torch.full((v,), -1, dtype=torch.int32, device=dev)
        term_head[head_terms] = torch.arange(n_head, dtype=torch.int32, device=dev)
        term_max = self._term_max
        head = ops.build_splade_head(self.doc_ptr, term, w.contiguous(), term_head, head_dim)
        m = term_head[term.long()] < 0
        row_t, term_t, w_t = row[m], term[m], w[m]
        del m
        order, term_ptr_t, _ = _term_major_csr(row_t, term_t, n, v)
        tail_tile_docs = max(256, min(tail_tile_docs, (n + 255) // 256 * 256))
        # (tail tiles are small: terms rarer than one posting per tile stay plain doc-ascending lists with coarse marks)
        tail_tiled_min = self.tiled_min if self.tiled_min is not None else int(os.environ.get("FZ_TAIL_TILED_MIN", 512))
        tail = build_postings(term_ptr_t, row_t[order].to(torch.int32), w_t[order].contiguous(), n, tail_tile_docs,
                              tail_tiled_min, dense_frac=0.0)
        self.head = ops.SpladeHeadView(head, term_head, term_max, self.doc_ptr, self.doc_post, tail, head_dim, v, n,
                                       unit_rows=self.similarity == "cos_sim")

    def save(self, path: str) -> None:
        """Persist the shard as its doc-major CSR (normalised weights) + parameters; ``load`` rebuilds the device forms."""
        hd = self.head.head_dim if self.head is not None else 0
        boot = self.head.boot.n_docs if (self.head is not None and self.head.boot is not None) else 0
        np.savez(path, doc_ptr=self.doc_ptr.cpu().numpy(), doc_term=self.doc_post[:, 0].cpu().numpy(),
                 doc_weight=self.doc_post[:, 1].contiguous().view(torch.float32).cpu().numpy(),
                 meta=np.array([self.vocab_size, self.doc_base, self.tile_docs, hd, boot], dtype=np.int64),
                 similarity=np.array(self.similarity), dense_frac=np.array(self.dense_frac))

    @classmethod
    def load(cls, path: str, device="cuda"):
        g = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
        vocab, doc_base, tile_docs, hd, boot = (int(x) for x in g["meta"])
        # the stored weights are already normalised: build with "dot" (no renormalisation), then restore the similarity
        self = cls(g["doc_ptr"], g["doc_term"], g["doc_weight"], vocab, "dot", device=device, doc_base=doc_base, tile_docs=tile_docs,
                   dense_frac=float(g["dense_frac"]), head_dim=hd, boot_docs=boot)
        self.similarity = str(g["similarity"])
        if self.head is not None:
            self.head.unit_rows = self.similarity == "cos_sim"
        return self

    def _general_view(self, n_docs: int) -> ops.PostingsView:
        """The three-form inverted index over all terms of the first ``n_docs`` docs."""
        nnz = int(self.doc_ptr[n_docs])
        lens = self.doc_ptr[1:n_docs + 1] - self.doc_ptr[:n_docs]
        row = torch.repeat_interleave(torch.arange(n_docs, device=self.device), lens)
        term = self.doc_post[:nnz, 0].long()
        w = self.doc_post[:nnz, 1].contiguous().view(torch.float32)
        order, term_ptr, _ = _term_major_csr(row, term, n_docs, self.vocab_size)
        return build_postings(term_ptr, row[order].to(torch.int32), w[order].contiguous(), n_docs, self.tile_docs,
                              self.tiled_min, self.dense_frac)

    def view(self) -> ops.PostingsView:
        """The general inverted index over all terms (built on first use)."""
        if self._view is None:
            self._view = self._general_view(self.n_docs)
        return self._view

    def topk(self, q_ptr, q_term, q_weight, k: int, cap: int = ops.DEFAULT_CAP, sync: "ops.ShardSync | None" = None,
             defer: bool = False):
        """Top-k of one query batch: the head/tail pipeline when the index has one, else the general inverted index."""
        if self.head is not None:
            return ops.splade_topk(self, q_ptr, q_term, q_weight, k, self.doc_base, cap=cap, sync=sync, defer=defer)
        return ops.sparse_topk(self.view(), q_ptr, q_term, q_weight, k, self.doc_base, cap=cap, sync=sync, defer=defer)

    def nbytes(self) -> int:
        b = self.doc_ptr.numel() * 8 + self.doc_post.numel() * 4
        if self.head is not None:
            b += self.head.nbytes()
        if self._view is not None:
            b += self._view.nbytes()
        return b


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
    _err: tuple | None = None        # (max |d - bf16(d)|_2, max |d|_2) over the shard's rows

    @classmethod
    def build(cls, embeddings: torch.Tensor, similarity: str = "cos_sim", keep_f32: bool = True, doc_base: int = 0):
        if similarity not in ("cos_sim", "dot"):
            raise FusionB200Error(f"unknown similarity {similarity!r}")
        d32, d16 = ops.normalize_rows(embeddings, normalize=similarity == "cos_sim", want_f32=keep_f32)
        return cls(d32, d16, similarity, doc_base)

    @property
    def n_docs(self) -> int:
        return self.d_bf16.shape[0]

    def save(self, path: str) -> None:
        """Persist the (normalised) fp32 rows, or the bf16 rows of a throughput-mode index, + parameters."""
        rows = self.d_f32 if self.d_f32 is not None else self.d_bf16.view(torch.int16)
        np.savez(path, rows=rows.cpu().numpy(), is_f32=np.array(self.d_f32 is not None), similarity=np.array(self.similarity),
                 doc_base=np.array(self.doc_base))

    @classmethod
    def load(cls, path: str, device="cuda"):
        g = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
        rows = torch.from_numpy(g["rows"]).to(device)
        if bool(g["is_f32"]):
            d32, d16 = ops.normalize_rows(rows, normalize=False)       # stored rows are already normalised: convert only
            return cls(d32, d16, str(g["similarity"]), int(g["doc_base"]))
        return cls(None, rows.view(torch.bfloat16), str(g["similarity"]), int(g["doc_base"]))

    def prepare_queries(self, q: torch.Tensor):
        return ops.normalize_rows(q, normalize=self.similarity == "cos_sim")

    def exact_margin(self, q32: torch.Tensor, q16: torch.Tensor) -> float:
        """How far below the k-th best bf16 score the filter must reach so that exact (fp32) rescoring of the survivors
        returns the true top-k.  With r = x - bf16(x):  q.d - bf16(q).bf16(d) = r_q.d + bf16(q).r_d, so
        |error| <= |r_q||d| + |bf16(q)||r_d| (+ the tensor core's fp32 accumulation of exact bf16 products, bounded by
        dim * 2^-23 * |q||d|); the measured residual norms are ~0.0011 for unit vectors, the worst case 2^-9 per element
        would be 0.0039.  The cut moves by at most the error on either side, hence twice the bound."""
        if self.d_f32 is None:
            return 0.0
        if self._err is None:
            e_d, n_d = 0.0, 0.0
            for lo in range(0, self.d_f32.shape[0], 1 << 18):
                blk = self.d_f32[lo:lo + (1 << 18)]
                e_d = max(e_d, float((blk - self.d_bf16[lo:lo + (1 << 18)].float()).norm(dim=1).max()))
                n_d = max(n_d, float(blk.norm(dim=1).max()))
            self._err = (e_d, n_d)
        e_d, n_d = self._err
        q16f = q16.float()
        e_q = float((q32 - q16f).norm(dim=1).max())
        n_q = max(float(q32.norm(dim=1).max()), float(q16f.norm(dim=1).max()))
        err = e_q * n_d + n_q * e_d + q32.shape[1] * 2.0 ** -23 * n_q * n_d
        return 2.0 * err * (1.0 + 1e-3)


@dataclass
class TokenStore:
    """ColBERT token embeddings of one corpus shard (rows of doc d are tok_ptr[d]:tok_ptr[d+1]).  The MaxSim kernel
    streams the PACKED image (ops.pack_tokens), built on first use; ``drop_plain`` then frees the [T, 128] matrix."""
    tok_ptr: torch.Tensor            # int64 [N+1]
    tok_emb: torch.Tensor | None     # bf16 [T, 128]
    doc_base: int = 0
    _packed: tuple | None = None

    @property
    def n_docs(self) -> int:
        return self.tok_ptr.numel() - 1

    @property
    def n_tokens(self) -> int:
        return int(self.tok_ptr[-1])

    def save(self, path: str) -> None:
        """Persist the per-passage offsets and the bf16 token rows (the packed image is rebuilt on load)."""
        if self.tok_emb is None:
            raise FusionB200Error("the plain token matrix was dropped (packed(drop_plain=True)): nothing to save")
        np.savez(path, tok_ptr=self.tok_ptr.cpu().numpy(), tok_emb=self.tok_emb.view(torch.int16).cpu().numpy(),
                 doc_base=np.array(self.doc_base))

    @classmethod
    def load(cls, path: str, device="cuda"):
        g = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
        return cls(torch.from_numpy(g["tok_ptr"]).to(device), torch.from_numpy(g["tok_emb"]).to(device).view(torch.bfloat16),
                   int(g["doc_base"]))

    def packed(self, drop_plain: bool = False):
        if self._packed is None:
            self._packed = ops.pack_tokens(self.tok_ptr, self.tok_emb)
        if drop_plain:
            self.tok_emb = None
        return self._packed
