"""Drop-in mirror of ``src/retrievers/bm25.py`` (classes TFIDF :33-126, BM25 :129-161, AtireBM25 :164-173).

Same constructor and method signatures, same return shapes (``list[{'corpus_id', 'score'}]`` sorted by score
descending with ties broken by the lower document index, every document ranked, zero and negative scores
included) and the same ZeroDivisionError for degenerate (k1, b).  Underneath, the corpus is a device-resident
CSR inverted index with precomputed fp64 impacts and queries are scored by the CUDA library
(``fz_sparse_topk_f64`` / ``fz_sparse_scores_f64`` + ``fz_rank_rows_f64``).  There is no CPU scoring path.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .. import ops
from ..index import LexicalIndex

# above this many results per query the whole score row is materialised and sorted (the reference's
# `top_k = len(documents)` mode, hybrid.py:74); below it the threshold-filter top-k pipeline runs.
FULL_RANKING_MIN_K = 2049
FULL_RANKING_MAX_DOCS = 1 << 22


class TFIDF:
    """TF-IDF retrieval model (bm25.py:33-126)."""
    _variant = "tfidf"

    def __init__(self, corpus: list[str], device: str = "cuda", device_tokenizer: bool = False, **index_kwargs):
        self.corpus = corpus
        self.corpus_size = len(corpus)
        self._device_vocab = None
        if device_tokenizer:
            # tokenise on the device (fusion_b200.text): no per-word Python loop; the vocabulary is a sorted hash table,
            # so the string-keyed views (`vocab`, `df`, `idf`) are not available in this mode
            from .. import text
            self._device_vocab, doc_ptr, doc_tok = text.tokenize_corpus(corpus, device)
            self._vocab = None
            n_terms = len(self._device_vocab)
        else:
            self._vocab, doc_ptr, doc_tok = self._tokenize(corpus)
            n_terms = len(self._vocab)
        self.index = LexicalIndex(doc_ptr, doc_tok, n_terms, variant=self._variant,
                                  k1=getattr(self, "k1", 0.0), b=getattr(self, "b", 0.0), device=device, **index_kwargs)
        self.doc_len = self.index.doc_len.cpu().tolist()
        self.avgdl = self.index.avgdl

    def __repr__(self):
        return f"{self.__class__.__name__}".lower()

    # -- index build (bm25.py:53-83): whitespace tokens, vocabulary in first-appearance order
    @staticmethod
    def _tokenize(corpus):
        vocab: dict[str, int] = {}
        ptr = np.zeros(len(corpus) + 1, dtype=np.int64)
        toks: list[int] = []
        for i, doc in enumerate(corpus):
            for w in doc.split():
                t = vocab.get(w)
                if t is None:
                    t = vocab[w] = len(vocab)
                toks.append(t)
            ptr[i + 1] = len(toks)
        return vocab, ptr, np.asarray(toks, dtype=np.int32)

    @property
    def vocab(self):
        self._need_string_vocab("vocab")
        return set(self._vocab)

    def get_vocab(self):
        """Return the vocabulary sorted by alphabetical order (bm25.py:48-50)."""
        self._need_string_vocab("get_vocab")
        return sorted(self._vocab)

    @property
    def df(self):
        self._need_string_vocab("df")
        return {w: int(self.index.df[t]) for w, t in self._vocab.items()}

    @property
    def idf(self):
        self._need_string_vocab("idf")
        idf = self.index.idf.cpu().numpy()
        return {w: float(idf[t]) for w, t in self._vocab.items()}

    # -- queries
    def _encode_queries(self, queries: list[str]):
        if self._device_vocab is not None:
            from .. import text
            return text.tokenize_queries(queries, self._device_vocab, self.index.device)
        ptr = np.zeros(len(queries) + 1, dtype=np.int32)
        toks: list[int] = []
        for i, q in enumerate(queries):
            toks.extend(self._vocab.get(w, -1) for w in q.split())      # OOV: idf 0, tf 0 -> contributes 0 (bm25.py:111-112)
            ptr[i + 1] = len(toks)
        dev = self.index.device
        return torch.from_numpy(ptr).to(dev), torch.tensor(toks, dtype=torch.int32, device=dev)

    def search_all_tensors(self, queries: list[str], top_k: int):
        """-> (scores float64 [Q, k], ids int32 [Q, k]) on the device, k = min(top_k, N)."""
        self.index.check_params()
        q_ptr, q_term = self._encode_queries(queries)
        return self._search_ids(q_ptr, q_term, top_k)

    def _search_ids(self, q_ptr, q_term, top_k: int):
        k = min(top_k, self.index.n_docs)
        view = self.index.view()
        if k >= FULL_RANKING_MIN_K or 2 * k > ops.DEFAULT_CAP:
            if self.index.n_docs > FULL_RANKING_MAX_DOCS:
                raise ops.FusionB200Error(f"top_k={top_k} on {self.index.n_docs} documents: full ranking is limited to "
                                          f"{FULL_RANKING_MAX_DOCS} documents")
            nq = q_ptr.numel() - 1
            out_s = torch.empty((nq, k), dtype=torch.float64, device=self.index.device)
            out_i = torch.empty((nq, k), dtype=torch.int32, device=self.index.device)
            step = max(1, (1 << 30) // (8 * self.index.n_docs))
            ptr_h = q_ptr.cpu()
            for lo in range(0, nq, step):
                hi = min(nq, lo + step)
                sub_ptr = (q_ptr[lo:hi + 1] - q_ptr[lo]).contiguous()
                sub_term = q_term[int(ptr_h[lo]):int(ptr_h[hi])].contiguous()
                if sub_term.numel() == 0:
                    sub_term = torch.full((1,), -1, dtype=torch.int32, device=self.index.device)
                scores = ops.sparse_scores(view, sub_ptr, sub_term)
                out_s[lo:hi], out_i[lo:hi] = ops.rank_rows(scores, k, self.index.doc_base)
            return out_s, out_i
        if q_term.numel() == 0:
            q_term = torch.full((1,), -1, dtype=torch.int32, device=self.index.device)
        return ops.sparse_topk(view, q_ptr, q_term, None, k, self.index.doc_base)

    def search_all(self, queries: list[str], top_k: int) -> list:
        """Perform retrieval on all provided queries (bm25.py:89-98)."""
        t0 = time.perf_counter()
        scores, ids = self.search_all_tensors(queries, top_k)
        scores, ids = scores.cpu().tolist(), ids.cpu().tolist()
        t1 = time.perf_counter()
        if queries:
            print(f"Avg. latency (ms/quey): {((t1 - t0) / len(queries)) * 1000}")
        return [[{'corpus_id': i, 'score': s} for i, s in zip(qi, qs)] for qi, qs in zip(ids, scores)]

    def search(self, query: str, top_k: int) -> list:
        """Perform retrieval on a single query (bm25.py:100-106)."""
        scores, ids = self.search_all_tensors([query], top_k)
        return [{'corpus_id': i, 'score': s} for i, s in zip(ids[0].cpu().tolist(), scores[0].cpu().tolist())]

    def score(self, query: str, doc_idx: int) -> float:
        """Score of one (query, document) pair (bm25.py:108-115, :149-156)."""
        self.index.check_params()
        q_ptr, q_term = self._encode_queries([query])
        if q_term.numel() == 0:
            return 0.0
        return float(ops.sparse_scores(self.index.view(), q_ptr, q_term)[0, doc_idx])

    @property
    def tf(self):
        """``{word: {doc_idx: count}}`` like the reference's tf index (bm25.py:62-70), materialised from the device CSR."""
        self._need_string_vocab("tf")
        ix = self.index
        ptr, doc, cnt = ix.term_ptr.cpu().numpy(), ix.post_doc.cpu().numpy(), ix.post_tf.cpu().numpy()
        return {w: dict(zip(doc[ptr[t]:ptr[t + 1]].tolist(), cnt[ptr[t]:ptr[t + 1]].tolist())) for w, t in self._vocab.items()}

    def _need_string_vocab(self, what: str) -> None:
        if self._vocab is None:
            raise ops.FusionB200Error(f"`{what}` needs the string vocabulary, which device_tokenizer=True does not keep "
                                      "(the device vocabulary is a table of 128-bit token hashes)")

    def _index_path(self, output_dir: str, dataset: str) -> str:
        return os.path.join(output_dir, f'{self.__repr__()}_index_{dataset}.npz')

    def save_indexes(self, output_dir: str, dataset: str, reference_pickles: bool = True) -> None:
        """Save the indexes to disk (bm25.py:117-126).  Writes the reference's four pickles
        ``<name>_{vocab,tf,df,idf}_<dataset>.pkl`` (a set, a dict of dicts, a Counter and a dict, like the reference's
        attributes) and, next to them, ``<name>_index_<dataset>.npz``: the vocabulary and the CSR arrays
        :meth:`load_indexes` restores the device index from without re-tokenising the corpus."""
        import pickle
        from collections import Counter
        self._need_string_vocab("save_indexes")
        ix = self.index
        words = np.array(sorted(self._vocab, key=self._vocab.get))
        np.savez_compressed(self._index_path(output_dir, dataset), vocab=words, term_ptr=ix.term_ptr.cpu().numpy(),
                            post_doc=ix.post_doc.cpu().numpy(), post_tf=ix.post_tf.cpu().numpy(),
                            doc_len=ix.doc_len.cpu().numpy(), df=ix.df, idf=ix.idf.cpu().numpy(),
                            params=np.array([getattr(self, "k1", 0.0), getattr(self, "b", 0.0)]))
        if reference_pickles:
            for name, obj in (("vocab", self.vocab), ("tf", self.tf), ("df", Counter(self.df)), ("idf", self.idf)):
                with open(os.path.join(output_dir, f'{self.__repr__()}_{name}_{dataset}.pkl'), 'wb') as f:
                    pickle.dump(obj, f)

    @classmethod
    def load_indexes(cls, output_dir: str, dataset: str, corpus: list[str] | None = None, device: str = "cuda", **kw):
        """Restore a retriever saved by :meth:`save_indexes` (the reference has no loader: it rebuilds its dicts from the
        corpus every run).  ``corpus`` is only kept for ``self.corpus``; the index comes from the file."""
        g = np.load(os.path.join(output_dir, f'{cls.__name__.lower()}_index_{dataset}.npz'), allow_pickle=False)
        self = cls.__new__(cls)
        self.corpus = corpus
        self._device_vocab = None
        self._vocab = {str(w): i for i, w in enumerate(g["vocab"])}
        k1, b = (float(x) for x in g["params"])
        if cls._variant != "tfidf":
            self.k1, self.b = kw.pop("k1", k1), kw.pop("b", b)
        self.index = LexicalIndex.from_postings(g["term_ptr"], g["post_doc"], g["post_tf"], g["doc_len"], len(self._vocab),
                                                variant=cls._variant, k1=getattr(self, "k1", 0.0), b=getattr(self, "b", 0.0),
                                                device=device, **kw)
        self.corpus_size = self.index.n_docs
        self.doc_len = self.index.doc_len.cpu().tolist()
        self.avgdl = self.index.avgdl
        return self


class BM25(TFIDF):
    """BM25 retrieval model (bm25.py:129-161)."""
    _variant = "bm25"

    def __init__(self, corpus: list[str], k1: float, b: float, device: str = "cuda", **index_kwargs):
        self.b = b
        self.k1 = k1
        super().__init__(corpus, device=device, **index_kwargs)

    def update_params(self, k1: float, b: float) -> None:
        """Update the BM25 parameters (bm25.py:158-161)."""
        self.k1 = k1
        self.b = b
        self.index.update_params(k1, b)


class AtireBM25(BM25):
    """ATIRE BM25: idf = log10((N + 1) / (df + 1)) (bm25.py:164-173)."""
    _variant = "atire"
