"""Mirror of the scoring section of ``src/utils/sentence_transformers.py``:
``InformationRetrievalEvaluatorCustom.compute_metrices`` (:314-393) and ``compute_metrics`` (:395-485).

The reference walks the corpus in chunks of 50,000, scores every query against every chunk with one GEMV,
takes ``torch.topk`` and merges through Python heaps (:334-364).  Here the whole corpus is scored by the
tcgen05 filter GEMM with the top-``max_k`` fused into its epilogue; the result lists have the reference's shape
(``{'corpus_id', 'score'}``), and the metric bookkeeping below reproduces its definitions.
"""
from __future__ import annotations

import inspect
import logging
import time

import numpy as np
import torch

from ..retrievers.hybrid import Ranker

logger = logging.getLogger(__name__)


def _score_matrix(a, b, normalize: bool):
    from .. import ops
    a = torch.as_tensor(a).float().cuda()
    b = torch.as_tensor(b).float().cuda()
    a, b = (a[None] if a.dim() == 1 else a), (b[None] if b.dim() == 1 else b)
    a32, _ = ops.normalize_rows(a, normalize=normalize, want_bf16=False)
    b32, _ = ops.normalize_rows(b, normalize=normalize, want_bf16=False)
    return ops.dense_scores(a32, b32)


def cos_sim(a, b):
    """``sentence_transformers.util.cos_sim``: [A, d] x [B, d] -> cosine matrix [A, B] (exact fp32, CUDA library).  Also
    the key under which ``InformationRetrievalEvaluatorCustom`` selects cosine scoring."""
    return _score_matrix(a, b, True)


def dot_score(a, b):
    """``sentence_transformers.util.dot_score``: [A, d] x [B, d] -> dot-product matrix [A, B]."""
    return _score_matrix(a, b, False)


class InformationRetrievalEvaluatorCustom:
    def __init__(self, queries, corpus, relevant_docs, corpus_chunk_size: int = 50000, mrr_at_k=[10], ndcg_at_k=[10],
                 accuracy_at_k=[1, 3, 5, 10], precision_recall_at_k=[1, 3, 5, 10], map_at_k=[100],
                 show_progress_bar: bool = False, batch_size: int = 32, name: str = '', write_csv: bool = True,
                 score_functions={'cos_sim': cos_sim, 'dot_score': dot_score}, main_score_function: str = None,
                 log_callback=None):
        self.queries_ids = [qid for qid in queries if qid in relevant_docs and len(relevant_docs[qid]) > 0]
        self.queries = [queries[qid] for qid in self.queries_ids]
        self.corpus_ids = list(corpus.keys())
        self.corpus = [corpus[cid] for cid in self.corpus_ids]
        self.relevant_docs = relevant_docs
        self.corpus_chunk_size = corpus_chunk_size
        self.mrr_at_k, self.ndcg_at_k, self.accuracy_at_k = mrr_at_k, ndcg_at_k, accuracy_at_k
        self.precision_recall_at_k, self.map_at_k = precision_recall_at_k, map_at_k
        self.show_progress_bar, self.batch_size, self.name, self.write_csv = show_progress_bar, batch_size, name, write_csv
        self.score_functions = score_functions
        self.score_function_names = sorted(list(self.score_functions.keys()))
        self.main_score_function = main_score_function
        self.log_callback = log_callback

    def __call__(self, model, output_path: str = None, epoch: int = -1, steps: int = -1, *args, **kwargs) -> float:
        scores = self.compute_metrices(model, *args, **kwargs)
        if self.main_score_function is None:
            return max(scores[name]['map@k'][max(self.map_at_k)] for name in self.score_function_names)
        return scores[self.main_score_function]['map@k'][max(self.map_at_k)]

    def search_tensors(self, query_embeddings: torch.Tensor, corpus_embeddings: torch.Tensor, max_k: int, name: str):
        """Device-side scoring + top-``max_k`` for one score function -> (scores [Q,k], corpus rows [Q,k])."""
        sim = "cos_sim" if "cos" in name else "dot"
        return Ranker.dense_search_tensors(query_embeddings.cuda(), corpus_embeddings.cuda(), max_k, sim)

    def compute_metrices(self, model, corpus_model=None, corpus_embeddings: torch.Tensor = None):
        if corpus_model is None:
            corpus_model = model
        # max([]) raises for an empty accuracy_at_k exactly like the reference
        max_k = max(max(self.mrr_at_k), max(self.ndcg_at_k), max(self.accuracy_at_k), max(self.precision_recall_at_k),
                    max(self.map_at_k))
        kwargs = {'sentences': self.queries, 'show_progress_bar': self.show_progress_bar, 'batch_size': self.batch_size,
                  'convert_to_tensor': True}
        if 'query_mode' in inspect.signature(model.encode).parameters:
            kwargs.update({'query_mode': True})
        t0 = time.perf_counter()
        query_embeddings = model.encode(**kwargs)
        encoding_latency = ((time.perf_counter() - t0) / len(self.queries)) * 1000
        if corpus_embeddings is None:
            chunks = []
            for start in range(0, len(self.corpus), int(self.corpus_chunk_size)):
                kw = {'sentences': self.corpus[start:start + self.corpus_chunk_size], 'show_progress_bar': self.show_progress_bar,
                      'batch_size': 128, 'convert_to_tensor': True}
                if 'query_mode' in inspect.signature(model.encode).parameters:
                    kw.update({'query_mode': False})
                chunks.append(corpus_model.encode(**kw))
            corpus_embeddings = torch.cat(chunks, 0)
        queries_result_list = {}
        t0 = time.perf_counter()
        for name in self.score_functions:
            scores, rows = self.search_tensors(query_embeddings, corpus_embeddings, max_k, name)
            torch.cuda.synchronize()
            scoring_latency = ((time.perf_counter() - t0) / len(self.queries)) * 1000
            t1 = time.perf_counter()
            queries_result_list[name] = [[{'corpus_id': self.corpus_ids[r], 'score': s} for r, s in zip(rr, ss)]
                                         for rr, ss in zip(rows.cpu().tolist(), scores.cpu().tolist())]
            formatting_latency = ((time.perf_counter() - t1) / len(self.queries)) * 1000
        latency = encoding_latency + scoring_latency + formatting_latency
        if self.log_callback:
            self.log_callback(0, 0, 'latency (ms/q)', latency)
        logger.info(f"Avg. latency (ms/query): {latency:.2f} (Encoding: {encoding_latency:.2f}; Scoring: {scoring_latency:.2f}; Formatting: {formatting_latency:.2f})")
        self.last_results = queries_result_list
        return {name: self.compute_metrics(queries_result_list[name]) for name in self.score_functions}

    @staticmethod
    def compute_dcg_at_k(relevances, k):
        return sum(r / np.log2(i + 2) for i, r in enumerate(relevances[:k]))

    def compute_metrics(self, queries_result_list):
        """Accuracy / precision / recall / MRR / nDCG / MAP @k and R-precision with the reference's definitions (:395-485)."""
        acc = {k: 0 for k in self.accuracy_at_k}
        prec = {k: [] for k in self.precision_recall_at_k}
        rec = {k: [] for k in self.precision_recall_at_k}
        mrr = {k: 0 for k in self.mrr_at_k}
        ndcg = {k: [] for k in self.ndcg_at_k}
        avgp = {k: [] for k in self.map_at_k}
        rp = []
        for qi, hits in enumerate(queries_result_list):
            rel = self.relevant_docs[self.queries_ids[qi]]
            top = sorted(hits, key=lambda x: x["score"], reverse=True)
            flags = np.array([h["corpus_id"] in rel for h in top], dtype=bool)
            n_rel = len(rel)
            for k in acc:
                acc[k] += bool(flags[:k].any())
            for k in prec:
                c = int(flags[:k].sum())
                prec[k].append(c / k)
                rec[k].append(c / n_rel)
            for k in mrr:
                pos = np.flatnonzero(flags[:k])
                if len(pos):
                    mrr[k] += 1.0 / (pos[0] + 1)
            for k in ndcg:
                ndcg[k].append(self.compute_dcg_at_k(flags[:k].astype(int).tolist(), k) / self.compute_dcg_at_k([1] * n_rel, k))
            for k in avgp:
                f = flags[:k]
                cum = np.cumsum(f)
                avgp[k].append(float((cum[f] / (np.flatnonzero(f) + 1)).sum()) / min(k, n_rel))
            rp.append(int(flags[:n_rel].sum()) / n_rel)
        nq = len(self.queries)
        return {"accuracy@k": {k: v / nq for k, v in acc.items()}, "precision@k": {k: np.mean(v) for k, v in prec.items()},
                "recall@k": {k: np.mean(v) for k, v in rec.items()}, "ndcg@k": {k: np.mean(v) for k, v in ndcg.items()},
                "mrr@k": {k: v / nq for k, v in mrr.items()}, "map@k": {k: np.mean(v) for k, v in avgp.items()},
                "r-precision": np.mean(rp)}
