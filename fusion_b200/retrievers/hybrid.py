"""Drop-in mirror of ``src/retrievers/hybrid.py``: ``Ranker`` (:45-163) and ``Aggregator`` (:166-307).

Signatures, argument meaning, return shapes (``list[list[{'corpus_id', 'score'}]]``) and the reference's
observable quirks are kept: ``fuse`` slices the list of QUERIES with ``return_topk``
(hybrid.py:220), the Borda score of rank 0 is ``(n+1)/n`` (:249), the ``assert`` on the weight names is a no-op
(:195-197).  All scoring, normalisation, union-summing and sorting runs in the CUDA library through
``fusion_b200.ops``; the transformer encoders stay stock PyTorch and are passed in (or loaded by name when the
third-party packages are installed).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..index import DenseIndex, SparseIndex, TokenStore, sparse_queries
from .bm25 import BM25, FULL_RANKING_MIN_K

_INT_MIN = -(1 << 31)


def _lists_to_tensors(results: list[list[dict]], device):
    """list[list[{'corpus_id','score'}]] -> (ids int32 [Q, n], scores float64 [Q, n], lens int32 [Q])."""
    nq = len(results)
    n = max((len(r) for r in results), default=0)
    n = max(n, 1)
    ids = np.full((nq, n), -1, dtype=np.int64)
    sc = np.zeros((nq, n), dtype=np.float64)
    lens = np.zeros(nq, dtype=np.int32)
    for qi, r in enumerate(results):
        m = len(r)
        lens[qi] = m
        if m:
            ids[qi, :m] = [x['corpus_id'] for x in r]
            sc[qi, :m] = [x['score'] for x in r]
    return ids, sc, lens


def _tensors_to_lists(ids, scores, lens=None, cast=float):
    ids, scores = ids.cpu().tolist(), scores.cpu().tolist()
    lens = None if lens is None else lens.cpu().tolist()
    out = []
    for qi, (ri, rs) in enumerate(zip(ids, scores)):
        m = len(ri) if lens is None else lens[qi]
        out.append([{'corpus_id': i, 'score': cast(s)} for i, s in zip(ri[:m], rs[:m])])
    return out


class Ranker:
    """Ranking queries against a corpus (hybrid.py:45-163)."""

    @staticmethod
    def bm25_search(queries: list[str], corpus: dict[int, str], do_preprocessing: bool, k1: float, b: float,
                    return_topk: int = None, device: str = "cuda", preprocessor=None):
        """BM25 retrieval (hybrid.py:50-75): every document is ranked unless ``return_topk`` is given.

        ``do_preprocessing``: the reference lemmatises with spaCy ``fr_core_news_md`` through its ``TextPreprocessor``
        (src/data/preprocessor.py:15-39), a text pipeline outside the scoring path.  Pass an object with the same
        ``preprocess(list[str], lemmatize=True) -> list[str]`` method as ``preprocessor``; without one, preprocessing
        raises ImportError (as the reference does where spaCy is not installed)."""
        documents = list(corpus.values())
        idx2id = {i: pid for i, pid in enumerate(corpus.keys())}
        if do_preprocessing:
            if preprocessor is None:
                raise ImportError("do_preprocessing=True needs a text preprocessor (the reference's spaCy TextPreprocessor, "
                                  "src/data/preprocessor.py): pass it as `preprocessor=`")
            documents = preprocessor.preprocess(documents, lemmatize=True)
            queries = preprocessor.preprocess(queries, lemmatize=True)
        retriever = BM25(corpus=documents, k1=k1, b=b, device=device)
        scores, ids = retriever.search_all_tensors(queries, top_k=return_topk or len(documents))
        return [[{'corpus_id': idx2id.get(i), 'score': s} for i, s in zip(ri, rs)]
                for ri, rs in zip(ids.cpu().tolist(), scores.cpu().tolist())]

    @staticmethod
    def _load_single_vector_model(model_name_or_path):
        """An encoder object (anything with the reference's ``encode(sentences=, batch_size=, convert_to_tensor=,
        show_progress_bar=[, query_mode=])``) is used as is.  A name is loaded with stock PyTorch / Hugging Face code:
        SPLADE checkpoints through ``fusion_b200.retrievers.splade.SPLADE.from_pretrained`` (transformers masked-LM +
        the CUDA activation head), bi-encoders through sentence-transformers when that package is installed."""
        if not isinstance(model_name_or_path, str):
            return model_name_or_path
        if 'splade' in model_name_or_path.lower():
            from .splade.splade import SPLADE
            return SPLADE.from_pretrained(model_name_or_path, max_query_length=64, max_doc_length=512)
        from sentence_transformers import SentenceTransformer
        model = SentenceTransformer(model_name_or_path)
        model.max_seq_length = 512
        return model

    @staticmethod
    def dense_search_tensors(q_embs: torch.Tensor, d_embs: torch.Tensor, top_k: int, similarity: str = "cos_sim",
                             exact: bool = True, margin: float | None = None, doc_base: int = 0):
        """Embeddings in, (scores f32 [Q,k], ids int32 [Q,k]) out - the arithmetic of ``util.semantic_search``
        (hybrid.py:103).  Small k runs the tcgen05 filter GEMM (+ fp32 rescoring when ``exact``); a full ranking
        materialises the exact fp32 score matrix and sorts its rows."""
        index = DenseIndex.build(d_embs.float(), similarity, keep_f32=True, doc_base=doc_base)
        q32, q16 = index.prepare_queries(q_embs.float())
        return Ranker._dense_search_index(index, q32, q16, top_k, exact, margin)

    @staticmethod
    def _dense_search_index(index: DenseIndex, q32, q16, top_k: int, exact: bool = True, margin: float | None = None):
        n, dim = index.d_bf16.shape
        k = min(top_k, n)
        if k >= FULL_RANKING_MIN_K or 2 * k > ops.DEFAULT_CAP or dim % 64 != 0:
            out_s = torch.empty((q32.shape[0], k), dtype=torch.float32, device=q32.device)
            out_i = torch.empty((q32.shape[0], k), dtype=torch.int32, device=q32.device)
            step = max(1, (1 << 30) // (4 * n))
            for lo in range(0, q32.shape[0], step):
                sc = ops.dense_scores(q32[lo:lo + step], index.d_f32)
                out_s[lo:lo + step], out_i[lo:lo + step] = ops.rank_rows(sc, k, index.doc_base)
            return out_s, out_i
        if margin is None:
            margin = index.exact_margin(q32, q16) if exact else 0.0      # from the measured bf16 residual norms
        return ops.dense_topk(q16, index.d_bf16, q32 if exact else None, index.d_f32 if exact else None, k,
                              margin=margin, doc_base=index.doc_base)

    @staticmethod
    def single_vector_search(queries: list[str], corpus: dict[int, str], model_name_or_path, return_topk: int = None):
        """DPR / SPLADE retrieval (hybrid.py:78-106): encode with the stock model, score on the GPU library."""
        documents = list(corpus.values())
        idx2id = {i: pid for i, pid in enumerate(corpus.keys())}
        model = Ranker._load_single_vector_model(model_name_or_path)
        is_splade = model.__class__.__name__.upper().startswith('SPLADE')
        extra_d = {'query_mode': False} if is_splade else {}
        extra_q = {'query_mode': True} if is_splade else {}
        d_embs = model.encode(sentences=documents, batch_size=64, convert_to_tensor=True, show_progress_bar=True, **extra_d)
        q_embs = model.encode(sentences=queries, batch_size=64, convert_to_tensor=True, show_progress_bar=True, **extra_q)
        d_embs, q_embs = d_embs.cuda(), q_embs.cuda()
        top_k = return_topk or len(documents)
        if is_splade:
            scores, ids = Ranker.sparse_vector_search_tensors(q_embs, d_embs, top_k)
        else:
            scores, ids = Ranker.dense_search_tensors(q_embs, d_embs, top_k, "cos_sim")
        del model, d_embs, q_embs
        torch.cuda.empty_cache()
        return [[{'corpus_id': idx2id.get(i), 'score': s} for i, s in zip(ri, rs)]
                for ri, rs in zip(ids.cpu().tolist(), scores.cpu().tolist())]

    @staticmethod
    def sparse_vector_search_tensors(q_acts: torch.Tensor, d_acts: torch.Tensor, top_k: int, similarity: str = "cos_sim"):
        """SPLADE activations [*, V] (mostly zeros, splade.py:88-99) -> CSR -> inverted-index scoring."""
        from ..activations import activations_to_csr as to_csr         # csrc/activations.cu: ordered compaction per row
        dp, dt, dw = to_csr(d_acts.float().cuda())
        index = SparseIndex(dp, dt, dw, d_acts.shape[1], similarity, device=dp.device)
        qp, qt, qw = sparse_queries(*to_csr(q_acts.float().cuda()), similarity, dp.device)
        k = min(top_k, index.n_docs)
        if k >= FULL_RANKING_MIN_K or 2 * k > ops.DEFAULT_CAP:
            return ops.rank_rows(ops.sparse_scores(index.view(), qp, qt, qw), k, 0)
        return index.topk(qp, qt, qw, k)

    @staticmethod
    def multi_vector_search(queries: list[str], corpus: dict[int, str], model_name_or_path, output_dir: str = 'output',
                            return_topk: int = None):
        """ColBERT retrieval (hybrid.py:109-137).  ``model_name_or_path`` is an encoder object with
        ``encode_queries(list[str]) -> [Q, Lq, 128]`` and ``encode_docs(list[str]) -> (tok_ptr, tok_emb)``;
        every document is scored exhaustively with MaxSim (no PLAID candidate generation)."""
        documents = list(corpus.values())
        idx2id = {i: pid for i, pid in enumerate(corpus.keys())}
        if isinstance(model_name_or_path, str):
            raise ImportError("loading a ColBERT checkpoint by name needs colbert-ai; pass an encoder object")
        q_tok = model_name_or_path.encode_queries(queries).cuda()
        tok_ptr, tok_emb = model_name_or_path.encode_docs(documents)
        store = TokenStore(tok_ptr.cuda().to(torch.int64), tok_emb.cuda().to(torch.bfloat16))
        scores, ids = Ranker.maxsim_search_tensors(q_tok, store, return_topk or len(documents))
        return [[{'corpus_id': idx2id.get(i), 'score': s} for i, s in zip(ri, rs)]
                for ri, rs in zip(ids.cpu().tolist(), scores.cpu().tolist())]

    @staticmethod
    def maxsim_search_tensors(q_tok: torch.Tensor, store: TokenStore, top_k: int, cand_ids: torch.Tensor | None = None,
                              group=None, chunk_pairs: int = 1 << 24):
        """MaxSim-score candidates and rank them.  ``cand_ids`` None = EXHAUSTIVE search (``CustomSearcher.search_all`` without
        PLAID candidate generation, colbert_ir.py:245-255): every passage of the store is scored, in chunks of passages so
        that the [Q, chunk] score block stays bounded, and the chunks' top-k lists are merged on the device.  With a process
        ``group`` (None = the default group when torch.distributed is initialised, False = never) the store is this rank's shard
        of the collection: the shards' lists are merged over NCCL and every rank gets the global top-k of all queries."""
        from .. import sharding
        q16 = q_tok.to(torch.bfloat16).contiguous()
        nq = q16.shape[0]
        if cand_ids is not None:
            sc = ops.maxsim(q16, store.tok_ptr, None, cand_ids, store.doc_base, packed=store.packed())
            k = min(top_k, cand_ids.shape[1])
            order_s, order_i = ops.rank_rows(sc, k, 0)
            return order_s, torch.gather(cand_ids, 1, order_i.long())
        n = store.n_docs
        world = 1 if group is False else sharding._world(group)[0]       # group=False: this store is the whole collection
        n_total = n if world == 1 else sharding.allreduce_max_ints([n], q16.device, group)[0] * world    # (bound on the global size)
        k = min(top_k, n_total)
        step = max(1, chunk_pairs // max(nq, 1))
        parts_s, parts_i = [], []
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            ids = (torch.arange(lo, hi, dtype=torch.int32, device=q16.device) + store.doc_base).expand(nq, -1).contiguous()
            sc = ops.maxsim(q16, store.tok_ptr, None, ids, store.doc_base, packed=store.packed())
            kk = min(k, hi - lo)
            s2, i2 = ops.rank_rows(sc, kk, store.doc_base + lo)
            if kk < k:      # pad to a common width for the merge
                s2 = torch.cat([s2, torch.full((nq, k - kk), float("-inf"), device=s2.device)], 1)
                i2 = torch.cat([i2, torch.full((nq, k - kk), -1, dtype=torch.int32, device=i2.device)], 1)
            parts_s.append(s2)
            parts_i.append(i2)
        if not parts_s:
            parts_s = [torch.full((nq, k), float("-inf"), device=q16.device)]
            parts_i = [torch.full((nq, k), -1, dtype=torch.int32, device=q16.device)]
        sc, ids = (parts_s[0], parts_i[0]) if len(parts_s) == 1 else ops.merge_topk(torch.stack(parts_s), torch.stack(parts_i), k)
        if world > 1:
            sc, ids = sharding.gather_merge_topk(sc.contiguous(), ids.contiguous(), k, group)
        return sc, ids


def weight_grid(systems: list[str], step: float = 0.05) -> list[dict[str, float]]:
    """Every weight combination on a ``step`` grid that sums to one (hybrid.py:405-409)."""
    import itertools
    return [{name: float(w) for name, w in zip(systems, comb)}
            for comb in itertools.product(np.arange(0, 1 + step, step), repeat=len(systems)) if np.isclose(sum(comb), 1.0)]


def tune_linear_fusion_weights(results: dict[str, list[list[dict]]], labels: list[list[int]], normalization: str,
                               step: float = 0.05, percentile_distributions: dict | None = None,
                               device: str = "cuda", return_topk: int | None = 1000) -> list[dict]:
    """The linear-fusion weight sweep of ``hybrid.main`` (hybrid.py:404-426): one row per weight combination holding
    ``run_evaluation``'s metrics plus ``weight_<system>`` columns, in the reference's row order.  The reference fuses
    and evaluates once per combination (1,771 for four systems); here every system is normalised once and ONE kernel
    evaluates all combinations per query (``fz_fuse_sweep``).  The ranked lists must be in descending score order.

    ``return_topk``: the reference calls ``Aggregator.fuse`` with its default ``return_topk=1000``, which slices the list of
    QUERIES (hybrid.py:220), and ``run_evaluation`` zips predictions with labels - so with more than 1000 queries its
    sweep evaluates the first 1000 only.  The default reproduces that; ``None`` evaluates every query."""
    systems = list(results.keys())
    if return_topk is not None:
        results = {s: r[:return_topk] for s, r in results.items()}
        labels = labels[:return_topk]
    host = [_lists_to_tensors(results[s], device) for s in systems]
    lists = [(torch.from_numpy(h[0].astype(np.int32)).to(device), torch.from_numpy(h[1]).to(device),
              torch.from_numpy(h[2]).to(device)) for h in host]
    combos = weight_grid(systems, step)
    w = torch.tensor([[c[s] for s in systems] for c in combos], dtype=torch.float64, device=device)
    gp = np.zeros(len(labels) + 1, dtype=np.int32)
    np.cumsum([len(g) for g in labels], out=gp[1:])
    gi = np.fromiter((int(x) for g in labels for x in g), dtype=np.int64, count=int(gp[-1])).astype(np.int32)
    distrs = None
    if normalization in ('percentile-rank', 'normal-curve-equivalent'):
        distrs = [np.asarray(percentile_distributions.get(s), dtype=np.float64) for s in systems]
    vals = ops.fuse_sweep(lists, normalization, w, torch.from_numpy(gp).to(device), torch.from_numpy(gi).to(device),
                          distrs).cpu().numpy()
    names = ops.metric_names()
    return [{**dict(zip(names, row.tolist())), **{f'weight_{k}': v for k, v in c.items()}} for row, c in zip(vals, combos)]


class Aggregator:
    """Aggregating ranked lists (hybrid.py:166-307)."""

    @classmethod
    def fuse(cls, ranked_lists: dict[str, list[list[dict]]], method: str, normalization: str = None,
             linear_weights: dict[str, float] = None, percentile_distributions: dict[str, np.array] = None,
             return_topk: int = 1000, device: str = "cuda", numpy_promotion: str | None = None) -> list[dict[int, float]]:
        """Fuse the ranked lists of different retrieval systems (hybrid.py:170-220).

        ``numpy_promotion``: ``weight_scores`` multiplies the np.float32 normalised scores by the Python-float weights
        (hybrid.py:291) and ``aggregate_scores`` sums them.  Under NumPy >= 2 (NEP 50) that stays float32; under the
        NumPy 1.x the reference pins it promotes to float64, which can order near-tied fused scores differently.
        ``"nep50"`` / ``"legacy"`` select the behaviour; the default follows the NumPy installed next to this package, i.e.
        what the reference itself would compute in the same environment."""
        num_queries = len(next(iter(ranked_lists.values())))
        assert all(len(system_res) == num_queries for system_res in ranked_lists.values()), (
            "Ranked results from different retrieval systems have varying lenghts across systems (i.e., some systems have been run on more queries)."
        )
        if method not in ('bcf', 'rrf', 'nsf'):
            # the reference appends the untransformed lists when the method is unknown; sums of raw scores
            method, normalization, linear_weights = 'nsf', 'none', {s: 1.0 for s in ranked_lists}
        systems = list(ranked_lists.keys())
        if num_queries == 0:
            return []
        id_arrays, remap = [], None
        host = [_lists_to_tensors(ranked_lists[s], device) for s in systems]
        lo = min(int(h[0].min()) for h in host)
        hi = max(int(h[0].max()) for h in host)
        if lo < -1 or hi >= (1 << 31) - 1:
            # ids outside int32: fuse on dense surrogate ids and map back
            uniq = np.unique(np.concatenate([h[0].ravel() for h in host]))
            remap = uniq
            host = [(np.searchsorted(uniq, h[0]), h[1], h[2]) for h in host]
        lists = [(torch.from_numpy(h[0].astype(np.int32)).to(device), torch.from_numpy(h[1]).to(device),
                  torch.from_numpy(h[2]).to(device)) for h in host]
        weights = distrs = None
        if method == 'nsf':
            weights = [linear_weights[s] for s in systems]
            if normalization in ('percentile-rank', 'normal-curve-equivalent'):
                distrs = [np.asarray(percentile_distributions.get(s), dtype=np.float64) for s in systems]
        if numpy_promotion is None:
            numpy_promotion = "nep50" if type(np.float32(1) * 1.0) is np.float32 else "legacy"
        if numpy_promotion not in ("nep50", "legacy"):
            raise ValueError("numpy_promotion must be 'nep50' or 'legacy'")
        torch_norm = method == 'nsf' and normalization in ('min-max', 'z-score', 'arctan', 'percentile-rank', 'normal-curve-equivalent')
        ids, scores, lens = ops.fuse(lists, method, normalization, weights, distrs,
                                     promote_f64=torch_norm and numpy_promotion == "legacy")
        fp32 = torch_norm and numpy_promotion == "nep50"
        if remap is not None:
            ids = torch.from_numpy(remap[ids.cpu().numpy().clip(min=0)])
        final_results = _tensors_to_lists(ids, scores, lens, cast=np.float32 if fp32 else float)
        return final_results[:return_topk]

    @staticmethod
    def convert2dict(results: list[dict]) -> dict[int, float]:
        """list of {'corpus_id', 'score'} -> {corpus_id: score} (hybrid.py:223-233)."""
        return {res['corpus_id']: res['score'] for res in results}

    @staticmethod
    def _run_single(results: dict, method: str, normalization, distr, device="cuda"):
        ids = np.fromiter(results.keys(), dtype=np.int64, count=len(results))
        sc = np.array([float(v) for v in results.values()], dtype=np.float64)
        lists = [(torch.from_numpy(ids.astype(np.int32))[None].to(device), torch.from_numpy(sc)[None].to(device), None)]
        out_i, out_s, out_n = ops.fuse(lists, method, normalization, [1.0], None if distr is None else [distr])
        got = dict(zip(out_i[0].cpu().tolist(), out_s[0].cpu().tolist()))
        return ids, got

    @staticmethod
    def transform_scores(results: dict[int, float], transformation: str, percentile_distr: np.array = None) -> dict[int, float]:
        """Transform the scores of one result dict (hybrid.py:236-280), insertion order preserved."""
        if len(results) == 0:
            return results
        if transformation == 'borda-count':
            ids, got = Aggregator._run_single(results, 'bcf', None, None)
            return {int(i): got[int(i)] for i in ids}
        if transformation == 'reciprocal-rank':
            ids, got = Aggregator._run_single(results, 'rrf', None, None)
            return {int(i): got[int(i)] for i in ids}
        if transformation in ('min-max', 'z-score', 'arctan', 'percentile-rank', 'normal-curve-equivalent'):
            ids, got = Aggregator._run_single(results, 'nsf', transformation, percentile_distr)
            return {int(i): np.float32(got[int(i)]) for i in ids}
        return results

    @staticmethod
    def weight_scores(results: dict[int, float], w: float) -> dict[int, float]:
        """score * w (hybrid.py:283-291)."""
        return {corpus_id: score * w for corpus_id, score in results.items()}

    @staticmethod
    def aggregate_scores(*args: dict[int, float], device: str = "cuda") -> list[dict]:
        """Union-sum of result dicts, sorted by score descending, ties by first insertion (hybrid.py:294-307)."""
        if not args or all(len(a) == 0 for a in args):
            return []
        all_f32 = all(isinstance(v, np.float32) for a in args for v in a.values())
        lists = []
        for a in args:
            ids = np.fromiter(a.keys(), dtype=np.int64, count=len(a)).astype(np.int32)
            sc = np.array([float(v) for v in a.values()], dtype=np.float64)
            if len(a) == 0:
                ids, sc = np.full(1, -1, np.int32), np.zeros(1)
            lists.append((torch.from_numpy(ids)[None].to(device), torch.from_numpy(sc)[None].to(device),
                          torch.tensor([len(a)], dtype=torch.int32, device=device)))
        out_i, out_s, out_n = ops.fuse(lists, 'nsf', 'identity-f32' if all_f32 else 'none', [1.0] * len(args))
        n = int(out_n[0])
        cast = np.float32 if all_f32 else float
        return [{'corpus_id': i, 'score': cast(s)} for i, s in zip(out_i[0, :n].cpu().tolist(), out_s[0, :n].cpu().tolist())]
