#!/usr/bin/env python
"""Headline benchmark: hybrid BM25 + DPR + SPLADE + ColBERT top-1000 retrieval with rank fusion, queries/sec, on a
synthetic mMARCO-fr-shaped corpus (8,841,823 passages, 6,980 queries) - BASELINE.json config 5 - on N B200s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--docs D --queries Q]  (one rank per GPU
    under torchrun for N > 1).  Rank 0 prints ONE JSON line.

A "step" is one pass of the whole hot path over all queries.  `value` is measured with the queries resident in HBM;
`e2e` goes through HybridSearcher.search_host: pinned host query buffers in, fused top-1000 back to pinned host
memory, both copies inside the timed region.  Per-kernel durations come from CUDA events recorded on the launching
stream by the library's profiler (fz_profile_enable).  `--impl reference` / `cpu_baseline` time the oracle port of
the reference's CPU algorithms on the host cores over a bounded sample (the reference is pure Python and its tree is
not present on the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

N_DOCS, N_QUERIES, DIM, TOP_K = 8_841_823, 6_980, 768, 1000
BM25_VOCAB, SPLADE_VOCAB = 500_000, 32_005
COLBERT_POOL, COLBERT_LQ = 1_000_000, 64


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--docs", type=int, default=int(os.environ.get("FZ_BENCH_DOCS", N_DOCS)))
    p.add_argument("--queries", type=int, default=int(os.environ.get("FZ_BENCH_QUERIES", N_QUERIES)))
    p.add_argument("--pool", type=int, default=int(os.environ.get("FZ_BENCH_POOL", COLBERT_POOL)))
    p.add_argument("--systems", default=os.environ.get("FZ_BENCH_SYSTEMS", "bm25,dpr,splade,colbert"))
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--dense-mode", default=os.environ.get("FZ_BENCH_DENSE_MODE", "exact"), choices=["exact", "bf16"],
                   help="DPR: bf16 tensor-core filter + exact fp32 rescoring (default) or the bf16 throughput mode (config 2)")
    p.add_argument("--parity-queries", type=int, default=int(os.environ.get("FZ_BENCH_PARITY", 16)),
                   help="queries checked against an independent exhaustive computation after the timed region (0 = off)")
    p.add_argument("--ncu-range", action="store_true",
                   help="bracket ONE extra step with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ device-side synthetic data
def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def zipf_cdf(vocab, s, device):
    c = torch.cumsum(torch.arange(1, vocab + 1, device=device, dtype=torch.float64) ** (-s), 0)
    return (c / c[-1]).float()


def zipf_draw(cdf, n, gen):
    u = torch.rand(n, device=cdf.device, generator=gen)
    return torch.searchsorted(cdf, u).clamp_(max=cdf.numel() - 1)


GEN_CHUNK = 500_000      # docs per generation chunk: chunk c of the GLOBAL corpus is seeded by (seed, c) whatever the sharding


def _chunks(lo, hi, n_total):
    """(chunk index, first doc of the chunk, docs in the chunk, slice of [lo, hi) inside it)"""
    for c in range(lo // GEN_CHUNK, (hi + GEN_CHUNK - 1) // GEN_CHUNK):
        c0 = c * GEN_CHUNK
        m = min(GEN_CHUNK, n_total - c0)
        yield c, c0, m, max(lo, c0) - c0, min(hi, c0 + m) - c0


def make_lexical(lo, hi, n_total, seed, device):
    """Docs [lo, hi) of the global corpus: doc lengths clip(lognormal(3.3, 0.5), 3, 256), Zipf(1.07) over 500k terms (C3)."""
    cdf = zipf_cdf(BM25_VOCAB, 1.07, device)
    lens_all, toks_all = [], []
    for c, c0, m, a, b in _chunks(lo, hi, n_total):
        g = _gen(device, seed * 100003 + c)
        lens = torch.exp(torch.randn(m, device=device, generator=g) * 0.5 + 3.3).clamp_(3, 256).long()
        ptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
        ptr[1:] = torch.cumsum(lens, 0)
        toks = zipf_draw(cdf, int(ptr[-1]), g).to(torch.int32)
        lens_all.append(lens[a:b])
        toks_all.append(toks[int(ptr[a]):int(ptr[b])])
    lens = torch.cat(lens_all)
    ptr = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    return ptr, torch.cat(toks_all)


def make_lexical_queries(nq, seed, device):
    g = _gen(device, seed)
    lens = 1 + torch.poisson(torch.full((nq,), 4.0, device=device), generator=g).long()
    ptr = torch.zeros(nq + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    toks = zipf_draw(zipf_cdf(BM25_VOCAB, 1.07, device), int(ptr[-1]), g).to(torch.int32)
    return ptr.to(torch.int32), toks


def _splade_chunk(m, mean_nnz, lo_nnz, hi_nnz, g, cdf, device):
    want = torch.poisson(torch.full((m,), float(mean_nnz), device=device), generator=g).clamp_(lo_nnz, hi_nnz).long()
    over = want * 2 + 8
    optr = torch.zeros(m + 1, dtype=torch.int64, device=device)
    optr[1:] = torch.cumsum(over, 0)
    row = torch.repeat_interleave(torch.arange(m, device=device), over)
    key = torch.unique(row * SPLADE_VOCAB + zipf_draw(cdf, int(optr[-1]), g))
    del row
    urow, uterm = key // SPLADE_VOCAB, key % SPLADE_VOCAB
    prio = torch.rand(key.numel(), device=device, generator=g)
    order = torch.argsort(urow.double() + prio.double())            # by (row, random priority)
    cnt = torch.bincount(urow, minlength=m)
    start = torch.zeros(m + 1, dtype=torch.int64, device=device)
    start[1:] = torch.cumsum(cnt, 0)
    rank = torch.arange(key.numel(), device=device) - start[urow[order]]
    keep = torch.zeros(key.numel(), dtype=torch.bool, device=device)
    keep[order] = rank < want[urow[order]]
    w = torch.log1p(torch.relu(torch.randn(key.numel(), device=device, generator=g) * 0.7 + 0.5))
    keep &= w > 0
    urow, uterm, w = urow[keep], uterm[keep], w[keep]
    p = torch.zeros(m + 1, dtype=torch.int64, device=device)
    p[1:] = torch.cumsum(torch.bincount(urow, minlength=m), 0)
    return p, uterm.to(torch.int32), w.float()


def make_splade(lo, hi, n_total, mean_nnz, lo_nnz, hi_nnz, seed, device):
    """Rows [lo, hi) of the global matrix of CSR sparse vectors: ~Poisson(mean_nnz) distinct Zipf(1.05) terms per row,
    weights log1p(relu(N(0.5, 0.7))) > 0 (C3)."""
    cdf = zipf_cdf(SPLADE_VOCAB, 1.05, device)
    lens_all, terms, ws = [], [], []
    for c, c0, m, a, b in _chunks(lo, hi, n_total):
        p, t, w = _splade_chunk(m, mean_nnz, lo_nnz, hi_nnz, _gen(device, seed * 100003 + c), cdf, device)
        lens_all.append(p[a + 1:b + 1] - p[a:b])
        terms.append(t[int(p[a]):int(p[b])])
        ws.append(w[int(p[a]):int(p[b])])
    lens = torch.cat(lens_all)
    ptr = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    return ptr, torch.cat(terms), torch.cat(ws)


def make_dense_index(lo, hi, n_total, dim, seed, device):
    from fusion_b200 import ops
    from fusion_b200.index import DenseIndex
    d32 = torch.empty((hi - lo, dim), dtype=torch.float32, device=device)
    d16 = torch.empty((hi - lo, dim), dtype=torch.bfloat16, device=device)
    pos = 0
    for c, c0, m, a, b in _chunks(lo, hi, n_total):
        x = torch.randn((m, dim), device=device, generator=_gen(device, seed * 100003 + c))[a:b]
        f, h = ops.normalize_rows(x)
        d32[pos:pos + b - a], d16[pos:pos + b - a] = f, h
        pos += b - a
    return DenseIndex(d32, d16, "cos_sim", lo)


def make_token_store(lo, hi, n_total, seed, device):
    """Passages [lo, hi) of the global ColBERT token store: clip(Poisson(70), 8, 180) unit bf16 token vectors each (C4)."""
    from fusion_b200.index import TokenStore
    lens_all, embs = [], []
    for c, c0, m, a, b in _chunks(lo, hi, n_total):
        g = _gen(device, seed * 100003 + c)
        lens = torch.poisson(torch.full((m,), 70.0, device=device), generator=g).clamp_(8, 180).long()
        ptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
        ptr[1:] = torch.cumsum(lens, 0)
        x = torch.randn((int(ptr[-1]), 128), device=device, generator=g)
        x = x[int(ptr[a]):int(ptr[b])]
        lens_all.append(lens[a:b])
        embs.append((x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16))
        del x
    lens = torch.cat(lens_all)
    ptr = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    return TokenStore(ptr, torch.cat(embs), lo)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.rows, self.stop, self.dev = [], threading.Event(), dev_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_reference_sample(n_docs_full, n_queries_full, systems, seconds_budget=25.0):
    """Time the oracle port of the reference's CPU algorithms on a bounded sample and extrapolate linearly to the
    full workload (every loop is exactly O(N) in the corpus and O(Q) in the queries).  -> dict with per-system
    seconds per query at full corpus size and the hybrid queries/sec."""
    from fusion_b200 import synth
    from oracle import bm25 as obm25, dense as odense, fusion as ofusion, maxsim as omaxsim
    torch.set_num_threads(os.cpu_count() or 1)
    per_q, sample = {}, {}
    if "bm25" in systems:
        n = 200_000
        (dptr, dtok), (qptr, qtok) = synth.c3_lexical(n, 16, BM25_VOCAB)
        o = obm25.LexicalOracle(dptr, dtok, BM25_VOCAB, "bm25", 0.9, 0.4)
        t0 = time.perf_counter()
        for qi in range(16):
            t = qtok[qptr[qi]:qptr[qi + 1]]
            o.search_ids(np.where(t < BM25_VOCAB, t, -1), TOP_K)
        per_q["bm25"] = (time.perf_counter() - t0) / 16 * (n_docs_full / n)
        sample["bm25"] = f"16 queries x {n} docs (numpy port of bm25.py:100-156), scaled x{n_docs_full / n:.1f} in N"
    if "dpr" in systems:
        n, nq = 500_000, 64
        d = torch.nn.functional.normalize(torch.randn(n, DIM), dim=1)
        q = torch.randn(nq, DIM)
        t0 = time.perf_counter()
        odense.topk_tensors(q, d, TOP_K, "cos_sim", chunk=50_000)
        per_q["dpr"] = (time.perf_counter() - t0) / nq * (n_docs_full / n)
        sample["dpr"] = f"{nq} queries x {n} docs x {DIM} in 50k-doc chunks (torch CPU sgemm + topk), scaled x{n_docs_full / n:.1f}"
        del d
    if "splade" in systems:
        n, nq = 20_000, 32
        dp, dt, dw = synth.splade_vectors(n, SPLADE_VOCAB, 120, 8, 512, seed=311)
        qp, qt, qw = synth.splade_vectors(nq, SPLADE_VOCAB, 24, 2, 64, seed=312)
        dd = torch.from_numpy(synth.densify(dp, dt, dw, SPLADE_VOCAB))
        qd = torch.from_numpy(synth.densify(qp, qt, qw, SPLADE_VOCAB))
        t0 = time.perf_counter()
        odense.topk_tensors(qd, dd, TOP_K, "cos_sim", chunk=10_000)
        per_q["splade"] = (time.perf_counter() - t0) / nq * (n_docs_full / n)
        sample["splade"] = f"{nq} queries x {n} docs, dense [.,{SPLADE_VOCAB}] cosine as the reference scores SPLADE (hybrid.py:101-103), scaled x{n_docs_full / n:.0f}"
        del dd
    if "colbert" in systems:
        ptr, emb = synth.colbert_tokens(4000, 128, 70, 8, 180, seed=401)
        q = synth.colbert_queries(2, COLBERT_LQ, 128, seed=402)
        cand = torch.from_numpy(np.random.default_rng(0).integers(0, 4000, (2, TOP_K)).astype(np.int32))
        t0 = time.perf_counter()
        omaxsim.maxsim_scores(torch.from_numpy(q), torch.from_numpy(ptr), torch.from_numpy(emb), cand)
        per_q["colbert"] = (time.perf_counter() - t0) / 2
        sample["colbert"] = "2 queries x 1000 candidates x ~70 tokens (torch restatement of colbert_score)"
    rng = np.random.default_rng(1)
    n_sys = len(systems)
    ids = [np.stack([rng.choice(4 * TOP_K, TOP_K, replace=False) for _ in range(16)]) for _ in range(n_sys)]
    sc = [-np.sort(-rng.normal(0, 1, (16, TOP_K)), axis=1) for _ in range(n_sys)]
    t0 = time.perf_counter()
    for qi in range(16):
        ofusion.fuse_query([i[qi] for i in ids], [s[qi] for s in sc], "nsf", "z-score", [1.0 / n_sys] * n_sys)
        ofusion.fuse_query([i[qi] for i in ids], [s[qi] for s in sc], "rrf")
    per_q["fusion"] = (time.perf_counter() - t0) / 16
    sample["fusion"] = "16 queries x (nsf z-score + rrf) over the systems' top-1000 lists (port of hybrid.py:170-307)"
    total = sum(per_q.values())
    return {"sec_per_query": per_q, "qps": 1.0 / total, "sample": sample}


def workload_name(systems, n_docs, n_queries):
    return (f"C5 hybrid {'+'.join(systems)} top-{TOP_K} + nsf z-score and rrf fusion, {n_docs} docs, {n_queries} queries, d={DIM}")


def run_reference(args):
    """The reference arm: the reference's CPU algorithms (oracle port) on the host cores.  A step is one bounded sample of the
    workload (a few queries against a corpus slice, per system), extrapolated linearly to the full workload; W warm-up and K
    timed steps like the GPU arm."""
    systems = args.systems.split(",")
    t0 = time.perf_counter()
    for _ in range(max(0, min(args.warmup, 1))):            # one warm-up sample pages in torch / numpy; more would only repeat it
        cpu_reference_sample(args.docs, args.queries, systems)
    runs = [cpu_reference_sample(args.docs, args.queries, systems) for _ in range(max(1, min(args.steps, 3)))]
    wall = time.perf_counter() - t0
    per_q = {k: float(np.mean([r["sec_per_query"][k] for r in runs])) for k in runs[0]["sec_per_query"]}
    qps = 1.0 / sum(per_q.values())
    line = {
        "impl": "reference", "metric": "hybrid top-1000 queries/sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.queries / qps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
        "config": {"workload": workload_name(systems, args.docs, args.queries), "docs": args.docs, "queries": args.queries,
                   "sampled": f"{len(runs)} timed sample(s) of the workload, extrapolated linearly in docs and queries"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "; ".join(f"{k}: {v}" for k, v in runs[0]["sample"].items()),
                         "sec_per_query": per_q, "sample_wall_s": wall},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ parity at benchmark scale
def parity_block(world, rank, dev, nq, lexical, sparse, dense, tokens, q, lists, fused_nsf, fused_rrf, pool, dense_exact,
                 n_sample=16):
    """Outside the timed region: for ``n_sample`` queries, every retriever's top-1000 at the benchmark's corpus size against
    an INDEPENDENT exhaustive computation in plain torch on the device (no fusion_b200 kernel): BM25 in fp64 in query-token
    order from the CSR arrays (bit-exact ids and scores expected), DPR as an fp32 matmul, SPLADE as an fp32 scatter-add over
    the doc-major postings, MaxSim as fp32 matmuls over the candidates' token rows; the fused lists against the CPU
    restatement of Aggregator.fuse (oracle/fusion.py, used as the checker).  Sharded runs compute per-shard exact top-k
    and merge them."""
    import torch.distributed as dist
    from fusion_b200 import sharding
    per = (nq + world - 1) // world
    sample = sorted(set(int(x) for x in np.linspace(0, nq - 1, n_sample)))
    out = {"queries": len(sample)}

    def owner(qi):
        return qi // per, qi % per          # (rank that holds the query's merged list, row in its slice)

    def merge_exact(sc, ids):
        """[S, k] local exact top-k (global ids) -> global [S, k]: score desc, ties by lower id."""
        if world > 1:
            gs = [torch.empty_like(sc) for _ in range(world)]
            gi = [torch.empty_like(ids) for _ in range(world)]
            dist.all_gather(gs, sc.contiguous())
            dist.all_gather(gi, ids.contiguous())
            sc, ids = torch.cat(gs, 1), torch.cat(gi, 1)
        o = torch.sort(ids, dim=1, stable=True).indices
        sc, ids = torch.gather(sc, 1, o), torch.gather(ids, 1, o)
        o = torch.sort(sc, dim=1, descending=True, stable=True).indices[:, :TOP_K]
        return torch.gather(sc, 1, o), torch.gather(ids, 1, o)

    def local_topk(score, base):
        k = min(TOP_K, score.numel())
        o = torch.sort(score, descending=True, stable=True).indices[:k]
        sc, ids = score[o], (o + base).to(torch.int64)
        if k < TOP_K:
            sc = torch.cat([sc, torch.full((TOP_K - k,), float("-inf"), dtype=sc.dtype, device=dev)])
            ids = torch.cat([ids, torch.full((TOP_K - k,), 1 << 40, dtype=torch.int64, device=dev)])
        return sc, ids

    def mine(name, qi):
        r, row = owner(qi)
        if r != rank:
            return None
        return lists[name][0][row], lists[name][1][row].to(torch.int64)

    def reduce_flags(d):
        """AND / MAX over the ranks (each sampled query is checked by the rank that owns it)."""
        keys = sorted(d)
        t = torch.tensor([float(d[k]) for k in keys], dtype=torch.float64, device=dev)
        if world > 1:
            neg = torch.tensor([k.endswith("_exact") or k.endswith("_min") for k in keys], device=dev)
            t = torch.where(neg, -t, t)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t = torch.where(neg, -t, t)
        return {k: (bool(v) if k.endswith("_exact") else float(v)) for k, v in zip(keys, t.tolist())}

    res = {}
    if lexical is not None:
        ix = lexical
        ptr_h = q.lex_ptr.cpu().tolist()
        k1, b, avgdl = float(ix.k1), float(ix.b), float(ix.avgdl)
        exact, smax = True, 0.0
        for qi in sample:
            score = torch.zeros(ix.n_docs, dtype=torch.float64, device=dev)
            for t in q.lex_term[ptr_h[qi]:ptr_h[qi + 1]].tolist():         # query-token order, duplicates counted (bm25.py:152-155)
                if t < 0:
                    continue
                lo, hi = int(ix.term_ptr[t]), int(ix.term_ptr[t + 1])
                docs = ix.post_doc[lo:hi].long()
                tf = ix.post_tf[lo:hi].double()
                dl = ix.doc_len[docs].double()
                num = ix.idf[t] * (tf * (k1 + 1))
                den = tf + k1 * ((1 - b) + (b * dl) / avgdl)
                score[docs] += num / den
            esc, eid = merge_exact(*[x[None] for x in local_topk(score, ix.doc_base)])
            got = mine("bm25", qi)
            if got is not None:
                exact &= bool(torch.equal(got[1], eid[0])) and bool(torch.equal(got[0], esc[0]))
                smax = max(smax, float((got[0] - esc[0]).abs().max()))
        res["bm25_exact"] = exact
        res["bm25_max_abs"] = smax
    if dense is not None:
        qs = torch.nn.functional.normalize(q.dense[sample].float(), dim=1)
        sc = qs @ dense.d_f32.t()
        lsc, lid = zip(*[local_topk(sc[i], dense.doc_base) for i in range(len(sample))])
        esc, eid = merge_exact(torch.stack(lsc), torch.stack(lid))
        del sc
        mx, ov_hit, ov_tot = 0.0, 0, 0
        for j, qi in enumerate(sample):
            got = mine("dpr", qi)
            if got is None:
                continue
            cut = float(esc[j, -1])
            tie = 1e-6 if dense_exact else 2e-3 * abs(cut)
            want = {int(i) for i, v in zip(eid[j].tolist(), esc[j].tolist()) if v > cut + tie}
            have = dict(zip(got[1].tolist(), got[0].tolist()))
            ov_hit += len(want & set(have))
            ov_tot += len(want)
            ref = dict(zip(eid[j].tolist(), esc[j].tolist()))
            mx = max([mx] + [abs(v - ref[i]) for i, v in have.items() if i in ref])
        res["dpr_max_abs"] = mx
        res["dpr_overlap_min"] = ov_hit / ov_tot if ov_tot else 1.0
    if sparse is not None:
        sp_ptr = q.sp_ptr.cpu().tolist()
        lens = sparse.doc_ptr[1:] - sparse.doc_ptr[:-1]
        row = torch.repeat_interleave(torch.arange(sparse.n_docs, device=dev), lens)
        term = sparse.doc_post[:, 0].long()
        w = sparse.doc_post[:, 1].contiguous().view(torch.float32)
        mx, ov_hit, ov_tot = 0.0, 0, 0
        for qi in sample:
            qd = torch.zeros(SPLADE_VOCAB, dtype=torch.float32, device=dev)
            qd[q.sp_term[sp_ptr[qi]:sp_ptr[qi + 1]].long()] = q.sp_weight[sp_ptr[qi]:sp_ptr[qi + 1]]
            score = torch.zeros(sparse.n_docs, dtype=torch.float32, device=dev).index_add_(0, row, w * qd[term])
            esc, eid = merge_exact(*[x[None] for x in local_topk(score, sparse.doc_base)])
            got = mine("splade", qi)
            if got is None:
                continue
            cut = float(esc[0, -1])
            want = {int(i) for i, v in zip(eid[0].tolist(), esc[0].tolist()) if v > cut + 1e-5}
            have = dict(zip(got[1].tolist(), got[0].tolist()))
            ov_hit += len(want & set(have))
            ov_tot += len(want)
            ref = dict(zip(eid[0].tolist(), esc[0].tolist()))
            mx = max([mx] + [abs(v - ref[i]) for i, v in have.items() if i in ref])
        del row, term, w
        res["splade_max_abs"] = mx
        res["splade_overlap_min"] = ov_hit / ov_tot if ov_tot else 1.0
    if tokens is not None and tokens.tok_emb is not None and "colbert" in lists:
        mx = 0.0
        for qi in sample:
            r, rowi = owner(qi)
            cand = lists["colbert"][1][rowi].clone() if r == rank else torch.empty(TOP_K, dtype=torch.int32, device=dev)
            have = lists["colbert"][0][rowi].clone() if r == rank else torch.empty(TOP_K, dtype=torch.float32, device=dev)
            if world > 1:
                dist.broadcast(cand, src=r)
            qt = q.colbert[qi].float()                                                  # [Lq, 128]
            part = torch.zeros(TOP_K, dtype=torch.float32, device=dev)
            pid = torch.where(cand >= 0, cand.long() % pool, cand.long()) if pool else cand.long()
            own = (pid >= tokens.doc_base) & (pid < tokens.doc_base + tokens.n_docs)
            for j in torch.nonzero(own).flatten().tolist():
                d = int(pid[j]) - tokens.doc_base
                e = tokens.tok_emb[int(tokens.tok_ptr[d]):int(tokens.tok_ptr[d + 1])].float()
                part[j] = (e @ qt.t()).max(dim=0).values.sum()
            if world > 1:
                dist.all_reduce(part)
            if r == rank:
                ok = cand >= 0
                mx = max(mx, float((part[ok] - have[ok]).abs().max()))
        res["colbert_max_abs"] = mx
    # ---- fusion: the kernel's output against the CPU restatement of Aggregator.fuse on the SAME input lists
    from oracle import fusion as ofusion
    names = list(lists.keys())
    ok_rrf = True
    smax, nsf_equal, nsf_total, nsf_gap = 0.0, 0, 0, 0.0
    for qi in sample:
        r, rowi = owner(qi)
        if r != rank:
            continue
        ids = [lists[n][1][rowi].cpu().numpy() for n in names]
        scs = [lists[n][0][rowi].double().cpu().numpy() for n in names]
        keep = [i >= 0 for i in ids]
        ids = [i[k_] for i, k_ in zip(ids, keep)]
        scs = [v[k_] for v, k_ in zip(scs, keep)]
        for tag, fused, kw in (("nsf", fused_nsf, dict(method="nsf", normalization="z-score", weights=[1.0 / len(names)] * len(names))),
                               ("rrf", fused_rrf, dict(method="rrf"))):
            eids, esc = ofusion.fuse_query(ids, scs, **kw)
            n = min(TOP_K, len(eids))
            got = fused[0][rowi, :n].cpu().tolist()
            smax = max(smax, float(np.abs(fused[1][rowi, :n].cpu().numpy() - np.asarray(esc[:n], dtype=np.float64)).max()))
            if tag == "rrf":            # Python-float arithmetic on integer ranks: the id sequence must be exact
                ok_rrf &= got == eids[:n]
            else:
                # z-score runs in fp32 (torch on the reference side): mean / std reduce in a different order, so fused scores
                # that differ by an ulp or two may swap.  Report how many positions agree and the largest oracle-score gap
                # between a swapped pair (a real ordering error would show as a gap far above fp32 resolution).
                ref_sc = dict(zip(eids, esc))
                nsf_total += n
                for a, b in zip(got, eids[:n]):
                    if a == b:
                        nsf_equal += 1
                    else:
                        nsf_gap = max(nsf_gap, abs(float(ref_sc.get(a, float("inf"))) - float(ref_sc[b])))
    res["fuse_rrf_ids_exact"], res["fuse_max_abs"] = ok_rrf, smax
    res["fuse_nsf_ids_equal_min"] = nsf_equal / nsf_total if nsf_total else 1.0
    res["fuse_nsf_swapped_pairs_max_gap"] = nsf_gap
    out.update(reduce_flags(res))
    return out


# ------------------------------------------------------------------------------------------ main (ours)
def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    import torch.distributed as dist
    from fusion_b200 import _lib, ops, sharding
    from fusion_b200.hybrid_engine import HybridQueries, HybridSearcher
    from fusion_b200.index import LexicalIndex, SparseIndex, sparse_queries

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    systems = args.systems.split(",")
    nq, n_total = args.queries, args.docs
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    n_local = hi - lo
    lib = _lib.load()
    dense_exact = args.dense_mode == "exact"
    # the ColBERT token store: the whole corpus when its shard fits next to the other indexes (~18 KB per passage), else a pool
    # of `--pool` passages that candidate ids are mapped into (1 GPU: 8.8M passages would need 158 GB)
    pool = 0 if (args.pool <= 0 or args.pool >= n_total) else args.pool
    if "colbert" in systems and args.pool == COLBERT_POOL and n_total / world * 17.9e3 * (2 if args.parity_queries else 1) < 90e9:
        pool = 0
    tok_total = pool or n_total

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    lexical = sparse = dense = tokens = None
    q = HybridQueries()
    algo, build_s, gen_s = {}, {}, {}
    t_setup = time.perf_counter()
    if "bm25" in systems:
        (dptr, dtok), gen_s["bm25"] = timed(lambda: make_lexical(lo, hi, n_total, 301, dev))
        lexical, build_s["bm25"] = timed(lambda: LexicalIndex(
            dptr, dtok, BM25_VOCAB, "bm25", 0.9, 0.4, device=dev, doc_base=lo,
            stats_reduce=lambda n, df, sdl: sharding.allreduce_lexical_stats(n, df, sdl, dev)))
        del dptr, dtok
        q.lex_ptr, q.lex_term = make_lexical_queries(nq, 302, dev)
        # algorithmic bytes: per query sum over unique terms of df_t * 8 B + k * 8 B
        ptr_h, term_h = q.lex_ptr.cpu().numpy(), q.lex_term.cpu().numpy()
        df_loc = np.diff(lexical.term_ptr.cpu().numpy())
        algo["bm25_bytes"] = float(sum(df_loc[np.unique(term_h[ptr_h[i]:ptr_h[i + 1]])].sum() for i in range(nq)) * 8 + nq * TOP_K * 8)
        # the same lists read ONCE for the whole query batch (the second denominator for kernels that share
        # posting reads between queries): sum over the batch's distinct terms
        algo["bm25_union_bytes"] = float(df_loc[np.unique(term_h[term_h >= 0])].sum()) * 8 + nq * TOP_K * 8
        algo["bm25_index_bytes"] = lexical.nbytes()
    if "splade" in systems:
        (dp, dt, dw), gen_s["splade"] = timed(lambda: make_splade(lo, hi, n_total, 120, 8, 512, 311, dev))
        # threshold bootstrap over the first docs of every shard: the SAME number on every rank (the round schedule, hence the
        # collectives of the cross-shard threshold exchange, start there), at most 1/8 of the smallest shard
        boot = min(262144, (n_total // world) // 8 // 256 * 256)
        boot = int(os.environ.get("FZ_BENCH_BOOT", boot))      # tuning knob (same value on every rank)
        sparse, build_s["splade"] = timed(lambda: SparseIndex(dp, dt, dw, SPLADE_VOCAB, "cos_sim", device=dev, doc_base=lo,
                                                              boot_docs=boot))
        del dp, dt, dw
        qp, qt, qw = make_splade(0, nq, nq, 24, 2, 64, 312, dev)
        q.sp_ptr, q.sp_term, q.sp_weight = sparse_queries(qp, qt, qw, "cos_sim", dev)
        ptr_h, term_h = q.sp_ptr.cpu().numpy(), q.sp_term.cpu().numpy()
        df_loc = torch.bincount(sparse.doc_post[:, 0].long(), minlength=SPLADE_VOCAB).cpu().numpy()
        algo["splade_bytes"] = float(sum(df_loc[term_h[ptr_h[i]:ptr_h[i + 1]]].sum() for i in range(nq)) * 8 + nq * TOP_K * 8)
        algo["splade_union_bytes"] = float(df_loc[np.unique(term_h[term_h >= 0])].sum()) * 8 + nq * TOP_K * 8
        algo["splade_index_bytes"] = sparse.nbytes()
        if sparse.head is not None:
            is_head = (sparse.head.term_head >= 0).cpu().numpy()
            algo["splade_head_dim"] = sparse.head.head_dim
            algo["splade_head_flops"] = 2.0 * nq * n_local * sparse.head.head_dim
            # tail: every (query tail term, posting) pair reads one 6-byte posting (uint16 offset + fp32 weight) and the pass
            # writes one 4-bit code per (query, doc)
            algo["splade_tail_bytes"] = float(sum(df_loc[term_h[ptr_h[i]:ptr_h[i + 1]]][~is_head[term_h[ptr_h[i]:ptr_h[i + 1]]]].sum()
                                                  for i in range(nq))) * 6 + nq * n_local * 0.5
    if "dpr" in systems:
        dense, t = timed(lambda: make_dense_index(lo, hi, n_total, DIM, 201, dev))
        gen_s["dpr"], build_s["dpr"] = t, 0.0            # (normalisation + bf16 copy happen while generating)
        if not dense_exact:
            dense.d_f32 = dense.d_f32 if args.parity_queries else None
        q.dense = torch.randn((nq, DIM), device=dev, generator=_gen(dev, 202))
        algo["dpr_flops"] = 2.0 * nq * n_local * DIM
    if "colbert" in systems:
        plo, phi = sharding.shard_bounds(tok_total, world, rank)
        tokens, gen_s["colbert"] = timed(lambda: make_token_store(plo, phi, tok_total, 401, dev))
        x = torch.randn((nq, COLBERT_LQ, 128), device=dev, generator=_gen(dev, 402))
        q.colbert = (x / x.norm(dim=2, keepdim=True)).to(torch.bfloat16)
        algo["colbert_avg_tokens"] = float(tokens.n_tokens) / max(1, tokens.n_docs)
        # the kernel streams the packed image; the plain matrix is only kept for the parity block's torch recomputation
        _, build_s["colbert"] = timed(lambda: tokens.packed(drop_plain=not args.parity_queries))
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    searcher = HybridSearcher(lexical, sparse, dense, tokens, k=TOP_K, fusion="nsf", normalization="z-score",
                              colbert_pool=(pool or None) if tokens is not None else None, dense_exact=dense_exact)
    searcher_rrf = HybridSearcher(k=TOP_K, fusion="rrf")

    def step(qq):
        lists = searcher.retrieve(qq)
        f1 = searcher._timed("fuse_nsf_zscore", lambda: searcher.fuse(lists))
        f2 = searcher._timed("fuse_rrf", lambda: searcher_rrf.fuse(lists))
        return lists, f1, f2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(q)
    if args.ncu_range:          # profiler capture range: exactly one step of the hot path, after warm-up
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step(q)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    # ---- device-resident timing
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step(q)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps

    # ---- per-kernel profile pass (CUDA events on the launching stream around every kernel launch)
    lib.fz_profile_enable(1)
    searcher.timing = True
    searcher.stage_ms = {}
    last = step(q)
    searcher.collect_stage_ms()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.fz_profile_summary(buf, len(buf))
    lib.fz_profile_enable(0)
    searcher.timing = False
    kern = {}
    for line in buf.value.decode().splitlines():
        name, cnt, tot = line.split()
        kern[name] = {"launches": int(cnt), "ms": float(tot)}
    gpu_launches = sum(v["launches"] for v in kern.values())
    stage = dict(searcher.stage_ms)
    if world > 1:           # the slowest rank's stage times (what the step waits for)
        keys = sorted(stage)
        tt = torch.tensor([stage[k] for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stage = dict(zip(keys, tt.tolist()))

    # ---- parity at this corpus size (outside every timed region)
    parity = None
    if args.parity_queries:
        parity = parity_block(world, rank, dev, nq, lexical, sparse, dense, tokens, q, last[0], last[1], last[2], pool,
                              dense_exact, args.parity_queries)

    # ---- end-to-end timing through the public API with host buffers
    q_host = q.pin()
    per = (nq + world - 1) // world
    out_ids = torch.empty((per, TOP_K), dtype=torch.int32).pin_memory()
    out_sc = torch.empty((per, TOP_K), dtype=torch.float64).pin_memory()
    searcher.search_host(q_host, out_ids, out_sc)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        searcher.search_host(q_host, out_ids, out_sc)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t) / args.steps
    h2d = searcher.last_h2d_bytes if hasattr(searcher, "last_h2d_bytes") else q_host.nbytes()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = ("measured (MEASURED_PEAKS.json: HBM copy GB/s for hbm-bound kernels, sustained bf16 TFLOP/s for tensor-bound ones)"
                if peaks else "fallback (B200_PROFILING.md)")

    # DRAM bytes of ONE captured launch (the largest round) from a committed `ncu --set full` pass, profiles/traffic.json: NOT
    # measured in this run - it belongs to the profiled configuration (1 GPU, full size) and is reported as captured
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh) if world == 1 and n_total == N_DOCS and nq == N_QUERIES else {}
    except Exception:
        pass

    def dram(name):
        return traffic.get(name, {}).get("dram_bytes")

    def with_traffic(out, name):
        out["traffic"] = dram(name)
        if dram(name):
            out["traffic_scope"] = f"largest launch of the step, from profiles/traffic.json ({traffic[name].get('round', 'r01')} ncu capture), not this run"
        return out

    def hbm(name, nbytes, union_bytes=None):
        if name in kern and kern[name]["ms"] > 0:
            ach = nbytes / (kern[name]["ms"] * 1e-3) / 1e9
            out = with_traffic({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                "ms": kern[name]["ms"], "launches": kern[name]["launches"], "algorithmic_bytes": nbytes}, name)
            if union_bytes:
                # posting lists shared by the queries of the batch counted once: what DRAM has to deliver at least
                out["batch_union_bytes"] = union_bytes
                out["frac_batch_union"] = union_bytes / (kern[name]["ms"] * 1e-3) / 1e9 / hbm_peak
            return out
        return None

    def tensor(name, flops):
        if name in kern and kern[name]["ms"] > 0:
            ach = flops / (kern[name]["ms"] * 1e-3) / 1e12
            return with_traffic({"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                                 "ms": kern[name]["ms"], "launches": kern[name]["launches"], "algorithmic_flops": flops}, name)
        return None

    rooflines = {}
    if "dpr" in systems:
        rooflines["dense_filter_gemm"] = tensor("dense_filter_gemm", algo["dpr_flops"])
    if "bm25" in systems:
        rooflines["sparse_tile_f64"] = hbm("sparse_tile_f64", algo["bm25_bytes"], algo.get("bm25_union_bytes"))
    if "splade" in systems:
        if "splade_head_gemm" in kern:
            # the SPLADE stage as a whole against SURVEY 8d's per-query posting bytes (what the round-1 kernel was measured on),
            # then its three kernels against what each of them actually moves / computes
            st_ms = sum(kern[n]["ms"] for n in kern if n.startswith("splade_"))
            ach = algo["splade_bytes"] / (st_ms * 1e-3) / 1e9
            rooflines["splade_stage"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                         "ms": st_ms, "algorithmic_bytes": algo["splade_bytes"], "traffic": None,
                                         "note": "SURVEY 8d bytes (every query re-reads its posting lists); the head terms are now a "
                                                 "tensor-core GEMM, so this is a work-equivalent rate, not a DRAM rate"}
            rooflines["splade_head_gemm"] = tensor("splade_head_gemm", algo["splade_head_flops"])
            rooflines["splade_tail_codes"] = hbm("splade_tail_codes", algo["splade_tail_bytes"])
        else:
            rooflines["sparse_tile_f32"] = hbm("sparse_tile_f32", algo["splade_bytes"], algo.get("splade_union_bytes"))
    if "colbert" in systems:
        rooflines["maxsim"] = hbm("maxsim", nq * TOP_K * algo["colbert_avg_tokens"] * 256.0 / world)   # this rank's share
    n_sys = len(systems)
    rooflines["fuse"] = hbm("fuse", 2 * (per * n_sys * TOP_K * 8.0 + per * TOP_K * 12.0))
    rooflines = {k: v for k, v in rooflines.items() if v}
    single = {k: v for k, v in rooflines.items() if k != "splade_stage"}
    dominant = max(single, key=lambda k: single[k]["ms"]) if single else None

    qps = nq / (ms_per_step * 1e-3)
    merged = {"bm25": ("bm25", "bm25_merge"), "dpr": ("dpr", "dpr_merge"), "splade": ("splade", "splade_merge"), "colbert": ("colbert",)}
    per_system = {name: nq / (sum(stage.get(x, 0.0) for x in parts) * 1e-3)
                  for name, parts in merged.items() if name in systems and sum(stage.get(x, 0.0) for x in parts) > 0}
    line = {
        "metric": "hybrid top-1000 queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None,
        "dtype": ("bf16 tensor-core filter + f32 exact rescoring (DPR, SPLADE head)" if dense_exact else "bf16 (DPR throughput mode)") +
                 ", bf16 (ColBERT), f64 (BM25), f32 (SPLADE, fusion)",
        "data": "synthetic",
        "config": {"workload": workload_name(systems, n_total, nq),
                   "docs": n_total, "queries": nq, "dense_mode": args.dense_mode,
                   "colbert_store_docs": tok_total if "colbert" in systems else 0,
                   "colbert_candidates": "global ids" if not pool else f"ids mapped into a pool of {pool} passages (id % pool)",
                   "corpus": "one seeded global corpus (generation chunks of 500k docs), sharded by contiguous doc range: the same corpus at every N",
                   "l2": "inputs larger than L2 (indexes are GBs, streamed every step)", "setup_s": setup_s,
                   "synth_generation_s": gen_s, "index_build_s": build_s,
                   "index_build_docs_per_s": {k: n_local / v for k, v in build_s.items() if v > 0},
                   "parallelism": f"corpus-sharded x{world}"},
        "e2e": {"value": nq / (e2e_ms_per_step * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": out_ids.numel() * 4 + out_sc.numel() * 8, "ms_per_step": e2e_ms_per_step},
        "gpu_launches": gpu_launches * args.steps,
        "clocks": clocks.summary(),
        "roofline": dict(rooflines[dominant], kernel=dominant, peak_source=peak_src) if dominant else None,
        "kernels": rooflines,
        "kernel_ms": kern,
        "stage_ms": stage,
        "per_system_qps": per_system,
        "parity": parity,
    }
    if not args.no_cpu_baseline:
        r = cpu_reference_sample(n_total, nq, systems)
        line["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "; ".join(f"{k}: {v}" for k, v in r["sample"].items()),
                                "sec_per_query": r["sec_per_query"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
