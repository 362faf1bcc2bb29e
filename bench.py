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


def make_lexical(n_docs, seed, device):
    """doc lengths clip(lognormal(3.3, 0.5), 3, 256), Zipf(1.07) over 500k terms."""
    g = _gen(device, seed)
    lens = torch.exp(torch.randn(n_docs, device=device, generator=g) * 0.5 + 3.3).clamp_(3, 256).long()
    ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    toks = zipf_draw(zipf_cdf(BM25_VOCAB, 1.07, device), int(ptr[-1]), g).to(torch.int32)
    return ptr, toks


def make_lexical_queries(nq, seed, device):
    g = _gen(device, seed)
    lens = 1 + torch.poisson(torch.full((nq,), 4.0, device=device), generator=g).long()
    ptr = torch.zeros(nq + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    toks = zipf_draw(zipf_cdf(BM25_VOCAB, 1.07, device), int(ptr[-1]), g).to(torch.int32)
    return ptr.to(torch.int32), toks


def make_splade(n, mean_nnz, lo, hi, seed, device, chunk=500_000):
    """CSR sparse vectors: ~Poisson(mean_nnz) distinct Zipf(1.05) terms per row, weights log1p(relu(N(0.5, 0.7))) > 0."""
    cdf = zipf_cdf(SPLADE_VOCAB, 1.05, device)
    ptrs, terms, ws = [], [], []
    base = 0
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        g = _gen(device, seed * 1000 + c0 // chunk)
        want = torch.poisson(torch.full((m,), float(mean_nnz), device=device), generator=g).clamp_(lo, hi).long()
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
        ptrs.append(p[1:] + base)
        base += int(p[-1])
        terms.append(uterm.to(torch.int32))
        ws.append(w.float())
        del key, urow, uterm, prio, order, keep
    ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=device)] + ptrs)
    return ptr, torch.cat(terms), torch.cat(ws)


def make_dense_index(n, dim, seed, device, doc_base, chunk=1_000_000):
    from fusion_b200 import ops
    from fusion_b200.index import DenseIndex
    d32 = torch.empty((n, dim), dtype=torch.float32, device=device)
    d16 = torch.empty((n, dim), dtype=torch.bfloat16, device=device)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        x = torch.randn((m, dim), device=device, generator=_gen(device, seed * 1000 + (doc_base + c0) // chunk))
        a, b = ops.normalize_rows(x)
        d32[c0:c0 + m], d16[c0:c0 + m] = a, b
    return DenseIndex(d32, d16, "cos_sim", doc_base)


def make_token_store(n_docs, seed, device, doc_base, chunk=100_000):
    from fusion_b200.index import TokenStore
    g = _gen(device, seed)
    lens = torch.poisson(torch.full((n_docs,), 70.0, device=device), generator=g).clamp_(8, 180).long()
    ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(lens, 0)
    total = int(ptr[-1])
    emb = torch.empty((total, 128), dtype=torch.bfloat16, device=device)
    step = chunk * 70
    for t0 in range(0, total, step):
        m = min(step, total - t0)
        x = torch.randn((m, 128), device=device, generator=g)
        emb[t0:t0 + m] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    return TokenStore(ptr, emb, doc_base)


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


def run_reference(args):
    systems = args.systems.split(",")
    t0 = time.perf_counter()
    r = cpu_reference_sample(args.docs, args.queries, systems)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "hybrid top-1000 queries/sec", "value": r["qps"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.queries / r["qps"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
        "config": {"workload": f"C5 hybrid {'+'.join(systems)} top-{TOP_K}, {args.docs} docs, {args.queries} queries (CPU sample extrapolated)"},
        "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "; ".join(f"{k}: {v}" for k, v in r["sample"].items()),
                         "sec_per_query": r["sec_per_query"], "sample_wall_s": wall},
        "e2e": {"value": r["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    t_setup = time.perf_counter()

    lexical = sparse = dense = tokens = None
    q = HybridQueries()
    algo = {}
    if "bm25" in systems:
        dptr, dtok = make_lexical(n_local, 301 * 10 + rank, dev)
        lexical = LexicalIndex(dptr, dtok, BM25_VOCAB, "bm25", 0.9, 0.4, device=dev, doc_base=lo,
                               stats_reduce=lambda n, df, sdl: sharding.allreduce_lexical_stats(n, df, sdl, dev))
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
        dp, dt, dw = make_splade(n_local, 120, 8, 512, 311 * 10 + rank, dev)
        sparse = SparseIndex(dp, dt, dw, SPLADE_VOCAB, "cos_sim", device=dev, doc_base=lo)
        del dp, dt, dw
        qp, qt, qw = make_splade(nq, 24, 2, 64, 312, dev)
        q.sp_ptr, q.sp_term, q.sp_weight = sparse_queries(qp, qt, qw, "cos_sim", dev)
        ptr_h, term_h = q.sp_ptr.cpu().numpy(), q.sp_term.cpu().numpy()
        df_loc = torch.bincount(sparse.doc_post[:, 0].long(), minlength=SPLADE_VOCAB).cpu().numpy()
        algo["splade_bytes"] = float(sum(df_loc[term_h[ptr_h[i]:ptr_h[i + 1]]].sum() for i in range(nq)) * 8 + nq * TOP_K * 8)
        algo["splade_union_bytes"] = float(df_loc[np.unique(term_h[term_h >= 0])].sum()) * 8 + nq * TOP_K * 8
        algo["splade_index_bytes"] = sparse.nbytes()
    if "dpr" in systems:
        dense = make_dense_index(n_local, DIM, 201, dev, lo)
        q.dense = torch.randn((nq, DIM), device=dev, generator=_gen(dev, 202))
        algo["dpr_flops"] = 2.0 * nq * n_local * DIM
    if "colbert" in systems:
        plo, phi = sharding.shard_bounds(args.pool, world, rank)
        tokens = make_token_store(phi - plo, 401 * 10 + rank, dev, plo)
        x = torch.randn((nq, COLBERT_LQ, 128), device=dev, generator=_gen(dev, 402))
        q.colbert = (x / x.norm(dim=2, keepdim=True)).to(torch.bfloat16)
        algo["colbert_avg_tokens"] = float(tokens.n_tokens) / max(1, tokens.n_docs)
        tokens.packed(drop_plain=True)          # the kernel streams the packed image; the plain matrix is not needed again
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    searcher = HybridSearcher(lexical, sparse, dense, tokens, k=TOP_K, fusion="nsf", normalization="z-score",
                              colbert_pool=args.pool if tokens is not None else None)
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
    prof_steps = 1
    for _ in range(prof_steps):
        step(q)
        searcher.collect_stage_ms()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.fz_profile_summary(buf, len(buf))
    lib.fz_profile_enable(0)
    searcher.timing = False
    kern = {}
    for line in buf.value.decode().splitlines():
        name, cnt, tot = line.split()
        kern[name] = {"launches": int(cnt) // prof_steps, "ms": float(tot) / prof_steps}
    gpu_launches = sum(v["launches"] for v in kern.values())

    # ---- end-to-end timing through the public API with host buffers
    q_host = q.pin()
    qs_lo, qs_hi = sharding.query_slice(nq, world, rank)
    per = (nq + world - 1) // world
    out_ids = torch.empty((per, TOP_K), dtype=torch.int32).pin_memory()
    out_sc = torch.empty((per, TOP_K), dtype=torch.float64).pin_memory()
    searcher.search_host(q_host, out_ids, out_sc)
    barrier()
    t0 = time.perf_counter()
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

    # DRAM bytes of ONE captured launch (the largest round) from the committed `ncu --set full` pass, profiles/traffic.json;
    # it belongs to the profiled configuration (1 GPU, full size) and is reported as captured, not rescaled.
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh) if world == 1 and n_total == N_DOCS and nq == N_QUERIES else {}
    except Exception:
        pass

    def dram(name):
        return traffic.get(name, {}).get("dram_bytes")

    def hbm(name, nbytes, union_bytes=None):
        if name in kern and kern[name]["ms"] > 0:
            ach = nbytes / (kern[name]["ms"] * 1e-3) / 1e9
            out = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                   "traffic": dram(name), "traffic_scope": "largest launch of the step (ncu)" if dram(name) else None,
                   "ms": kern[name]["ms"], "launches": kern[name]["launches"], "algorithmic_bytes": nbytes}
            if union_bytes:
                # posting lists shared by the queries of the batch counted once: what DRAM has to deliver at least
                out["batch_union_bytes"] = union_bytes
                out["frac_batch_union"] = union_bytes / (kern[name]["ms"] * 1e-3) / 1e9 / hbm_peak
            return out
        return None

    rooflines = {}
    if "dense_filter_gemm" in kern:
        ach = algo["dpr_flops"] / (kern["dense_filter_gemm"]["ms"] * 1e-3) / 1e12
        rooflines["dense_filter_gemm"] = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                                          "frac": ach / tf_peak, "traffic": dram("dense_filter_gemm"), "ms": kern["dense_filter_gemm"]["ms"],
                                          "launches": kern["dense_filter_gemm"]["launches"], "algorithmic_flops": algo["dpr_flops"]}
    if "bm25" in systems:
        rooflines["sparse_tile_f64"] = hbm("sparse_tile_f64", algo["bm25_bytes"], algo.get("bm25_union_bytes"))
    if "splade" in systems:
        rooflines["sparse_tile_f32"] = hbm("sparse_tile_f32", algo["splade_bytes"], algo.get("splade_union_bytes"))
    if "colbert" in systems:
        rooflines["maxsim"] = hbm("maxsim", nq * TOP_K * algo["colbert_avg_tokens"] * 256.0 / world)   # this rank's share
    n_sys = len(systems)
    rooflines["fuse"] = hbm("fuse", 2 * (per * n_sys * TOP_K * 8.0 + per * TOP_K * 12.0))
    rooflines = {k: v for k, v in rooflines.items() if v}
    dominant = max(rooflines, key=lambda k: rooflines[k]["ms"]) if rooflines else None

    qps = nq / (ms_per_step * 1e-3)
    line = {
        "metric": "hybrid top-1000 queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 tensor-core filter + f32 rescoring (DPR, ColBERT bf16), f64 (BM25), f32 (SPLADE, fusion)",
        "data": "synthetic",
        "config": {"workload": f"C5 hybrid {'+'.join(systems)} top-{TOP_K} + nsf z-score and rrf fusion, {n_total} docs, {nq} queries, d={DIM}",
                   "docs": n_total, "queries": nq, "colbert_pool_docs": args.pool if "colbert" in systems else 0,
                   "l2": "inputs larger than L2 (indexes are GBs, streamed every step)", "setup_s": setup_s,
                   "parallelism": f"corpus-sharded x{world}"},
        "e2e": {"value": nq / (e2e_ms_per_step * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q_host.nbytes(),
                "d2h_bytes_per_step": out_ids.numel() * 4 + out_sc.numel() * 8, "ms_per_step": e2e_ms_per_step},
        "gpu_launches": gpu_launches * args.steps,
        "clocks": clocks.summary(),
        "roofline": dict(rooflines[dominant], kernel=dominant, peak_source=peak_src) if dominant else None,
        "kernels": rooflines,
        "kernel_ms": kern,
        "stage_ms": {k: v / prof_steps for k, v in searcher.stage_ms.items()},
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
