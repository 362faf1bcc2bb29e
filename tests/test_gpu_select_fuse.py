"""GPU parity: K5 merge / rank rows and K4 fusion against the oracle and the reference goldens (through the C ABI)."""
import os

import numpy as np
import pytest
import torch

from oracle import fusion as ofusion

pytestmark = pytest.mark.gpu


def _ops():
    from fusion_b200 import ops
    return ops


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("g,q,k_in,k_out", [(2, 5, 50, 50), (8, 33, 1000, 1000), (4, 3, 3000, 100)])
def test_merge_topk(dtype, g, q, k_in, k_out):
    ops = _ops()
    gen = torch.Generator().manual_seed(g * 1000 + k_in)
    sc = torch.randn(g, q, k_in, generator=gen, dtype=torch.float64).to(dtype)
    sc = (sc * 8).round() / 8                      # many exact ties
    ids = torch.stack([torch.randperm(g * k_in * 2, generator=gen)[: g * k_in].reshape(g, k_in) for _ in range(q)], 1).to(torch.int32)
    ids[0, 0, :3] = -1                             # padding entries are ignored
    out_s, out_i = ops.merge_topk(sc.cuda(), ids.cuda(), k_out)
    for qi in range(q):
        s, i = sc[:, qi].reshape(-1), ids[:, qi].reshape(-1).long()
        keep = i >= 0
        s, i = s[keep], i[keep]
        order = np.lexsort((i.numpy(), -s.double().numpy()))[:k_out]
        assert out_i[qi].cpu().tolist() == i[order].tolist()
        assert torch.equal(out_s[qi].cpu(), s[order])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n,k", [(700, 700), (5000, 123), (20000, 20000), (40000, 1000)])
def test_rank_rows(dtype, n, k):
    ops = _ops()
    gen = torch.Generator().manual_seed(n)
    sc = ((torch.randn(3, n, generator=gen, dtype=torch.float64) * 16).round() / 16).to(dtype)
    sc[1, : n // 2] = 0.0
    out_s, out_i = ops.rank_rows(sc.cuda(), k, doc_base=7)
    for qi in range(3):
        order = np.argsort(-sc[qi].double().numpy(), kind="stable")[:k]
        assert np.array_equal(out_i[qi].cpu().numpy(), order + 7)
        assert torch.equal(out_s[qi].cpu(), sc[qi][order])


FUSION_CASES = [("bcf", None), ("rrf", None)] + [("nsf", n) for n in ofusion.NORMALIZATIONS]


@pytest.mark.parametrize("method,norm", FUSION_CASES)
def test_fuse_golden(golden_dir, method, norm):
    """Fixtures produced by the verbatim reference Aggregator.fuse (oracle/make_golden.py)."""
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "fusion_small.npz"))
    systems = [str(s) for s in g["systems"]]
    tag = method if norm is None else f"{method}_{norm}"
    exp_ids, exp_sc = g[f"out_ids_{tag}"], g[f"out_scores_{tag}"]
    lists = [(torch.from_numpy(g[f"in_ids_{s}"]).cuda(), torch.from_numpy(g[f"in_scores_{s}"]).cuda(), None) for s in systems]
    ids, sc, lens = ops.fuse(lists, method, norm, list(g["weights"]), [g[f"distr_{s}"] for s in systems])
    ids, sc, lens = ids.cpu().numpy(), sc.cpu().numpy(), lens.cpu().numpy()
    for qi in range(exp_ids.shape[0]):
        n = int((exp_ids[qi] >= 0).sum())
        assert lens[qi] == n
        tol = 1e-12 if (method != "nsf" or norm == "none") else 1e-5
        np.testing.assert_allclose(sc[qi, :n], exp_sc[qi, :n], rtol=tol, atol=tol)
        # fusion must reproduce the EXACT id sequence, insertion-order tie-break included
        assert ids[qi, :n].tolist() == exp_ids[qi, :n].tolist(), (tag, qi)


@pytest.mark.parametrize("norm", ["min-max", "z-score", "arctan", "percentile-rank"])
def test_fuse_numpy1_promotion_golden(golden_dir, norm):
    """FZ_FUSE_PROMOTE_F64 / Aggregator.fuse(numpy_promotion='legacy'): the reference's arithmetic under the NumPy 1.x it
    pins (fp32 normalised scores weighted and summed in float64) - exact id sequence, scores to 1e-5 (fp32 normalisation)."""
    from fusion_b200.retrievers.hybrid import Aggregator
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "fusion_small.npz"))
    leg = np.load(os.path.join(golden_dir, "fusion_legacy.npz"))
    systems = [str(s) for s in g["systems"]]
    lists = [(torch.from_numpy(g[f"in_ids_{s}"]).cuda(), torch.from_numpy(g[f"in_scores_{s}"]).cuda(), None) for s in systems]
    ids, sc, lens = ops.fuse(lists, "nsf", norm, list(g["weights"]), [g[f"distr_{s}"] for s in systems], promote_f64=True)
    exp_ids, exp_sc = leg[f"out_ids_{norm}"], leg[f"out_scores_{norm}"]
    for qi in range(exp_ids.shape[0]):
        n = int((exp_ids[qi] >= 0).sum())
        assert int(lens[qi]) == n
        assert ids[qi, :n].cpu().tolist() == exp_ids[qi, :n].tolist(), (norm, qi)
        np.testing.assert_allclose(sc[qi, :n].cpu().numpy(), exp_sc[qi, :n], rtol=1e-5, atol=1e-5)
    # the list-of-dict adapter: 'legacy' returns Python floats of the float64 sums, 'nep50' np.float32 like NumPy >= 2
    ranked = {s: [[{"corpus_id": int(i), "score": float(v)} for i, v in zip(ri, rs)]
                  for ri, rs in zip(g[f"in_ids_{s}"], g[f"in_scores_{s}"])] for s in systems}
    w = dict(zip(systems, g["weights"].tolist()))
    d = {s: g[f"distr_{s}"] for s in systems}
    res = Aggregator.fuse(ranked, "nsf", norm, w, d, numpy_promotion="legacy")
    assert [x["corpus_id"] for x in res[0]] == exp_ids[0, :len(res[0])].tolist() and isinstance(res[0][0]["score"], float)
    res = Aggregator.fuse(ranked, "nsf", norm, w, d, numpy_promotion="nep50")
    assert isinstance(res[0][0]["score"], np.float32)
    assert Aggregator.fuse({s: [] for s in systems}, "nsf", norm, w, d) == []           # zero queries (hybrid.py returns [])


def test_metrics_reject_more_gold_ids_than_the_kernel_holds():
    """More than 256 distinct relevant docs for a query: an error, never a silently truncated recall / nDCG."""
    from fusion_b200._lib import FusionB200Error
    ops = _ops()
    ids = torch.arange(1000, dtype=torch.int32).repeat(2, 1).cuda()
    ok = ops.rank_metrics(ids, None, torch.tensor([0, 256, 300], dtype=torch.int32).cuda(),
                          torch.arange(300, dtype=torch.int32).cuda())
    assert float(ok[0]) > 0
    with pytest.raises(FusionB200Error, match="256"):
        ops.rank_metrics(ids, None, torch.tensor([0, 257, 300], dtype=torch.int32).cuda(), torch.arange(300, dtype=torch.int32).cuda())


def _random_lists(rng, q, n_list, pool):
    lists = []
    for n in n_list:
        ids = np.stack([rng.choice(pool, n, replace=False) for _ in range(q)]).astype(np.int32)
        sc = -np.sort(-rng.normal(0, 2, (q, n)), axis=1)
        lists.append((ids, sc))
    return lists


@pytest.mark.parametrize("n_list,pool", [([1000, 1000, 1000, 1000], 3000), ([3000, 2500], 4000), ([28000, 28000], 28000)])
@pytest.mark.parametrize("method,norm", [("rrf", None), ("bcf", None), ("nsf", "z-score"), ("nsf", "min-max"), ("nsf", "none")])
def test_fuse_vs_oracle(method, norm, n_list, pool):
    """Shared-memory path (4 x 1000), global-workspace path and the reference's full-length lists (n = N = 28k)."""
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(len(n_list) * 7 + pool))
    q = 3
    host = _random_lists(rng, q, n_list, pool)
    w = [1.0 / len(n_list)] * len(n_list)
    lists = [(torch.from_numpy(i).cuda(), torch.from_numpy(s).cuda(), None) for i, s in host]
    ids, sc, lens = ops.fuse(lists, method, norm, w)
    ids, sc, lens = ids.cpu().numpy(), sc.cpu().numpy(), lens.cpu().numpy()
    for qi in range(q):
        eids, esc = ofusion.fuse_query([h[0][qi] for h in host], [h[1][qi] for h in host], method, norm, w)
        esc = np.asarray(esc, dtype=np.float64)
        n = len(eids)
        assert lens[qi] == n
        tol = 1e-12 if (method != "nsf" or norm == "none") else 1e-5
        np.testing.assert_allclose(sc[qi, :n], esc, rtol=tol, atol=tol)
        if method != "nsf" or norm == "none":
            assert ids[qi, :n].tolist() == eids                     # exact sequence incl. insertion-order ties
        else:
            assert sorted(ids[qi, :n].tolist()) == sorted(eids)
            mism = np.flatnonzero(ids[qi, :n] != np.asarray(eids))
            for p in mism:                                          # swaps only between near-equal fused scores
                assert abs(esc[p] - esc[eids.index(int(ids[qi, p]))]) <= 2e-5


def test_fuse_ragged_and_float32_inputs():
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(5))
    q = 4
    host = _random_lists(rng, q, [40, 30], 60)
    lens0 = np.array([40, 0, 17, 2], dtype=np.int32)   # (a 1-element list gives z-score NaN in the reference)
    lists = [(torch.from_numpy(host[0][0]).cuda(), torch.from_numpy(host[0][1]).cuda(), torch.from_numpy(lens0).cuda()),
             (torch.from_numpy(host[1][0]).cuda(), torch.from_numpy(host[1][1].astype(np.float32)).cuda(), None)]
    ids, sc, lens = ops.fuse(lists, "nsf", "z-score", [0.3, 0.7])
    for qi in range(q):
        eids, esc = ofusion.fuse_query([host[0][0][qi, :lens0[qi]], host[1][0][qi]],
                                       [host[0][1][qi, :lens0[qi]], host[1][1][qi].astype(np.float32).astype(np.float64)],
                                       "nsf", "z-score", [0.3, 0.7])
        n = len(eids)
        assert int(lens[qi]) == n
        got = np.asarray(sc[qi, :n].cpu())
        exp = np.asarray(esc, dtype=np.float64)
        np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-5, equal_nan=True)


# ------------------------------------------------------------------------------------------------ metrics + weight sweep
def _sweep_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "sweep_small.npz"))
    systems = [str(s) for s in g["systems"]]
    nq = g[f"in_ids_{systems[0]}"].shape[0]
    results = {k: [[{"corpus_id": int(i), "score": float(s)} for i, s in zip(g[f"in_ids_{k}"][q], g[f"in_scores_{k}"][q])]
                   for q in range(nq)] for k in systems}
    gp, gi = g["gold_ptr"], g["gold_ids"]
    golds = [gi[gp[i]:gp[i + 1]].tolist() for i in range(nq)]
    return g, systems, results, golds


@pytest.mark.parametrize("norm", ["min-max", "z-score", "none"])
def test_weight_sweep_matches_reference(golden_dir, norm):
    """hybrid.py:404-426 executed verbatim (Aggregator.fuse + Metrics per weight vector) vs ONE fz_fuse_sweep launch."""
    from fusion_b200.retrievers.hybrid import tune_linear_fusion_weights, weight_grid
    g, systems, results, golds = _sweep_fixture(golden_dir)
    rows = tune_linear_fusion_weights(results, golds, norm, step=0.25)
    assert len(rows) == g["weights"].shape[0]
    names = [str(x) for x in g["metric_names"]]
    for r, w, exp in zip(rows, g["weights"], g[f"metrics_{norm}"]):
        assert [r[f"weight_{k}"] for k in systems] == w.tolist()
        np.testing.assert_allclose([r[n] for n in names], exp, rtol=1e-12, atol=1e-12)
    assert len(weight_grid(["a", "b", "c", "d"], 0.05)) == 1771          # the reference's four-system grid


def test_metrics_class_matches_reference(golden_dir):
    """Metrics.compute_all_metrics on fused rankings == the reference's values for the same rankings."""
    from fusion_b200.retrievers.hybrid import Aggregator
    from fusion_b200.utils.metrics import Metrics
    g, systems, results, golds = _sweep_fixture(golden_dir)
    w = g["weights"][5]
    ranked = Aggregator.fuse(results, method="nsf", normalization="z-score", linear_weights=dict(zip(systems, w.tolist())))
    ev = Metrics(recall_at_k=[5, 10, 20, 50, 100, 200, 500, 1000], map_at_k=[10, 100], mrr_at_k=[10, 100], ndcg_at_k=[10, 100])
    sc = ev.compute_all_metrics(golds, [[x["corpus_id"] for x in r] for r in ranked])
    names = [str(x) for x in g["metric_names"]]
    np.testing.assert_allclose([sc[n] for n in names], g["metrics_z-score"][5], rtol=1e-12, atol=1e-12)
    # per-query helpers against the oracle restatement
    from oracle import metrics as om
    res0 = [x["corpus_id"] for x in ranked[0]]
    exp = om.query_metrics(golds[0], res0)
    assert abs(ev.recall(golds[0], res0, 10) - exp[1]) < 1e-12
    assert abs(ev.average_precision(golds[0], res0, 100) - exp[9]) < 1e-12
    assert abs(ev.ndcg(golds[0], res0, 10) - exp[12]) < 1e-12
    assert abs(ev.r_precision(golds[0], res0) - exp[14]) < 1e-12
    with pytest.raises(ZeroDivisionError):
        ev.compute_all_metrics([[]], [[1, 2]])


def test_metrics_oracle_random_vs_kernel():
    from fusion_b200 import ops
    from oracle import metrics as om
    rng = np.random.Generator(np.random.PCG64(5))
    nq, n = 40, 300
    ids = np.stack([rng.permutation(1000)[:n] for _ in range(nq)]).astype(np.int32)
    golds = [rng.choice(1000, size=int(rng.integers(1, 9)), replace=False).tolist() for _ in range(nq)]
    golds[3] = golds[3] + golds[3][:1]                 # a duplicated gold id counts twice in len(gold) only
    gp = np.cumsum([0] + [len(g) for g in golds]).astype(np.int32)
    gi = np.concatenate(golds).astype(np.int32)
    got = ops.rank_metrics(torch.from_numpy(ids).cuda(), None, torch.from_numpy(gp).cuda(), torch.from_numpy(gi).cuda())
    exp = om.mean_metrics(golds, [r.tolist() for r in ids])
    np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("method,norm", [("nsf", "z-score"), ("rrf", None), ("nsf", "none"), ("bcf", None)])
def test_fuse_truncated_output_equals_head_of_full_fusion(method, norm):
    """out_stride < |union| takes the radix-select + short-sort path; it must equal the head of the full sort, ties
    (constant lists, rrf ranks shared between systems) included."""
    from fusion_b200 import ops
    rng = np.random.Generator(np.random.PCG64(77))
    nq, n, pool = 9, 700, 1500
    lists = []
    for s in range(4):
        ids = np.stack([rng.permutation(pool)[:n] for _ in range(nq)]).astype(np.int32)
        sc = -np.sort(-rng.normal(0, 2, (nq, n)), axis=1)
        if s == 2:
            sc[:] = 0.75                                  # a constant list: every entry ties
        lists.append((torch.from_numpy(ids).cuda(), torch.from_numpy(sc).cuda(), None))
    w = [0.4, 0.3, 0.2, 0.1] if method == "nsf" else None
    full_i, full_s, full_n = ops.fuse(lists, method, norm, w)
    for stride in (1, 37, 1000):
        ti, ts, tn = ops.fuse(lists, method, norm, w, out_stride=stride)
        assert torch.equal(ti, full_i[:, :stride]) and torch.equal(ts, full_s[:, :stride])
        assert torch.equal(tn, torch.clamp(full_n, max=stride))


def test_ranking_tsv_round_trip_and_msmarco_evaluation(tmp_path):
    """colbert_ir.py:261-345: ranking file format + recall@depth / MRR@10 / R-precision, against a plain-Python restatement
    of that loop."""
    from fusion_b200.utils import colbert_ir as ci
    rng = np.random.Generator(np.random.PCG64(3))
    nq, k, pool = 37, 60, 400
    ids = np.stack([rng.permutation(pool)[:k] for _ in range(nq)]).astype(np.int32)
    ids[5, 40:] = -1
    scores = -np.sort(-rng.random((nq, k)), axis=1)
    golds = [rng.choice(pool, size=int(rng.integers(1, 6)), replace=False).tolist() for _ in range(nq)]
    qids = list(range(100, 100 + nq))
    path = str(tmp_path / "ranking.tsv")
    n = ci.save_ranking_tsv(path, qids, torch.from_numpy(ids), torch.from_numpy(scores))
    assert n == nq * k - 20
    q2, i2, s2 = ci.load_ranking_tsv(path)
    assert q2 == qids and torch.equal(i2, torch.from_numpy(ids))
    assert torch.equal(s2[ids >= 0], torch.from_numpy(scores)[ids >= 0])
    # restatement of the reference loop
    depths = (5, 10, 50)
    mrr, rp, rec = 0.0, 0.0, {d: 0.0 for d in depths}
    for qi in range(nq):
        ranking = [p for p in ids[qi].tolist() if p >= 0]
        pos = golds[qi]
        for rank, pid in enumerate(ranking, start=1):
            if rank <= 10 and pid in pos:
                mrr += 1.0 / rank
                break
        for rank, pid in enumerate(ranking, start=1):
            if rank <= len(pos) and pid in pos:
                rp += 1.0 / len(pos)
            if pid in pos:
                for d in depths:
                    if rank <= d:
                        rec[d] += 1.0 / len(pos)
    gp = torch.tensor(np.concatenate([[0], np.cumsum([len(g) for g in golds])]), dtype=torch.int32).cuda()
    gi = torch.tensor([x for g in golds for x in g], dtype=torch.int32).cuda()
    got = ci.evaluate_ranking_tensors(i2.cuda(), gp, gi, depths=depths)
    assert got["mrr@10"] == pytest.approx(mrr / nq, abs=1e-12)
    assert got["rp"] == pytest.approx(rp / nq, abs=1e-12)
    for d in depths:
        assert got[f"recall@{d}"] == pytest.approx(rec[d] / nq, abs=1e-12)
