"""Golden fixture (TEST INFRASTRUCTURE): ``tests/golden/fusion_legacy.npz`` - nsf fusion under NumPy 1.x type promotion.

``Aggregator.weight_scores`` computes ``np.float32 score * python float weight`` (hybrid.py:291).  NumPy >= 2 (NEP 50, what
this container has) keeps float32; the NumPy 1.x the reference pins (torch 2.1.2 / pandas 2.1.4, requirements.txt) promotes
to float64 and ``aggregate_scores`` then sums doubles.  NumPy 1.x cannot be installed here, but its arithmetic is exactly
``float64(score) * float64(weight)``: running the VERBATIM reference with the weights given as ``np.float64`` scalars
produces it under any NumPy version (np.float32 * np.float64 -> float64 is version independent).
Inputs are those of fusion_small.npz.  Run:  python -m oracle.make_golden_promotion
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    mod = ref_loader.load_hybrid()
    g = np.load(os.path.join(ROOT, "tests", "golden", "fusion_small.npz"))
    systems = [str(s) for s in g["systems"]]
    lists = {s: [[{"corpus_id": int(i), "score": float(v)} for i, v in zip(ri, rs)]
                 for ri, rs in zip(g[f"in_ids_{s}"], g[f"in_scores_{s}"])] for s in systems}
    weights = {s: np.float64(w) for s, w in zip(systems, g["weights"])}
    distrs = {s: g[f"distr_{s}"] for s in systems}
    out = {}
    for norm in ("min-max", "z-score", "arctan", "percentile-rank"):
        res = mod.Aggregator.fuse(copy.deepcopy(lists), method="nsf", normalization=norm, linear_weights=weights,
                                  percentile_distributions=distrs)
        L = max(len(q) for q in res)
        ids = np.full((len(res), L), -1, dtype=np.int32)
        sc = np.full((len(res), L), np.nan, dtype=np.float64)
        for qi, q in enumerate(res):
            ids[qi, :len(q)] = [x["corpus_id"] for x in q]
            sc[qi, :len(q)] = [float(x["score"]) for x in q]
        assert type(res[0][0]["score"]).__name__ == "float64"
        out[f"out_ids_{norm}"], out[f"out_scores_{norm}"] = ids, sc
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_legacy.npz"), **out)
    print("wrote fusion_legacy.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
