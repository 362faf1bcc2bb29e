"""Golden fixture for ColBERT MaxSim (TEST INFRASTRUCTURE): ``tests/golden/maxsim_small.npz``.

colbert-ai (``requirements.txt:15``, unpinned ``@main``, not installed, absent from /root/reference) holds the
arithmetic, so no reference-owned vector exists.  This script pins ``oracle/maxsim.py`` (a per-pair loop over ragged
passages) against an INDEPENDENT formulation of the published algorithm - colbert-ai's ``colbert_score``: pad the
candidates' token matrices to one [C, Ld_max, dim] batch, ``D_padded @ Q^T``, set padded rows to -9999, max over document
tokens, sum over query tokens - evaluated in numpy float64 on bf16-rounded inputs.  Run:  python -m oracle.make_golden_maxsim
"""
from __future__ import annotations

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bf16(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(x).to(torch.bfloat16).float().numpy()


def colbert_score_padded(q: np.ndarray, tok_ptr: np.ndarray, emb: np.ndarray, cand: np.ndarray) -> np.ndarray:
    """q [Lq, dim], candidates of ONE query -> scores [C] (float64), padded-batch formulation."""
    lens = np.array([tok_ptr[c + 1] - tok_ptr[c] for c in cand])
    ld = max(int(lens.max()), 1)
    d_padded = np.zeros((len(cand), ld, emb.shape[1]), dtype=np.float64)
    mask = np.zeros((len(cand), ld), dtype=bool)
    for i, c in enumerate(cand):
        d_padded[i, :lens[i]] = emb[tok_ptr[c]:tok_ptr[c + 1]]
        mask[i, :lens[i]] = True
    scores = d_padded @ q.astype(np.float64).T                   # [C, Ld, Lq]
    scores[~mask] = -9999.0                                      # D_padding (colbert_score)
    return scores.max(axis=1).sum(axis=1)


def main():
    rng = np.random.Generator(np.random.PCG64(4242))
    n_docs, dim, nq, lq, n_cand = 120, 128, 5, 32, 40
    lens = rng.integers(1, 180, n_docs)
    lens[7] = 0                                                  # an empty passage: every row is padding
    lens[11] = 1
    lens[13] = 180
    tok_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=tok_ptr[1:])
    emb = rng.standard_normal((int(tok_ptr[-1]), dim)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    q = rng.standard_normal((nq, lq, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    emb, q = bf16(emb), bf16(q)
    cand = np.stack([rng.choice(n_docs, n_cand, replace=False) for _ in range(nq)]).astype(np.int32)
    cand[0, :3] = [7, 11, 13]
    cand[1, 5] = cand[1, 6]                                      # a repeated candidate
    scores = np.stack([colbert_score_padded(q[i], tok_ptr, emb, cand[i]) for i in range(nq)])
    out = os.path.join(ROOT, "tests", "golden", "maxsim_small.npz")
    np.savez_compressed(out, q=q, tok_ptr=tok_ptr, tok_emb=emb, cand=cand, scores=scores)
    print("wrote", out, scores.shape, float(scores.min()), float(scores.max()))


if __name__ == "__main__":
    main()
