"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of ColBERT late interaction.

The arithmetic lives in colbert-ai (``colbert/modeling/colbert.py::colbert_score``, package
``colbert-ai @ git+https://github.com/stanford-futuredata/ColBERT.git@main``, unpinned in
``requirements.txt:15`` and absent from /root/reference): ``scores = D_padded @ Q^T``; padded
document tokens are set to -9999; max over document tokens, sum over query tokens.  Reference call
sites: ``src/retrievers/hybrid.py:120-137``, ``src/utils/colbert_ir.py:245-255``.  PARITY UNPINNED by
any reference-owned test (third-party, not installed).  Inputs are rounded to bf16 first (what the
B200 kernel stores) and the arithmetic is fp32.
"""
from __future__ import annotations

import torch


def maxsim_scores(q_tok: torch.Tensor, tok_ptr: torch.Tensor, tok_emb: torch.Tensor,
                  cand_ids: torch.Tensor) -> torch.Tensor:
    """q_tok [Q,Lq,D], tok_ptr [N+1], tok_emb [T,D], cand_ids [Q,C] (-1 = empty slot) -> scores [Q,C] fp32."""
    q = q_tok.to(torch.bfloat16).float()
    e = tok_emb.to(torch.bfloat16).float()
    Q, C = cand_ids.shape
    out = torch.full((Q, C), float("-inf"), dtype=torch.float32)
    for qi in range(Q):
        for ci in range(C):
            d = int(cand_ids[qi, ci])
            if d < 0:
                continue
            D = e[int(tok_ptr[d]):int(tok_ptr[d + 1])]            # [Ld, dim]
            if len(D) == 0:
                out[qi, ci] = -9999.0 * q.shape[1]
                continue
            s = D @ q[qi].t()                                     # [Ld, Lq]
            out[qi, ci] = s.max(dim=0).values.sum()
    return out
