"""The product package never imports or calls the oracle, and has no CPU fallback for scoring."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_does_not_import_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fusion_b200")):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "/root/reference" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch

    from fusion_b200 import ops
    from fusion_b200._lib import FusionB200Error
    with pytest.raises(FusionB200Error):
        ops.merge_topk(torch.zeros(1, 1, 4), torch.zeros(1, 1, 4, dtype=torch.int32), 2)
    with pytest.raises(FusionB200Error):
        ops.dense_scores(torch.zeros(2, 8), torch.zeros(3, 8))
