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


def test_activation_head_refuses_cpu_tensors_and_bad_pooling():
    import pytest
    import torch

    from fusion_b200 import activations as A
    from fusion_b200._lib import FusionB200Error
    from fusion_b200.retrievers.splade import SPLADE
    with pytest.raises(FusionB200Error):
        A.splade_pool(torch.zeros(2, 3, 5), torch.ones(2, 3, dtype=torch.long), "max")
    with pytest.raises(FusionB200Error):
        A.activations_to_csr(torch.zeros(2, 5))
    with pytest.raises(FusionB200Error):
        A.prune_activations(torch.zeros(2, 5), 2)
    with pytest.raises(AssertionError):          # the reference's assertion (splade.py:74)
        A.splade_pool(torch.zeros(2, 3, 5), torch.ones(2, 3, dtype=torch.long), "mean")
    with pytest.raises(AssertionError):
        SPLADE(torch.nn.Identity(), pooling="avg")


def test_sparse_topk_shard_sync_is_inert_for_one_shard():
    """ops._SyncCall builds no C struct when there is nothing to exchange (one shard / no sync): the library then gets
    NULL and runs the plain single-GPU schedule."""
    import torch

    from fusion_b200 import ops
    assert ops._SyncCall(None, 4, torch.float32, "cpu").ref() is None
    assert ops._SyncCall(ops.ShardSync(lambda t: t, 1, 100), 4, torch.float32, "cpu").ref() is None
    sc = ops._SyncCall(ops.ShardSync(lambda t: t.fill_(1.0), 2, 100), 4, torch.float32, "cpu")
    assert sc.ref() is not None and sc.struct.n_shards == 2 and sc.struct.sched_docs == 100 and sc.struct.floor_rank == 0
    assert ops._SyncCall(ops.ShardSync(lambda t: t, 8, 100), 4, torch.float32, "cpu", k_global=1000).struct.floor_rank == 125
    assert sc.struct.hook(None) == 0 and sc.exchange.tolist() == [1.0] * 4          # the callback runs the reduction

    def boom(t):
        raise RuntimeError("collective failed")
    sc = ops._SyncCall(ops.ShardSync(boom, 2, 100), 4, torch.float32, "cpu")
    assert sc.struct.hook(None) == 1
    try:
        sc.reraise()
        raise AssertionError("expected the stored exception")
    except RuntimeError as e:
        assert "collective failed" in str(e)


def test_package_never_imports_the_reference_tree():
    """A drop-in replacement must not need /root/reference on sys.path: no `from src...` / `import src...` anywhere."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fusion_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+src(\.|\s)", src, flags=re.M), os.path.join(dirpath, f)
