"""GPU parity: SPLADE activation head - pooling, pruning, dense -> CSR - against the fixture made by
the verbatim reference (tests/golden/splade_head_small.npz) and against the torch-CPU oracle on seeded inputs."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

POOL_RTOL = 2e-6        # device log1pf vs torch CPU log1p: a couple of ulps; fp32 sum order differs for 'sum'


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "splade_head_small.npz"))


def _check_pruned(got: np.ndarray, act: np.ndarray, k: int):
    """Kept values equal the k largest of the row; everything else is zero.  Which member of a tie group at the cutoff
    survives is unspecified in torch.topk, so values are compared, not positions inside the tie."""
    for r in range(act.shape[0]):
        kept = got[r] != 0
        want = np.sort(act[r])[::-1][:k]
        assert np.array_equal(np.sort(got[r][kept])[::-1], want[want != 0])
        assert np.array_equal(got[r][kept], act[r][kept])
        cutoff = want[-1]
        assert kept[act[r] > cutoff].all() and not kept[act[r] < cutoff].any()


@pytest.mark.parametrize("pooling", ["max", "sum"])
def test_pool_prune_match_reference_fixture(golden_dir, pooling):
    from fusion_b200 import activations as A
    g = _golden(golden_dir)
    logits, mask = torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    act = A.splade_pool(logits, mask, pooling)
    want = g[f"act_{pooling}"]
    np.testing.assert_allclose(act.cpu().numpy(), want, rtol=POOL_RTOL, atol=1e-7)
    assert np.array_equal(act.cpu().numpy() == 0, want == 0)            # the sparsity pattern is exact
    ref_act = torch.from_numpy(want).cuda()
    for k in (1, 32, 517):
        pruned, idx = A.prune_activations(ref_act, k)
        _check_pruned(pruned.cpu().numpy(), want, k)
        assert idx.shape == (want.shape[0], k)
        vals = np.take_along_axis(want, idx.cpu().numpy(), axis=1)
        assert np.array_equal(vals, np.take_along_axis(want, g[f"topk_{pooling}_{k}"], axis=1))   # same values, best first


def test_csr_matches_oracle_and_feeds_the_index(golden_dir):
    from fusion_b200 import activations as A
    from oracle import splade_head as oh
    g = _golden(golden_dir)
    act = torch.from_numpy(g["act_max"])
    ptr, term, w = A.activations_to_csr(act.cuda())
    optr, oterm, ow = oh.to_csr(act)
    assert np.array_equal(ptr.cpu().numpy(), optr)
    assert np.array_equal(term.cpu().numpy(), oterm)
    assert np.array_equal(w.cpu().numpy(), ow)
    empty = torch.zeros((3, 40), device="cuda")
    p0, t0, w0 = A.activations_to_csr(empty)
    assert p0.tolist() == [0, 0, 0, 0] and t0.numel() == 0 and w0.numel() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pooling", ["max", "sum"])
def test_pool_random_shapes_vs_oracle(dtype, pooling):
    from fusion_b200 import activations as A
    from oracle import splade_head as oh
    g = torch.Generator().manual_seed(5)
    for (b, l, v) in [(1, 1, 1), (3, 7, 33), (5, 64, 1000), (2, 19, 32005)]:
        logits = (torch.randn((b, l, v), generator=g) * 3).to(dtype)
        lens = torch.randint(0, l + 1, (b,), generator=g)
        mask = (torch.arange(l)[None, :] < lens[:, None]).long()
        want = oh.pool(logits.float(), mask, pooling).numpy()
        got = A.splade_pool(logits.cuda(), mask.cuda(), pooling).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5 if pooling == "sum" else POOL_RTOL, atol=1e-7)


def test_splade_mirror_forward_and_search():
    """The class mirror: forward() with a stub encoder equals the oracle; its CSR output scores like the dense path."""
    from fusion_b200.retrievers.splade import SPLADE
    from oracle import splade_head as oh
    g = torch.Generator().manual_seed(9)
    logits = torch.randn((40, 12, 4100), generator=g) * 2 - 1.5
    mask = torch.ones((40, 12), dtype=torch.long)
    mask[::3, 6:] = 0

    class Enc(torch.nn.Module):
        def forward(self, input_ids=None, attention_mask=None):
            return types.SimpleNamespace(logits=logits.cuda())

    with pytest.raises(AssertionError):
        SPLADE(Enc(), pooling="mean")
    m = SPLADE(Enc(), pooling="max", pruning_topk=20)
    act = m.forward(None, mask.cuda())
    want_pool = oh.pool(logits, mask, "max")
    want, _ = oh.prune(want_pool, 20)
    np.testing.assert_allclose(act.cpu().numpy(), want.numpy(), rtol=POOL_RTOL, atol=1e-7)
    assert int((act != 0).sum(1).max()) <= 20
    ptr, term, w = m.encode_csr(None, mask.cuda())
    dense = torch.zeros_like(act)
    dense[torch.repeat_interleave(torch.arange(40, device="cuda"), ptr[1:] - ptr[:-1]), term.long()] = w
    assert torch.equal(dense, act)
    # sparse search over those activations == exact dense cosine of the same vectors
    sc, ids = m.search_tensors(act[:5], act, 10)
    qn = torch.nn.functional.normalize(act[:5], dim=1)
    dn = torch.nn.functional.normalize(act, dim=1)
    ref_s, ref_i = torch.topk(qn @ dn.t(), 10, dim=1)
    np.testing.assert_allclose(sc.cpu().numpy(), ref_s.cpu().numpy(), rtol=2e-5, atol=1e-6)
    assert (ids.cpu()[:, 0] == torch.arange(5)).all()                     # every vector's nearest neighbour is itself
