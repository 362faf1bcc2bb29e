"""Index build through the C-ABI (fz_build_*, csrc/build.cu) against independent implementations: the torch layout
specification (tests/_torch_postings.py) array for array, numpy for the token -> (term, doc, tf) postings, and the
reference's own index statistics (bm25.py:53-83: vocabulary df, per-doc term counts, doc lengths)."""
import numpy as np
import pytest
import torch

from fusion_b200 import ops, synth
from fusion_b200._lib import FusionB200Error
from fusion_b200.index import SparseIndex, build_postings

import _torch_postings as spec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _csr(n_docs, vocab, seed, mean_len=40, zipf=1.3):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 2 * mean_len, n_docs)
    doc = np.repeat(np.arange(n_docs), lens)
    term = np.minimum((rng.zipf(zipf, doc.size) - 1), vocab - 1)
    key = np.unique(term.astype(np.int64) * n_docs + doc)
    term, doc = key // n_docs, key % n_docs
    val = rng.random(key.size) + 0.1
    ptr = np.zeros(vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(term, minlength=vocab), out=ptr[1:])
    return torch.from_numpy(ptr), torch.from_numpy(doc.astype(np.int32)), torch.from_numpy(val)


@pytest.mark.parametrize("n_docs,tile,dtype,tiled_q,dense_frac", [
    (5000, 512, torch.float64, 40, 0.3), (3000, 1024, torch.float32, 40, 0.3), (700, 256, torch.float64, 40, 0.3),
    (40000, 2048, torch.float32, 10, 0.0), (9001, 4, torch.float32, 50, 0.25), (70000, 8192, torch.float32, 30, 0.0),
    (2500, 32768, torch.float64, 60, 0.5)])
def test_postings_layout_equals_specification(n_docs, tile, dtype, tiled_q, dense_frac):
    vocab = 300
    ptr, doc, val = _csr(n_docs, vocab, seed=n_docs)
    df = np.diff(ptr.numpy())
    tiled_min = min(int(np.percentile(df, tiled_q)) + 1, 65535)
    want = spec.build_postings(ptr, doc, val.to(dtype), n_docs, tile, tiled_min=tiled_min, dense_frac=dense_frac)
    got = build_postings(ptr.to(DEV), doc.to(DEV), val.to(dtype).to(DEV), n_docs, tile, tiled_min=tiled_min, dense_frac=dense_frac)
    assert (got.n_docs, got.tile_docs) == (want.n_docs, want.tile_docs)
    names = ("term_ptr", "post_doc", "post_val", "short_coarse", "term_slot", "tiled_base", "tiled_tile_off", "tiled_off",
             "tiled_val", "dense_val")
    for name, g, w in zip(names, got.tensors(), want.tensors()):
        assert g.dtype == w.dtype and tuple(g.shape) == tuple(w.shape), name
        assert torch.equal(g.cpu(), w), name


def test_postings_of_an_empty_and_an_all_short_index():
    ptr = torch.zeros(11, dtype=torch.int64, device=DEV)
    pv = build_postings(ptr, torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.float32, device=DEV), 100, 64)
    assert pv.post_doc.numel() == 0 and pv.tiled_base.numel() == 0 and int(pv.term_ptr[-1]) == 0
    assert tuple(pv.short_coarse.shape) == (10, 2) and not bool(pv.short_coarse.any())
    p, d, v = _csr(400, 50, seed=9)
    pv = build_postings(p.to(DEV), d.to(DEV), v.to(DEV), 400, 128, tiled_min=65535, dense_frac=0.0)
    assert pv.tiled_base.numel() == 0 and pv.dense_val.shape[0] == 0 and int(pv.term_ptr[-1]) == d.numel()
    with pytest.raises(FusionB200Error):
        build_postings(p.to(DEV), d.to(DEV), v.to(DEV), 400, 130)          # tile_docs not a multiple of 4


def test_term_major_transpose():
    rng = np.random.default_rng(4)
    n_rows, n_terms, nnz = 3000, 500, 60000
    key = rng.choice(n_rows * n_terms, nnz, replace=False)
    rng.shuffle(key)
    row, term = (key % n_rows).astype(np.int32), (key // n_rows).astype(np.int32)
    order, ptr = ops.build_term_major(torch.from_numpy(row).to(DEV), torch.from_numpy(term).to(DEV), n_rows, n_terms)
    want = np.argsort(term.astype(np.int64) * n_rows + row, kind="stable")
    assert np.array_equal(order.cpu().numpy(), want)
    assert np.array_equal(ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(term, minlength=n_terms))]))
    o, p = ops.build_term_major(torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV), 10, 7)
    assert o.numel() == 0 and not bool(p.any()) and p.numel() == 8
    with pytest.raises(FusionB200Error):
        ops.build_term_major(torch.tensor([0, 11], dtype=torch.int32, device=DEV), torch.tensor([1, 2], dtype=torch.int32, device=DEV), 10, 7)


def test_lexical_postings_from_tokens_match_the_reference_statistics():
    """bm25.py:53-83: vocab[t] = number of docs containing t, doc_term_freqs[d][t] = occurrences, doc_len[d] = tokens."""
    (dptr, dtok), _ = synth.c3_lexical(4000, 4, 900)
    tp, pd, tf, dl = ops.build_lexical_postings(torch.from_numpy(dptr).to(DEV), torch.from_numpy(dtok.astype(np.int32)).to(DEV), 900)
    tp, pd, tf, dl = tp.cpu().numpy(), pd.cpu().numpy(), tf.cpu().numpy(), dl.cpu().numpy()
    vocab, freqs = {}, []
    for d in range(4000):                                   # the reference's loop, on token ids
        toks = dtok[dptr[d]:dptr[d + 1]].tolist()
        f = {}
        for t in toks:
            f[t] = f.get(t, 0) + 1
        for t in f:
            vocab[t] = vocab.get(t, 0) + 1
        freqs.append(f)
        assert dl[d] == len(toks)
    assert tp[0] == 0 and tp[-1] == pd.size == sum(len(f) for f in freqs)
    for t in range(900):
        docs = pd[tp[t]:tp[t + 1]]
        assert docs.size == vocab.get(t, 0)
        assert np.all(np.diff(docs) > 0)
        for d, c in zip(docs.tolist(), tf[tp[t]:tp[t + 1]].tolist()):
            assert freqs[d][t] == c
    with pytest.raises(FusionB200Error):
        ops.build_lexical_postings(torch.tensor([0, 2], device=DEV), torch.tensor([1, 900], dtype=torch.int32, device=DEV), 900)
    e = ops.build_lexical_postings(torch.zeros(4, dtype=torch.int64, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV), 5)
    assert e[1].numel() == 0 and not bool(e[0].any()) and e[3].tolist() == [0, 0, 0]


def test_splade_head_stats_and_normalisation():
    ptr, term, w = synth.splade_vectors(3000, 800, 40, 4, 120, seed=5)
    dp = torch.from_numpy(ptr).to(DEV)
    t = torch.from_numpy(term.astype(np.int32)).to(DEV)
    wt = torch.from_numpy(w.astype(np.float32)).to(DEV)
    nw = ops.build_csr_normalize(dp, wt).cpu().numpy()
    row = np.repeat(np.arange(3000), np.diff(ptr))
    nrm = np.sqrt(np.bincount(row, weights=w.astype(np.float64) ** 2, minlength=3000))
    assert np.allclose(nw, w / np.maximum(nrm, 1e-12)[row], rtol=3e-7, atol=0)
    df, tmax, flags = ops.build_term_stats(t, wt, 800)
    assert flags == 0
    assert np.array_equal(df.cpu().numpy(), np.bincount(term, minlength=800))
    want_max = np.zeros(800, dtype=np.float32)
    np.maximum.at(want_max, term, w.astype(np.float32))
    assert np.array_equal(tmax.cpu().numpy(), want_max)
    assert ops.build_term_stats(t, -wt, 800)[2] & 1
    assert ops.build_term_stats(t + 1, wt, 800)[2] & 2 or int(term.max()) < 799
    term_head = torch.full((800,), -1, dtype=torch.int32, device=DEV)
    top = torch.topk(df, 64).indices
    term_head[top] = torch.arange(64, dtype=torch.int32, device=DEV)
    head = ops.build_splade_head(dp, t, wt, term_head, 64).float().cpu().numpy()
    want = np.zeros((3000, 64), dtype=np.float32)
    th = term_head.cpu().numpy()[term]
    m = th >= 0
    want[row[m], th[m]] = torch.from_numpy(w.astype(np.float32)[m]).bfloat16().float().numpy()
    assert np.array_equal(head, want)
    # an index built from the same vectors scores like before (the builders feed the scoring kernels end to end)
    ix = SparseIndex(ptr, term, w, 800, "cos_sim", device=DEV, head_dim=64, boot_docs=0)
    assert ix.head is not None and ix.nonneg
