"""CPU checks of the three-form postings layout on its torch specification (tests/_torch_postings.py): the layout the
scoring kernels consume can be verified without a GPU - every posting present exactly once, segments padded and aligned
as the vectorised loads need, and the bank ordering doing its job.  tests/test_gpu_build.py then holds the device
builder (fz_build_postings_*) to this specification array for array."""
import numpy as np
import pytest
import torch

from fusion_b200 import synth
from _torch_postings import build_postings


def _csr(n_docs, vocab, seed, mean_len=40):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 2 * mean_len, n_docs)
    doc = np.repeat(np.arange(n_docs), lens)
    term = np.minimum((rng.zipf(1.3, doc.size) - 1), vocab - 1)
    key = np.unique(term.astype(np.int64) * n_docs + doc)
    term, doc = key // n_docs, key % n_docs
    val = rng.random(key.size) + 0.1
    ptr = np.zeros(vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(term, minlength=vocab), out=ptr[1:])
    return torch.from_numpy(ptr), torch.from_numpy(doc.astype(np.int32)), torch.from_numpy(val), term


@pytest.mark.parametrize("n_docs,tile,dtype", [(5000, 512, torch.float64), (3000, 1024, torch.float32), (700, 256, torch.float64)])
def test_every_posting_exactly_once(n_docs, tile, dtype):
    vocab = 300
    ptr, doc, val, term = _csr(n_docs, vocab, seed=n_docs)
    df = np.diff(ptr.numpy())
    pv = build_postings(ptr, doc, val.to(dtype), n_docs, tile, tiled_min=int(np.percentile(df, 40)) + 1, dense_frac=0.3)
    n_tiles = pv.n_tiles
    got = {}
    slot = pv.term_slot.numpy()
    assert (slot >= 0).any() and (slot <= -2).any() and (slot == -1).any()          # all three forms are exercised
    sp = pv.term_ptr.numpy()
    for t in range(vocab):
        if slot[t] == -1:
            for p in range(sp[t], sp[t + 1]):
                got[(t, int(pv.post_doc[p]))] = float(pv.post_val[p])
            docs_t = pv.post_doc[sp[t]:sp[t + 1]].numpy()
            assert np.all(np.diff(docs_t) > 0)                                           # ascending: the kernel bisects
            marks = pv.short_coarse[t].numpy().astype(np.int64) & 0xffff                 # postings below tile 16*c
            assert marks[0] == 0 and marks[-1] == docs_t.size
            for c in range(marks.size):
                assert marks[c] == np.searchsorted(docs_t, c * 16 * tile)
        elif slot[t] >= 0:
            r = slot[t]
            base = int(pv.tiled_base[r])
            offs = pv.tiled_tile_off[r].numpy().astype(np.int64) & 0xffffffff
            assert base % 4 == 0 and np.all(offs % 4 == 0)                            # 8 / 16 / 32-byte aligned segments
            for ti in range(n_tiles):
                seg = pv.tiled_off[base + offs[ti]: base + offs[ti + 1]].numpy().astype(np.int64) & 0xffff
                vals = pv.tiled_val[base + offs[ti]: base + offs[ti + 1]].numpy()
                real = seg < tile
                assert np.all(seg[~real] == tile) and np.all(vals[~real] == 0)        # padding goes to the dump slot
                assert (~real).sum() < 4
                d = ti * tile + seg[real]
                assert len(set(d.tolist())) == real.sum()
                for dd, vv in zip(d.tolist(), vals[real].tolist()):
                    got[(t, dd)] = vv
        else:
            row = pv.dense_val[-2 - slot[t]].numpy()
            assert row.shape[0] == n_tiles * tile
            for dd in np.nonzero(row)[0].tolist():
                got[(t, dd)] = float(row[dd])
    exp = {(int(t), int(d)): float(v) for t, d, v in zip(term, doc.numpy(), val.to(dtype).numpy())}
    assert got == exp


def test_bank_ordering_reduces_conflicts():
    """Within a segment, the 32 postings a warp touches with one accumulate (lane l: posting 4*(l + 32*i) + j) should sit
    in (almost) 32 different banks; in doc order they collide ~3 ways."""
    n_docs, tile, vocab = 40000, 2048, 64
    ptr, doc, val, term = _csr(n_docs, vocab, seed=3, mean_len=12)
    pv = build_postings(ptr, doc, val.float(), n_docs, tile, tiled_min=64, dense_frac=0.0)
    slot = pv.term_slot.numpy()
    worst, n_groups = 0.0, 0
    for t in np.nonzero(slot >= 0)[0][:8]:
        r = slot[t]
        base = int(pv.tiled_base[r])
        offs = pv.tiled_tile_off[r].numpy().astype(np.int64) & 0xffffffff
        for ti in range(pv.n_tiles):
            seg = pv.tiled_off[base + offs[ti]: base + offs[ti + 1]].numpy().astype(np.int64) & 0xffff
            if seg.size < 256:
                continue
            for j in range(4):          # fp32 values: a thread owns 4 consecutive postings
                sub = seg[j::4]
                for g0 in range(0, sub.size - 31, 32):
                    banks = sub[g0:g0 + 32] % 32
                    worst += np.bincount(banks, minlength=32).max()
                    n_groups += 1
    assert n_groups > 50
    assert worst / n_groups < 1.6, worst / n_groups


def test_uint32_offsets_and_empty_forms():
    ptr, doc, val, _ = _csr(400, 50, seed=9)
    pv = build_postings(ptr, doc, val, 400, 128, tiled_min=65535, dense_frac=0.0)         # everything short
    assert pv.tiled_base.numel() == 0 and pv.dense_val.shape[0] == 0
    assert int(pv.term_ptr[-1]) == doc.numel()
    st = pv.c_struct()
    assert st.n_tiled == 0 and st.n_dense == 0 and st.dense_stride == pv.n_tiles * 128
