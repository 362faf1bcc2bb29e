"""Whitespace tokenisation on the device: the step in front of the lexical index build.

The reference tokenises with ``doc.split()`` inside Python loops over the corpus (``src/retrievers/bm25.py:54-60,72,81``,
4.9 s for 28k documents); here the corpus is ONE UTF-8 byte buffer on the device, ``fz_token_starts`` marks the token
starts with ``str.split()``'s whitespace rules, ``fz_hash_tokens`` hashes every token to 128 bits, and the distinct hashes
(a device sort) are the vocabulary.  Term ids are positions in that sorted hash table: they differ from the reference's
first-appearance numbering, which never leaves the index.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FusionB200Error, check
from .ops import _ptr, _stream


def _buffer(texts: list[str], device) -> tuple[torch.Tensor, torch.Tensor]:
    """-> (uint8 bytes of '\\n'.join(texts) + '\\n' on the device, int64 start offset of every text [n + 1])."""
    enc = [t.encode("utf-8") for t in texts]
    lens = np.fromiter((len(e) + 1 for e in enc), dtype=np.int64, count=len(enc))
    ptr = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    raw = b"\n".join(enc) + b"\n" if enc else b""
    buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8) if raw else torch.zeros(0, dtype=torch.uint8)
    return buf.to(device), torch.from_numpy(ptr).to(device)


def hash_tokens(texts: list[str], device="cuda"):
    """Tokenise ``texts`` -> (ptr int64 [n+1] tokens per text as CSR, h1 int64 [T], h2 int64 [T])."""
    lib = _lib.load()
    buf, text_ptr = _buffer(texts, device)
    n = buf.numel()
    flags = torch.empty(n, dtype=torch.uint8, device=buf.device)
    check(lib.fz_token_starts(_ptr(buf), n, _ptr(flags), _stream(flags)), "fz_token_starts")
    starts = torch.nonzero(flags).flatten()
    nt = starts.numel()
    h1 = torch.empty(nt, dtype=torch.int64, device=buf.device)
    h2 = torch.empty(nt, dtype=torch.int64, device=buf.device)
    check(lib.fz_hash_tokens(_ptr(buf), n, _ptr(starts), nt, _ptr(h1), _ptr(h2), None, _stream(h1)), "fz_hash_tokens")
    # a text's tokens start inside [text_ptr[i], text_ptr[i+1]): the separator '\n' is whitespace, tokens never span texts
    ptr = torch.searchsorted(starts, text_ptr)
    return ptr, h1, h2


class DeviceVocabulary:
    """Distinct token hashes of a corpus, sorted by the first hash: term id = position.  The second hash only guards
    against collisions of the first (two different tokens with equal h1 would need equal h2 too to be merged)."""

    def __init__(self, h1: torch.Tensor, h2: torch.Tensor):
        self.h1, inverse = torch.unique(h1, return_inverse=True)
        n = self.h1.numel()
        lo = torch.full((n,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=h1.device).scatter_reduce_(0, inverse, h2, "amin")
        hi = torch.full((n,), torch.iinfo(torch.int64).min, dtype=torch.int64, device=h1.device).scatter_reduce_(0, inverse, h2, "amax")
        if bool((lo != hi).any()):
            raise FusionB200Error("64-bit token hash collision inside the corpus vocabulary; re-tokenise on the host")
        self.h2 = lo
        self.inverse = inverse

    def __len__(self) -> int:
        return self.h1.numel()

    def lookup(self, h1: torch.Tensor, h2: torch.Tensor) -> torch.Tensor:
        """term id of every (h1, h2), -1 when the token is not in the vocabulary."""
        if h1.numel() == 0 or len(self) == 0:
            return torch.full((h1.numel(),), -1, dtype=torch.int32, device=h1.device)
        idx = torch.searchsorted(self.h1, h1).clamp(max=len(self) - 1)
        ok = (self.h1[idx] == h1) & (self.h2[idx] == h2)
        return torch.where(ok, idx, torch.full_like(idx, -1)).to(torch.int32)


def tokenize_corpus(corpus: list[str], device="cuda"):
    """-> (vocabulary, doc_ptr int64 [N+1], doc_tok int32 [T]) with the token sequence of every document in order."""
    ptr, h1, h2 = hash_tokens(corpus, device)
    vocab = DeviceVocabulary(h1, h2)
    return vocab, ptr, vocab.inverse.to(torch.int32)


def tokenize_queries(queries: list[str], vocab: DeviceVocabulary, device="cuda"):
    """-> (q_ptr int32 [Q+1], q_term int32) in query-token order, duplicates kept, -1 = out of vocabulary."""
    ptr, h1, h2 = hash_tokens(queries, device)
    return ptr.to(torch.int32), vocab.lookup(h1, h2)


def percentile_distribution(scores: torch.Tensor, n_points: int) -> torch.Tensor:
    """The percentile-based score distribution of one system (src/retrievers/hybrid.py:391-398): drop the zeros and every
    occurrence of the two smallest distinct scores, then ``quantile(np.linspace(0, 1, n_points + 1))`` (linear
    interpolation) -> float64 [n_points + 1], ascending: the ``percentile_distr`` input of the percentile-rank fusion."""
    lib = _lib.load()
    s = scores.flatten().to(torch.float64)
    if s.numel():
        uniq = torch.unique(s)                       # ascending; drop_duplicates().nsmallest(2)
        smallest = uniq[:2]
        keep = (s != 0.0) & ~torch.isin(s, smallest)
        s = s[keep]
    if s.numel() == 0:
        return torch.full((n_points + 1,), float("nan"), dtype=torch.float64, device=scores.device)
    s = torch.sort(s).values.contiguous()
    out = torch.empty(n_points + 1, dtype=torch.float64, device=s.device)
    check(lib.fz_quantiles_f64(_ptr(s), s.numel(), n_points + 1, _ptr(out), _stream(out)), "fz_quantiles_f64")
    return out
