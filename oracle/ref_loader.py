"""Import the UNMODIFIED reference modules from /root/reference (build container only).

The GPU box has no /root/reference: everything that must run there uses the committed fixtures in
``tests/golden`` instead.  Three absent third-party names are stubbed with empty modules and one
removed transformers attribute is restored (SURVEY.md preamble probe table); no reference source is
edited or copied.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOTS = ("/root/reference",)


def reference_root() -> str | None:
    for r in REF_ROOTS:
        if os.path.isdir(os.path.join(r, "src", "retrievers")):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


def _stub(name: str, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    return sys.modules[name]


def _prepare():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected on the build container only)")
    if root not in sys.path:
        sys.path.insert(0, root)
    _stub("ir_datasets")
    _stub("seaborn")
    sp = _stub("spacy")
    tok = _stub("spacy.tokens", Doc=type("Doc", (), {}))
    sp.tokens = tok
    return root


def load_bm25():
    """-> module ``src.retrievers.bm25`` (classes TFIDF, BM25, AtireBM25), verbatim."""
    _prepare()
    import importlib
    return importlib.import_module("src.retrievers.bm25")


def load_hybrid():
    """-> module ``src.retrievers.hybrid`` (Ranker, Aggregator), verbatim."""
    _prepare()
    import importlib
    return importlib.import_module("src.retrievers.hybrid")


def load_splade_base():
    """-> module ``src.retrievers.splade.base`` (BaseModel), verbatim, after restoring
    ``transformers.file_utils.default_cache_path`` (removed in transformers 5)."""
    _prepare()
    import importlib
    import transformers.file_utils as fu
    if not hasattr(fu, "default_cache_path"):
        fu.default_cache_path = os.path.expanduser("~/.cache/huggingface")
    return importlib.import_module("src.retrievers.splade.base")


def load_splade():
    """-> module ``src.retrievers.splade.splade`` (class SPLADE), verbatim.  Two names the module imports no longer
    exist where it looks for them: ``transformers.optimization.AdamW`` (removed in transformers 5; only used by
    ``fit``) and ``BaseModel`` / ``MmarcoReader`` on the ``splade`` package (its ``__init__`` exports nothing); both
    are set on the imported modules, the reference sources stay untouched."""
    base = load_splade_base()
    import importlib
    import torch
    import transformers.optimization as opt
    if not hasattr(opt, "AdamW"):
        opt.AdamW = torch.optim.AdamW
    pkg = importlib.import_module("src.retrievers.splade")
    if not hasattr(pkg, "BaseModel"):
        pkg.BaseModel = base.BaseModel
    if not hasattr(pkg, "MmarcoReader"):
        pkg.MmarcoReader = type("MmarcoReader", (), {})
    return importlib.import_module("src.retrievers.splade.splade")


def make_splade_head(logits, pooling: str, pruning_topk):
    """A verbatim ``SPLADE`` whose encoder is a stub returning ``logits``: ``forward`` (splade.py:80-99) and
    ``_prune_activations`` (:295-306) then run unmodified."""
    import types
    import torch
    mod = load_splade()

    class _Encoder(torch.nn.Module):
        def forward(self, input_ids=None, attention_mask=None):
            return types.SimpleNamespace(logits=logits)

    m = object.__new__(mod.SPLADE)
    torch.nn.Module.__init__(m)
    m.model, m.pooling, m.pruning_topk, m.relu, m.device = _Encoder(), pooling, pruning_topk, torch.nn.ReLU(), "cpu"
    return m


def make_injected_searcher(similarity: str, q_embs, d_embs):
    """A verbatim ``BaseModel`` whose ``encode`` returns injected tensors, so that
    ``BaseModel.search`` (splade/base.py:199-251) runs unmodified on synthetic embeddings."""
    base = load_splade_base()

    class _Injected(base.BaseModel):
        def __init__(self):
            self.similarity = similarity

        def forward(self, *a, **k):  # pragma: no cover
            raise NotImplementedError

        def fit(self, *a, **k):  # pragma: no cover
            raise NotImplementedError

        def encode(self, sentences, query_mode=True, **kwargs):
            return q_embs if query_mode else d_embs

    return _Injected()
