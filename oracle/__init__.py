"""CPU oracle for the retrieval-scoring + fusion hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / torch-CPU, the algorithms of maastrichtlawtech/fusion that the
CUDA library replaces.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package ``fusion_b200``
never does (``tests/test_no_oracle_in_product.py`` enforces that).

Pinning status: the reference ships no tests, golden vectors or fixtures.
  * BM25 / TF-IDF / ATIRE (``oracle.bm25``) and rank fusion (``oracle.fusion``) are pinned
    bit-for-bit / to 1e-6 against the reference's own classes executed verbatim in the build
    container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
  * dense score + top-k (``oracle.dense``) is pinned against the verbatim
    ``src/retrievers/splade/base.py::BaseModel.search`` with injected embeddings.
  * ``sentence_transformers.util.semantic_search`` (sentence-transformers==2.2.2) and ColBERT
    ``colbert_score`` (colbert-ai @ main, unpinned) are third-party and absent from the reference
    tree: those restatements follow the published algorithm and the in-repo call sites and are
    "parity unpinned" by any reference-owned test.
"""
