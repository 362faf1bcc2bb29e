"""ctypes binding of libfusion_b200.so (the C ABI declared in include/fusion_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfusion_b200.so")

FZ_STATUS_OVERFLOW, FZ_STATUS_NEED_ZERO, FZ_STATUS_NEED_NEG, FZ_STATUS_FALLBACK, FZ_STATUS_TOO_LONG = 1, 2, 4, 8, 16
FUSE_METHODS = {"bcf": 0, "rrf": 1, "nsf": 2}
FUSE_NORMS = {None: 0, "none": 0, "min-max": 1, "z-score": 2, "arctan": 3, "percentile-rank": 4,
              "normal-curve-equivalent": 5, "identity-f32": 6}
LEX_TFIDF, LEX_BM25 = 0, 1
ABI_VERSION = 2         # FZ_ABI_VERSION


class FusionB200Error(RuntimeError):
    pass


class Postings(C.Structure):
    """mirror of fz_postings_t"""
    _fields_ = [("term_ptr", C.c_void_p), ("post_doc", C.c_void_p), ("post_val", C.c_void_p),
                ("short_coarse", C.c_void_p), ("term_slot", C.c_void_p), ("tiled_base", C.c_void_p), ("tiled_tile_off", C.c_void_p),
                ("tiled_off", C.c_void_p), ("tiled_val", C.c_void_p), ("dense_val", C.c_void_p),
                ("dense_stride", C.c_int64), ("n_terms", C.c_int32), ("n_tiled", C.c_int32), ("n_dense", C.c_int32),
                ("tile_docs", C.c_int32), ("n_tiles", C.c_int32), ("n_coarse", C.c_int32), ("n_docs", C.c_int64)]


class BuildPlan(C.Structure):
    """mirror of fz_build_plan_t"""
    _fields_ = [("n_short", C.c_int64), ("n_tiled_entries", C.c_int64), ("dense_stride", C.c_int64), ("n_tiled", C.c_int32),
                ("n_dense", C.c_int32), ("n_tiles", C.c_int32), ("n_coarse", C.c_int32)]


class SpladeHead(C.Structure):
    """mirror of fz_splade_head_t"""
    _fields_ = [("head_bf16", C.c_void_p), ("term_head", C.c_void_p), ("term_max", C.c_void_p), ("doc_ptr", C.c_void_p),
                ("doc_post", C.c_void_p), ("head_dim", C.c_int32), ("n_terms", C.c_int32), ("n_docs", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


SHARD_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p)      # fz_shard_hook_t


class ShardSync(C.Structure):
    """mirror of fz_shard_sync_t"""
    _fields_ = [("hook", SHARD_HOOK), ("user", C.c_void_p), ("exchange", C.c_void_p), ("n_shards", C.c_int32),
                ("floor_rank", C.c_int32), ("sched_docs", C.c_int64)]


_p, _i, _i64, _sz, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol of include/fusion_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "fz_last_error": (C.c_char_p, []),
    "fz_abi_version": (_i, []),
    "fz_profile_enable": (_i, [_i]),
    "fz_profile_summary": (_i, [C.c_char_p, _sz]),
    "fz_debug_set_stats": (_i, [_p]),
    "fz_merge_topk_workspace_bytes": (_sz, [_i, _i, _i]),
    "fz_merge_topk_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "fz_merge_topk_f64": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "fz_rank_rows_workspace_bytes": (_sz, [_i, _i64]),
    "fz_rank_rows_f32": (_i, [_p, _i, _i64, _i, _i64, _p, _p, _p, _sz, _p]),
    "fz_rank_rows_f64": (_i, [_p, _i, _i64, _i, _i64, _p, _p, _p, _sz, _p]),
    "fz_fuse_workspace_bytes": (_sz, [_i, _i, _p]),
    "fz_fuse": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "fz_rank_metrics": (_i, [_p, _p, _i, _i, _p, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _p, _p]),
    "fz_fuse_sweep": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _p]),
    "fz_token_starts": (_i, [_p, _i64, _p, _p]),
    "fz_hash_tokens": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p]),
    "fz_quantiles_f64": (_i, [_p, _i64, _i, _p, _p]),
    "fz_splade_pool": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p]),
    "fz_prune_topk": (_i, [_p, _i, _i, _i, _p, _p]),
    "fz_csr_count": (_i, [_p, _i, _i, _p, _p]),
    "fz_csr_fill": (_i, [_p, _i, _i, _p, _p, _p, _p]),
    "fz_build_lexical_workspace_bytes": (_sz, [_i64, C.c_int32]),
    "fz_build_lexical_plan": (_i, [_p, _p, _i64, _i64, C.c_int32, _p, _p, _sz, _p]),
    "fz_build_lexical_fill": (_i, [_p, _i64, _i64, C.c_int32, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fz_build_term_major_workspace_bytes": (_sz, [_i64, C.c_int32]),
    "fz_build_term_major": (_i, [_p, _p, _i64, _i64, C.c_int32, _p, _p, _p, _sz, _p]),
    "fz_build_postings_workspace_bytes": (_sz, [C.c_int32]),
    "fz_build_postings_plan": (_i, [_p, _p, C.c_int32, _i64, C.c_int32, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "fz_build_postings_fill": (_i, [_p, _p, _p, _i, C.c_int32, _i64, C.c_int32, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p,
                                    _p, _sz, _p]),
    "fz_build_csr_normalize": (_i, [_p, _p, _i64, _p, _p]),
    "fz_build_term_stats": (_i, [_p, _p, _i64, C.c_int32, _p, _p, _p, _p]),
    "fz_build_splade_head": (_i, [_p, _p, _p, _i64, _i64, _p, C.c_int32, _p, _p]),
    "fz_lexical_impacts": (_i, [_p, _p, _p, _p, _p, C.c_int32, _i64, _d, _d, _d, _i, _p, _p]),
    "fz_sparse_topk_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "fz_sparse_topk_f64": (_i, [_p, _p, _p, _i, _i, _i64, _i, _i, _i, _p, _p, _p, _p, _sz, _p, _p]),
    "fz_sparse_topk_f32": (_i, [_p, _p, _p, _p, _i, _i, _i64, _i, _i, _i, _p, _p, _p, _p, _sz, _p, _p]),
    "fz_sparse_scores_f64": (_i, [_p, _p, _p, _i, _p, _i, _p]),
    "fz_sparse_scores_f32": (_i, [_p, _p, _p, _p, _i, _p, _i, _p]),
    "fz_splade_topk_workspace_bytes": (_sz, [_i, _i, _i, _i, _i64]),
    "fz_splade_topk": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _i, _p, _p, _p, _p, _sz, _p, _p]),
    "fz_dense_topk_workspace_bytes": (_sz, [_i, _i, _i]),
    "fz_dense_topk": (_i, [_p, _p, _p, _p, _i, _i64, _i, _i, _f, _i64, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "fz_dense_topk_filter": (_i, [_p, _p, _i, _i64, _i, _i, _f, _i64, _i, _i, _i, _p, _p, _p, _sz, _p, _p]),
    "fz_dense_topk_finish": (_i, [_p, _p, _p, _i, _i, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    "fz_dense_scores_f32": (_i, [_p, _p, _i, _i64, _i, _p, _p]),
    "fz_pairwise_dot_f32": (_i, [_p, _p, _i64, _i, _p, _p]),
    "fz_normalize_rows": (_i, [_p, _i64, _i, _i, _p, _p, _p]),
    "fz_maxsim_workspace_bytes": (_sz, [_i, _i]),
    "fz_maxsim_pack": (_i, [_p, _p, _p, _i64, _p, _p]),
    "fz_maxsim_bf16": (_i, [_p, _i, _p, _p, _p, _p, _i64, _i64, _i, _i, _p, _p, _sz, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FusionB200Error(
                f"{LIB_PATH} is missing - build it with `python -m fusion_b200.build` (nvcc, sm_100a). "
                "fusion_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.fz_abi_version() != ABI_VERSION:
            raise FusionB200Error("libfusion_b200.so ABI version mismatch - rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fz_last_error().decode(errors="replace")
        raise FusionB200Error(f"{what} failed ({rc}): {msg}")
