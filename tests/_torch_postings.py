"""Layout specification of ``fz_postings_t`` in plain torch (runs on CPU): the independent implementation the tests
compare ``fz_build_postings_*`` / ``fz_build_term_major`` against, array for array.  Test infrastructure only."""
import math
import os

import numpy as np
import torch

from fusion_b200 import ops
from fusion_b200._lib import FusionB200Error

DENSE_FRAC = 0.25


def _term_major_csr(row_of_entry: torch.Tensor, term_of_entry: torch.Tensor, n_rows: int, n_terms: int):
    """Sort (term, doc) pairs term-major / doc-ascending.  -> (order, term_ptr int64 [V+1])."""
    key = term_of_entry.to(torch.int64) * n_rows + row_of_entry.to(torch.int64)
    order = torch.argsort(key)
    counts = torch.bincount(term_of_entry.to(torch.int64), minlength=n_terms)
    term_ptr = torch.zeros(n_terms + 1, dtype=torch.int64, device=key.device)
    term_ptr[1:] = torch.cumsum(counts, 0)
    return order, term_ptr, counts


def build_postings(term_ptr: torch.Tensor, post_doc: torch.Tensor, post_val: torch.Tensor, n_docs: int, tile_docs: int,
                   tiled_min: int | None = None, dense_frac: float = DENSE_FRAC,
                   chunk_postings: int = 1 << 27) -> ops.PostingsView:
    """Term-major CSR (doc-ascending inside a term) -> the three storage forms of ``fz_postings_t``.

    short  df < tiled_min: kept as (doc, value) pairs.
    tiled  per (term, tile) segments of (uint16 tile-relative offset, value), padded to 4 with (tile_docs, 0) and ordered
           for conflict-free shared-memory scatter: the postings of a segment are dealt round-robin over the 32 banks
           (doc % 32), and that sequence is laid out so that lane l of a warp reads element l of a run of 32 with its
           j-th accumulate (a thread owns the postings of one 16-byte value vector: 4 fp32 or 2 fp64).
    dense  df >= dense_frac * n_docs: one value per document, zero where the term is absent.
    """
    dev = post_doc.device
    n_terms = term_ptr.numel() - 1
    vec = 16 // post_val.element_size()                     # 4 fp32 weights / 2 fp64 impacts per 16-byte load
    if tile_docs % 4 or not (4 <= tile_docs <= 32768):      # (the K2 kernels take <= 8192; the SPLADE tail kernel 32768)
        raise FusionB200Error(f"tile_docs={tile_docs} must be a multiple of 4 in [4, 32768]")
    n_tiles = (n_docs + tile_docs - 1) // tile_docs
    df = term_ptr[1:] - term_ptr[:-1]
    if tiled_min is None:
        tiled_min = int(os.environ.get("FZ_TILED_MIN", 512))
    dense_min = max(tiled_min, int(math.ceil(dense_frac * n_docs))) if dense_frac > 0 else (1 << 62)
    is_dense = df >= dense_min
    is_tiled = (df >= tiled_min) & ~is_dense
    tiled_terms = torch.nonzero(is_tiled).flatten()
    dense_terms = torch.nonzero(is_dense).flatten()
    n_tiled, n_dense = tiled_terms.numel(), dense_terms.numel()
    term_slot = torch.full((n_terms,), -1, dtype=torch.int32, device=dev)
    term_slot[tiled_terms] = torch.arange(n_tiled, dtype=torch.int32, device=dev)
    term_slot[dense_terms] = -2 - torch.arange(n_dense, dtype=torch.int32, device=dev)
    slot_of_post = torch.repeat_interleave(term_slot, df)                       # int32 per posting

    # ---- short lists
    m = slot_of_post == -1
    short_doc, short_val = post_doc[m].contiguous(), post_val[m].contiguous()
    short_ptr = torch.zeros(n_terms + 1, dtype=torch.int64, device=dev)
    short_ptr[1:] = torch.cumsum(torch.where(is_tiled | is_dense, torch.zeros_like(df), df), 0)
    if tiled_min > 65535:
        raise FusionB200Error("tiled_min must be <= 65535 (short-list offsets are 16 bits)")
    # coarse marks: postings of the term below tile 64*c, so the kernel bisects a handful of postings, not the list
    n_coarse = (n_tiles + ops.COARSE_TILES - 1) // ops.COARSE_TILES
    st = torch.repeat_interleave(torch.arange(n_terms, device=dev), short_ptr[1:] - short_ptr[:-1])
    bucket = short_doc.long() // (ops.COARSE_TILES * tile_docs)
    cnt = torch.bincount(st * n_coarse + bucket, minlength=n_terms * n_coarse).view(n_terms, n_coarse)
    coarse = torch.zeros((n_terms, n_coarse + 1), dtype=torch.int64, device=dev)
    coarse[:, 1:] = torch.cumsum(cnt, 1)
    short_coarse = torch.where(coarse >= (1 << 15), coarse - (1 << 16), coarse).to(torch.int16)     # uint16 payload
    del st, bucket, cnt, coarse

    # ---- dense rows
    stride = n_tiles * tile_docs
    dense_val = torch.zeros((n_dense, stride), dtype=post_val.dtype, device=dev)
    if n_dense:
        m = slot_of_post <= -2
        dense_val[(-2 - slot_of_post[m]).long(), post_doc[m].long()] = post_val[m]

    # ---- tiled segments
    tiled_base = torch.zeros(n_tiled, dtype=torch.int64, device=dev)
    tile_off = torch.zeros((n_tiled, n_tiles + 1), dtype=torch.int32, device=dev)
    tiled_off = torch.zeros(0, dtype=torch.int16, device=dev)
    tiled_val = torch.zeros(0, dtype=post_val.dtype, device=dev)
    if n_tiled:
        m = slot_of_post >= 0
        r_all = slot_of_post[m].long()
        d_all = post_doc[m].long()
        v_all = post_val[m]
        del m, slot_of_post
        seg_len = torch.bincount(r_all * n_tiles + d_all // tile_docs, minlength=n_tiled * n_tiles)
        seg_pad = (seg_len + 3) // 4 * 4
        seg_start = torch.zeros(n_tiled * n_tiles + 1, dtype=torch.int64, device=dev)
        seg_start[1:] = torch.cumsum(seg_pad, 0)
        seg_first = torch.cumsum(seg_len, 0) - seg_len                              # first unpadded posting of a segment
        total = int(seg_start[-1])
        tiled_base = seg_start[:-1:n_tiles].clone()
        rel = seg_start.view(-1)[: n_tiled * n_tiles].view(n_tiled, n_tiles) - tiled_base[:, None]
        last = seg_start[n_tiles::n_tiles] - tiled_base
        if int(torch.max(last)) >= (1 << 32):
            raise FusionB200Error("a tiled posting list exceeds 2^32 entries")
        tile_off = torch.cat([rel, last[:, None]], dim=1).to(torch.int64)
        tile_off = torch.where(tile_off >= (1 << 31), tile_off - (1 << 32), tile_off).to(torch.int32)   # uint32 payload
        tiled_off = torch.full((total,), tile_docs if tile_docs < 32768 else tile_docs - 65536, dtype=torch.int16, device=dev)   # uint16 payload
        tiled_val = torch.zeros(total, dtype=post_val.dtype, device=dev)
        # chunk over term rows so the sort temporaries stay bounded
        df_t = df[tiled_terms]
        row_end = torch.cumsum(df_t, 0).cpu().numpy()
        r0, p0 = 0, 0
        while r0 < n_tiled:
            r1 = int(np.searchsorted(row_end, p0 + chunk_postings, side="right"))
            r1 = min(n_tiled, max(r1, r0 + 1))
            p1 = int(row_end[r1 - 1])
            r, d, v = r_all[p0:p1], d_all[p0:p1], v_all[p0:p1]
            tile = d // tile_docs
            off = d - tile * tile_docs
            seg = r * n_tiles + tile                                                # global segment id, non-decreasing
            bank = off & 31
            o1 = torch.sort(seg * 32 + bank, stable=True).indices                  # (segment, bank), offsets ascending
            sb = (seg * 32 + bank)[o1]
            pos = torch.arange(p1 - p0, device=dev)
            is_start = torch.ones_like(sb, dtype=torch.bool)
            is_start[1:] = sb[1:] != sb[:-1]
            start_pos = torch.cummax(torch.where(is_start, pos, torch.zeros_like(pos)), 0).values
            rank_in_bank = torch.empty_like(pos)
            rank_in_bank[o1] = pos - start_pos
            del sb, is_start, start_pos, o1
            o2 = torch.sort(((seg << 16) | rank_in_bank) * 32 + bank).indices      # round-robin over the banks
            q = torch.empty_like(pos)
            q[o2] = pos - (seg_first[seg[o2]] - p0)
            del o2, rank_in_bank
            # a thread owns `vec` consecutive postings (one 16-byte vector of values): element e of the round-robin
            # sequence goes to slot vec * (e mod P/vec) + e div (P/vec), so the lanes' j-th accumulates read a run of it
            part = seg_pad[seg] // vec
            dest = seg_start[seg] + vec * (q % part) + q // part
            tiled_off[dest] = off.to(torch.int16)
            tiled_val[dest] = v
            r0, p0 = r1, p1
    return ops.PostingsView(short_ptr, short_doc, short_val, short_coarse, term_slot, tiled_base, tile_off, tiled_off, tiled_val,
                            dense_val, n_docs, tile_docs)
