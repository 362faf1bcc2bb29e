"""Cycle-counter probe of the persistent tensor-core kernels (debugging aid, GPU only)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from fusion_b200 import _lib, ops, synth
import numpy as np
lib = _lib.load()
dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "maxsim"
stats = torch.zeros((148, 8), dtype=torch.int64, device=dev)
names = ["prod_wait_empty", "mma_wait_tempty", "mma_wait_full", "mma_issue", "epi0_wait_tfull", "epi0_total", "epi0_pass2", "epi0_pass2_atomic_wait (plain) / pass2 count (codes)"]
g = torch.Generator(device=dev); g.manual_seed(1)
if which == "maxsim":
    nq, nc, pool = 148 * 8, 1000, 200000
    lens = torch.poisson(torch.full((pool,), 70.0, device=dev), generator=g).clamp_(8, 180).long()
    ptr = torch.zeros(pool + 1, dtype=torch.int64, device=dev); ptr[1:] = torch.cumsum(lens, 0)
    emb = torch.randn((int(ptr[-1]), 128), device=dev, generator=g)
    emb = (emb / emb.norm(dim=1, keepdim=True)).bfloat16()
    q = torch.randn((nq, 64, 128), device=dev, generator=g); q = (q / q.norm(dim=2, keepdim=True)).bfloat16()
    cand = torch.randint(0, pool, (nq, nc), device=dev, generator=g, dtype=torch.int32)
    packed = ops.pack_tokens(ptr, emb)
    for it in range(3):
        stats.zero_()
        lib.fz_debug_set_stats(ctypes.c_void_p(stats.data_ptr()))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.maxsim(q, ptr, None, cand, packed=packed); b.record(); torch.cuda.synchronize()
        lib.fz_debug_set_stats(ctypes.c_void_p(0))
    ms = a.elapsed_time(b)
    units = nq * nc / 148
    print(f"maxsim {ms:.3f} ms, {nq*nc} candidates, per-SM candidates {units:.0f}, bytes {float(lens.float().mean())*256*nq*nc/1e9:.2f} GB -> {float(lens.float().mean())*256*nq*nc/ms/1e6:.0f} GB/s")
else:
    nq, nd, dim = 6980, int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000, 768
    d16 = torch.empty((nd, dim), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, nd, 1_000_000):
        m = min(1_000_000, nd - c0)
        d16[c0:c0 + m] = ops.normalize_rows(torch.randn((m, dim), device=dev, generator=g), want_f32=False)[1]
    qq = torch.randn((nq, dim), device=dev, generator=g); _, q16 = ops.normalize_rows(qq, want_f32=False)
    for it in range(3):
        stats.zero_()
        lib.fz_debug_set_stats(ctypes.c_void_p(stats.data_ptr()))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.dense_topk(q16, d16, None, None, 1000); b.record(); torch.cuda.synchronize()
        lib.fz_debug_set_stats(ctypes.c_void_p(0))
    ms = a.elapsed_time(b)
    units = (nq + 127) // 128 * (nd / 256) / 148
    print(f"dense_topk {ms:.3f} ms ({2*nq*nd*dim/ms/1e9:.0f} TFLOP/s incl. select), tiles per SM {units:.0f}")
s = stats.cpu().numpy().astype(np.float64)
for i, n in enumerate(names):
    print(f"  {n:18s} mean {s[:, i].mean()/1e6:9.3f} Mcyc   per-unit {s[:, i].mean()/units:9.1f} cyc")
