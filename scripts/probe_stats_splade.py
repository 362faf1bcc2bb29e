"""Cycle-counter probe of the SPLADE head GEMM (codes-mode filter_gemm_kernel); needs a FZ_KERNEL_STATS=1 build."""
import ctypes, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fusion_b200 import _lib, ops
from fusion_b200.index import SparseIndex, sparse_queries
lib = _lib.load()
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = 6980
dp, dt, dw = bench.make_splade(0, n, n, 120, 8, 512, 311, dev)
ix = SparseIndex(dp, dt, dw, bench.SPLADE_VOCAB, "cos_sim", device=dev, boot_docs=min(262144, n // 8 // 256 * 256))
qp, qt, qw = bench.make_splade(0, nq, nq, 24, 2, 64, 312, dev)
q_ptr, q_term, q_w = sparse_queries(qp, qt, qw, "cos_sim", dev)
stats = torch.zeros((148, 8), dtype=torch.int64, device=dev)
names = ["prod_wait_empty", "mma_wait_tempty", "mma_wait_full", "mma_issue", "epi0_wait_tfull", "epi0_total", "epi0_append", "epi0_append_n"]
for it in range(3):
    stats.zero_()
    lib.fz_debug_set_stats(ctypes.c_void_p(stats.data_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ix.topk(q_ptr, q_term, q_w, 1000); b.record(); torch.cuda.synchronize()
    lib.fz_debug_set_stats(ctypes.c_void_p(0))
units = (nq + 127) // 128 * ((n - 262144) / 256) / 148
print(f"splade_topk {a.elapsed_time(b):.2f} ms, GEMM tiles per SM {units:.0f}")
s = stats.cpu().numpy().astype(np.float64)
for i, nm in enumerate(names):
    print(f"  {nm:18s} mean {s[:, i].mean()/1e6:9.3f} Mcyc   per-tile {s[:, i].mean()/units:9.1f} cyc")
