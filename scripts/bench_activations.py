"""HBM roofline of the SPLADE activation head (csrc/activations.cu) on one GPU: CUDA-event timing, inputs >> L2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_b200 import activations as A  # noqa: E402

peak = 6484.6
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
out = {}
for dtype in (torch.float32, torch.bfloat16):
    for b, l, v in [(32, 256, 32005), (128, 64, 32005)]:
        logits = (torch.randn((b, l, v), device=dev, generator=g) * 2).to(dtype)
        lens = torch.randint(l // 2, l + 1, (b,), device=dev, generator=g)
        mask = (torch.arange(l, device=dev)[None, :] < lens[:, None]).int()
        nbytes = int(mask.sum()) * v * logits.element_size() + b * v * 4
        for pooling in ("max", "sum"):
            for _ in range(3):
                A.splade_pool(logits, mask, pooling)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                A.splade_pool(logits, mask, pooling)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[f"pool_{pooling}_{str(dtype)[6:]}_{b}x{l}x{v}"] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1),
                                                                   "frac_of_measured_hbm": round(nbytes / ms / 1e6 / peak, 3)}
        del logits
act = torch.relu(torch.randn((65536, 32005), device=dev, generator=g) - 2.4)      # ~0.8 % non-zeros, 8.4 GB
for name, fn, nb in [("csr", lambda: A.activations_to_csr(act), 2 * act.numel() * 4),
                     ("prune_topk_128", lambda: A.prune_activations(act, 128, want_indices=False), 6 * act.numel() * 4)]:
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[name + "_65536x32005"] = {"ms": round(ms, 3), "GB/s": round(nb / ms / 1e6, 1), "frac_of_measured_hbm": round(nb / ms / 1e6 / peak, 3)}
print(json.dumps(out, indent=1))
