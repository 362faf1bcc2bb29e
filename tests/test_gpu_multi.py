"""Multi-GPU parity over REAL NCCL: the corpus-sharded hybrid pipeline must return, for every rank's query slice,
what one unsharded index returns - BM25 bit-exact, the float systems to 1e-5.  Needs >= 2 visible GPUs (skipped on the
1-GPU tier); spawns one process per GPU under torchrun and keeps the full log (gpurun_out/parity_g<N>.log)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_sharded_pipeline_equals_single_index_over_nccl(n_gpus):
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "multi_gpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    log = r.stdout + "\n--- stderr ---\n" + r.stderr
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_g{n_gpus}.log"), "w") as fh:
            fh.write(log)
    assert r.returncode == 0 and "MULTI_GPU_PARITY PASS" in r.stdout, log[-4000:]
