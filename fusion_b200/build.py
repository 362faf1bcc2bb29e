"""Build recipe for the CUDA library: ``python -m fusion_b200.build`` -> fusion_b200/libfusion_b200.so.

nvcc cross-compiles for sm_100a without a GPU.  The library is a plain C-ABI shared object (no torch, no
pybind) linked against the static CUDA runtime; it is built in-tree so that it travels with the repo snapshot
to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfusion_b200.so")
STAMP = os.path.join(HERE, ".libfusion_b200.stamp")
SOURCES = ["select.cu", "fuse.cu", "metrics.cu", "textbuild.cu", "activations.cu", "sparse.cu", "build.cu", "dense.cu", "splade.cu", "maxsim.cu"]
NVCC_FLAGS = ([] if not os.environ.get("FZ_KERNEL_STATS") else ["-DFZ_KERNEL_STATS"]) + (
    [] if not os.environ.get("FZ_PREFETCH_SCATTER") else ["-DFZ_PREFETCH_SCATTER"]) + [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-cudart", "static",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, file=sys.stderr)
        if p.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xcompiler", "-fPIC", *objs, "-o", LIB]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
