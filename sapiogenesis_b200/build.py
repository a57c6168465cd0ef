"""In-tree build of libsapio_b200.so (hand-written CUDA for sm_100a behind the C-ABI of include/sapio.h)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsapio_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    out = []
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".def")):
                out.append(os.path.join(d, f))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    stale = force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in sources())
    if stale:
        cmd = [NVCC] + FLAGS + ["-o", SO, os.path.join(CSRC, "sapio_abi.cu"), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed")
        if verbose:
            print(log)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
