"""In-tree build of libsapio_b200.so (hand-written CUDA for sm_100a behind the C-ABI of include/sapio.h).

Every .cu under csrc/ is one translation unit (the size classes of the organism solver and the coupled kernels are compiled on
their own: the templates dominate the build time); objects are cached under csrc/_obj/ and rebuilt when the unit or a header it
includes (transitively) changed; units compile in parallel."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
_TAG = os.environ.get("SAPIO_BUILD_TAG", "")            # diagnostic / A-B builds: their own object cache and library name (SAPIO_SO selects one at run time)
OBJ = os.path.join(CSRC, "_obj" + ("_" + _TAG if _TAG else ""))
SO = os.path.join(HERE, "libsapio_b200" + ("_" + _TAG if _TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("SAPIO_NVCC_DEFS", "").split()      # e.g. -DORG_PHASE_PROF (diagnostic builds)
_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def units():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def deps(path, seen=None):
    """The unit and every header it includes through #include "..." (transitively)."""
    seen = seen if seen is not None else set()
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def sources():
    out = set()
    for u in units():
        out |= deps(u)
    return sorted(out)


def source_hash() -> str:
    """sha256 over the CUDA sources: keys measurements taken under a profiler (profiles/ncu_traffic.json) to the code they measured."""
    import hashlib
    h = hashlib.sha256()
    for path in sources():
        h.update(os.path.basename(path).encode()); h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


def _compile(unit, force):
    obj = os.path.join(OBJ, os.path.basename(unit)[:-3] + ".o")
    if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps(unit)):
        return obj, "", 0
    cmd = [NVCC] + FLAGS + ["-c", "-o", obj, unit]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, " ".join(cmd) + "\n" + r.stdout + r.stderr, r.returncode


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 16))) as pool:
        res = list(pool.map(lambda u: _compile(u, force), units()))
    log = "".join(r[1] for r in res)
    objs = [r[0] for r in res]
    failed = any(r[2] for r in res)
    relink = failed is False and (force or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs))
    if relink:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + r.stdout + r.stderr
        failed = failed or r.returncode != 0
    if log:
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(log)
    if failed:
        sys.stderr.write(log[-20000:])
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
