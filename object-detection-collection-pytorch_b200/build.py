"""Build libyolohead.so in-tree with nvcc for sm_100a.

    python -m odcp_b200.build            (from the repository root)

The library has no Python or torch dependency: plain C ABI (include/yolohead.h), CUDA runtime
linked statically.  It is built next to its sources so that it travels with the repository
snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libyolohead.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), out=None):
    """`out`: build an experimental variant (extra -D flags) next to the product library.
    The translation units are compiled in parallel (one nvcc -c each), then linked into the shared library."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    import concurrent.futures
    import tempfile
    nvcc = find_nvcc()
    target = out or LIB_PATH
    common = [*NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    log = []

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        log.append(res.stdout + res.stderr)

    with tempfile.TemporaryDirectory(prefix="yh_build_") as tmp:
        objs = [os.path.join(tmp, os.path.basename(src)[:-3] + ".o") for src in sources()]
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda so: run([nvcc, "-c", *common, "-o", so[1], so[0]]), zip(sources(), objs)))
        run([nvcc, "-shared", *NVCC_FLAGS, "-o", target, *objs])
    if verbose and any(log):
        print("".join(log))
    return target


if __name__ == "__main__":
    ex = ["-Xptxas", "-v"] if "-v" in sys.argv else []
    print(build(force=True, verbose=True, extra=ex))
