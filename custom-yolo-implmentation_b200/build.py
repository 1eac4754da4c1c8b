"""Build ``csrc/*.cu`` into ``csrc/libyolo_boxpath.so`` for sm_100a with nvcc (in-tree, so the
library travels with the repo snapshot to the GPU box).  Cross-compiles without a GPU.

    python custom-yolo-implmentation_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libyolo_boxpath.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _stamp(sources) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sorted(sources) + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "yolo_boxpath.h")]:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    stamp_path = os.path.join(CSRC, "build", "stamp")
    stamp = _stamp(sources)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        return LIB
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(CSRC, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(CSRC, "build", os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, res.stderr))
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart",
           "-Xlinker", "--no-undefined"]          # a declared-but-deleted entry point fails here, not at dlopen
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stderr)
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
