"""In-tree build of libdcg_b200.so (sm_100a only).

    python -m deep_cartograph_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  The shared library is written to
``deep_cartograph_b200/lib/`` so it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdcg_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


OBJ_DIR = os.path.join(PKG_DIR, "build")


def _compile_one(src: str, verbose: bool):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", INCLUDE, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return obj, cmd, res


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (one nvcc process per file, in parallel) and link
    them into one shared library; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, verbose), srcs))
    objs = []
    for obj, cmd, res in results:
        if verbose:
            sys.stderr.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        objs.append(obj)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-o", LIB_PATH + ".tmp"] + objs + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
