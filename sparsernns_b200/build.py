"""Build libs5b200.so in-tree with nvcc for sm_100a (B200).

`python -m sparsernns_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The library is linked against the static CUDA runtime so that it does not depend on which
libcudart the host process (PyTorch, jaxlib) has loaded.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libs5b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libs5b200 needs the CUDA 12.9 toolkit to build")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _extra_flags():
    return os.environ.get("S5B200_NVCC_FLAGS", "").split()   # e.g. -DS5_TC_TRACE (phase timeline printf, debugging)


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(_extra_flags()).encode())
    for f in sorted(os.listdir(CSRC)) + ["../../include/s5b200.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    srcs = _sources()

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *ARCH, *FLAGS, *_extra_flags(), "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, cmd, r

    # one nvcc process per translation unit, side by side (the largest file bounds the build time)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, srcs))
    objs = []
    logs = []
    for src, obj, cmd, r in results:
        logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
