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


def _up_to_date(stamp: str, digest: str) -> bool:
    try:
        with open(stamp) as f:
            return os.path.exists(LIB) and f.read().strip() == digest
    except OSError:
        return False


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build (or reuse) the library.  Safe to call from several processes at once (one rank per GPU under torchrun):
    the build runs under an exclusive file lock, objects go to a per-process directory, and the .so and its stamp are
    moved into place with os.replace, so nobody can dlopen a half-written file."""
    import fcntl

    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()
    if not force and _up_to_date(stamp, digest):
        return LIB
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(stamp, digest):      # another process built it while we waited
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, digest: str, verbose: bool) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, f"obj.{os.getpid()}")
    os.makedirs(objdir, exist_ok=True)
    srcs = _sources()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
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
    tmp_lib = os.path.join(objdir, "libs5b200.so")
    cmd = [nvcc, *ARCH, "-shared", "-o", tmp_lib, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(logs))
    if os.path.exists(stamp):
        os.remove(stamp)                                       # never a new .so beside an old stamp or vice versa
    os.replace(tmp_lib, LIB)
    tmp_stamp = os.path.join(objdir, "build.stamp")
    with open(tmp_stamp, "w") as f:
        f.write(digest)
    os.replace(tmp_stamp, stamp)
    shutil.rmtree(objdir, ignore_errors=True)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
