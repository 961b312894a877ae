"""Build libhmfe.so (sm_100a) in-tree with nvcc.  No JIT, no torch extension machinery:
the library is a plain C-ABI shared object loaded with ctypes (``_lib.py``)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.path.join(HERE, "libhmfe.so")
HOST_CHECK = os.path.join(ROOT, "build", "host_check")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libhmfe.so")
    return exe


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu") and f != "host_check.cu"
    )


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "hmfe.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, log: list) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(src), _deps_mtime())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((src, r.stdout + r.stderr))
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a and link libhmfe.so."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    log: list = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, log), sources()))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for src, out in log:
            print(f"== {os.path.basename(src)}\n{out}")
    return LIB_PATH


def build_host_check(force: bool = False) -> str:
    """CPU emulation harness of the warp algorithms (test infrastructure, g++ only)."""
    src = os.path.join(CSRC, "host_check.cu")
    os.makedirs(os.path.dirname(HOST_CHECK), exist_ok=True)
    if not force and os.path.exists(HOST_CHECK) and os.path.getmtime(HOST_CHECK) >= max(os.path.getmtime(src), _deps_mtime()):
        return HOST_CHECK
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-x", "c++", f"-I{cuda_inc}", "-o", HOST_CHECK, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"host_check build failed:\n{r.stdout}\n{r.stderr}")
    return HOST_CHECK


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
