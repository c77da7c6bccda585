"""Builds libsmx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m summarymixing_b200.build [--force] [-v]

Every csrc/*.cu becomes an object under build/ (compiled in parallel, rebuilt only when the source or a header is
newer), then one link step produces summarymixing_b200/libsmx.so.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libsmx.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + [
    "-lineinfo", "-O3", "-std=c++17", "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return (glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))
            + glob.glob(os.path.join(ROOT, "include", "*.h")))


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def _compile(src: str, verbose: bool) -> str:
    res = subprocess.run([NVCC] + CFLAGS + ["-c", "-o", _obj(src), src], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
    with open(_obj(src)[:-2] + ".ptxas.log", "w") as f:  # registers / spills / shared memory per kernel (-Xptxas -v)
        f.write(res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return _obj(src)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    todo = [s for s in sources()
            if force or not os.path.exists(_obj(s)) or os.path.getmtime(_obj(s)) < max(os.path.getmtime(s), hdr_t)]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [_obj(s) for s in sources()]
    res = subprocess.run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libsmx.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
