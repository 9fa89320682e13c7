"""Builds cama_b200/_lib/libcama_b200.so (sm_100a only) with nvcc, in-tree.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  `python -m cama_b200.build`
rebuilds unconditionally; `ensure_built()` rebuilds when a source is newer than the library.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libcama_b200.so")
STAMP_PATH = LIB_PATH + ".stamp"
SOURCES = ["ops.cu", "clip.cu", "overlay.cu", "densify.cu", "remap.cu", "lidar.cu", "peer.cu"]
HEADERS = ["common.cuh", "geom.cuh", "host_pool.h", os.path.join("..", "..", "include", "cama_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # every FMA in the arithmetic contract is written explicitly (csrc/geom.cuh)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",      # host loops run on the library's own worker pool (csrc/host_pool.h)
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libcama_b200.so cannot be built")
    return exe


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def source_digest():
    """sha256 over flags + every source/header: the staleness test (mtimes do not survive the
    snapshot copy to the GPU box)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sources() + [os.path.normpath(os.path.join(CSRC, x)) for x in HEADERS]:
        if os.path.exists(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as fh:
        return fh.read().strip() != source_digest()


def build(verbose=False, extra_flags=(), out=None):
    """out=<path>: build a variant there (with extra -D flags) and leave the product library alone."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if out is not None:
        res = subprocess.run([_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", out, *sources()], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return out
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, *sources()]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)
    with open(STAMP_PATH, "w") as fh:
        fh.write(source_digest())
    return LIB_PATH


def ensure_built():
    """Build if missing or stale and a compiler is available; never silently fall back.
    CAMA_B200_LIB=<path> loads that build of the library instead (A/B experiments with compile-time variants)."""
    override = os.environ.get("CAMA_B200_LIB")
    if override:
        if not os.path.exists(override):
            raise RuntimeError(f"CAMA_B200_LIB={override} does not exist")
        return override
    if is_stale():
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing and nvcc is not available to build it")
        else:
            raise RuntimeError(f"{LIB_PATH} was built from other sources than the ones in {CSRC} (digest mismatch) and nvcc is "
                               "not available to rebuild it; set CAMA_B200_LIB=<path> to load a specific build on purpose")
    return LIB_PATH


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "-v" in sys.argv else []
    print(build(verbose=True, extra_flags=flags))
