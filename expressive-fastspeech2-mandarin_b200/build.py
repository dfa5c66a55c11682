"""Compile csrc/*.cu into libfs2b200.so (in-tree, sm_100a only) with nvcc."""
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libfs2b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  [os.path.join(os.path.dirname(PKG_DIR), "include", "fs2_b200.h")])


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > built for s in _sources())


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libfs2b200.so.  Cross-compiles without a GPU."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libfs2b200.so cannot be built (and there is no CPU fallback)")
    trace = ["-DFS2_TRACE_BUILD"] if os.environ.get("FS2_TRACE_BUILD") else []   # tools/trace_*.py instrumentation
    cmd = [nvcc] + NVCC_FLAGS + trace + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH, os.path.join(CSRC, "fs2_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
