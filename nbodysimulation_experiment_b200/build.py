"""Builds libsphb200.so (hand-written CUDA for sm_100a + the C ABI of include/sphb200.h) in-tree.

nvcc cross-compiles without a GPU, so this also runs in the CPU-only authoring container.  The
.so stays inside the package directory (git-ignored, but it travels to the GPU box).
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsphb200.so")
SOURCES = [os.path.join(CSRC, "sphb200.cu")]
HEADERS = [
    os.path.join(CSRC, "sph_math.cuh"),
    os.path.join(CSRC, "sph_kernels.cuh"),
    os.path.join(os.path.dirname(PKG_DIR), "include", "sphb200.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
    "-ldl",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsphb200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile if missing or older than its sources; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
