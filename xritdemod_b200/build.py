"""Builds the in-tree native libraries of xritdemod_b200.

  libxrd.so     -- CUDA kernels + C ABI (include/xrd.h), sm_100a only
  libxrdsig.so  -- host-side synthetic IQ source (csrc/siggen.c)

`python -m xritdemod_b200.build` or __graft_entry__.build().  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBXRD = os.path.join(HERE, "libxrd.so")
LIBSIG = os.path.join(HERE, "libxrdsig.so")
LIBSHIM = os.path.join(HERE, "libxrd_shim_test.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # every fused multiply-add in the kernels is an explicit fmaf()
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared",
]


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found")
    return n


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_libxrd(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "xrd_api.cu")] + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "xrd.h")]
    if force or _stale(LIBXRD, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIBXRD, srcs[0]]
        subprocess.check_call(cmd)
    return LIBXRD


def build_siggen(force=False):
    src = os.path.join(CSRC, "siggen.c")
    if force or _stale(LIBSIG, [src]):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", LIBSIG, src, "-lm"])
    return LIBSIG


def build_shim_test(force=False):
    """C++ host-side mirror of the reference operator interface (include/xrd_sathelper.hpp),
    compiled into a small test driver library."""
    src = os.path.join(CSRC, "shim_test.cpp")
    hdr = os.path.join(ROOT, "include", "xrd_sathelper.hpp")
    if not os.path.exists(src):
        return None
    if force or _stale(LIBSHIM, [src, hdr, LIBXRD]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               "-o", LIBSHIM, src, "-L", HERE, "-lxrd", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return LIBSHIM


def build_all(force=False, verbose=False):
    build_libxrd(force, verbose)
    build_siggen(force)
    build_shim_test(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIBXRD, LIBSIG)
