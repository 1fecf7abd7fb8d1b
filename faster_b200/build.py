"""Builds faster_b200/lib/libfaster_b200.so (CUDA kernels + C ABI + host helpers) for sm_100a with nvcc.

In-tree build: the .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfaster_b200.so")
SOURCES = ["fq_kernels.cu", "fq_capi.cu", "fq_host.cpp", "fq_decomp.cpp", "fq_jps.cpp"]
PUBLIC_HEADER = os.path.join(HERE, "..", "include", "faster_b200.h")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-cudart", "static"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [PUBLIC_HEADER, os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in deps)      # every file of csrc/ (sources and all .cuh/.h)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags/out: build an experimental variant next to the product library (tuning runs only)."""
    if out is None and not (force or _stale()):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-o", out or LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
