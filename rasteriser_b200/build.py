"""Build the in-tree native libraries.

  rasteriser_b200/librast_b200.so   CUDA kernels + C ABI (include/rast.h), sm_100a only
  rasteriser_b200/renderer          reference-compatible C++ command line (host/), links the ABI

nvcc cross-compiles without a GPU.  The .so files are git-ignored but travel to the GPU box with
the gpurun snapshot."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "librast_b200.so")
RENDERER = os.path.join(PKG, "renderer")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",  # the reference's fp32 ops are never fused
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs):
    out = []
    for d in dirs:
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
                out.append(os.path.join(d, f))
    return out


def build_lib(force=False, verbose=False):
    csrc = os.path.join(PKG, "csrc")
    deps = _sources(csrc, os.path.join(ROOT, "include"))
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB, os.path.join(csrc, "rast_ctx.cu")]
        subprocess.check_call(cmd)
    return LIB


def build_renderer(force=False):
    host = os.path.join(PKG, "host")
    main = os.path.join(host, "main.cpp")
    if not os.path.exists(main):
        return None
    deps = _sources(host, os.path.join(ROOT, "include")) + [LIB]
    if force or _stale(RENDERER, deps):
        srcs = [os.path.join(host, f) for f in sorted(os.listdir(host)) if f.endswith(".cpp")]
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", RENDERER] + srcs + \
              ["-L", PKG, "-lrast_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"]
        subprocess.check_call(cmd)
    return RENDERER


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_renderer(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
