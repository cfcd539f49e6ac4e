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
HOST_LIB = os.path.join(PKG, "librast_host.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",  # the reference's fp32 ops are never fused
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas"]  # (#pragma unroll reaches the host pass through the __host__ __device__ functions)


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


def build_variant(name, defines, verbose=False):
    """build/variants/librast_b200_<name>.so: the same library with extra -D switches, selected at run time with
    RAST_LIB=<path> (A/B measurements of kernel variants in one GPU session)."""
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "librast_b200_%s.so" % name)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-shared", "-o", out, os.path.join(PKG, "csrc", "rast_ctx.cu")]
    subprocess.check_call(cmd)
    return out


HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-fPIC"]


def build_renderer(force=False):
    """rasteriser_b200/renderer: the reference-compatible command line (host/main.cpp) on the C ABI."""
    host = os.path.join(PKG, "host")
    deps = _sources(host, os.path.join(ROOT, "include")) + [LIB]
    if force or _stale(RENDERER, deps):
        srcs = [os.path.join(host, f) for f in ("main.cpp", "args.cpp", "loaders.cpp", "png.cpp")]
        subprocess.check_call(["g++"] + HOST_FLAGS + ["-o", RENDERER] + srcs + ["-L", PKG, "-lrast_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return RENDERER


def build_host_lib(force=False):
    """rasteriser_b200/librast_host.so: loaders, PNG codec and flag parser behind a C API (tests, bench)."""
    host = os.path.join(PKG, "host")
    if force or _stale(HOST_LIB, _sources(host)):
        srcs = [os.path.join(host, f) for f in ("host_capi.cpp", "args.cpp", "loaders.cpp", "png.cpp")]
        subprocess.check_call(["g++"] + HOST_FLAGS + ["-shared", "-o", HOST_LIB] + srcs + ["-lz"])
    return HOST_LIB


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_host_lib(force)
    build_renderer(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
