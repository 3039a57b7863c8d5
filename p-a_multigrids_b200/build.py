"""Builds libpamg_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpamg_cuda.so")
HOST_BIN = os.path.join(LIBDIR, "pamg_host")

SOURCES = ["pamg_api.cu", "pamg_mesh.cpp", "pamg_plan.cpp"]
DEPS = SOURCES + ["pamg_kernels.cuh", "pamg_unstr.cuh", "pamg_internal.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-cudart", "static",
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(ROOT, "include", "pamg.h"), __file__]
    if force or stale(LIB, deps):
        cmd = [nvcc(), "-ccbin", host_cxx()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libpamg_cuda.so")
    host_src = os.path.join(HERE, "host", "pamg_host.cpp")
    if os.path.exists(host_src) and (force or stale(HOST_BIN, [host_src, LIB])):
        cmd = [host_cxx(), "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), host_src, "-o", HOST_BIN,
               "-L", LIBDIR, "-lpamg_cuda", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building pamg_host")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
