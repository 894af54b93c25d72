"""Build recipe for the CUDA library (in-tree, sm_100a only).

    python -m lego_loam_sr_b200.build        (or: python lego-loam-sr_b200/build.py)

Outputs lego-loam-sr_b200/lib/liblegoloam_b200.so (the C ABI of include/legoloam_b200.h) and
lego-loam-sr_b200/lib/liblegoloam_host.so (the C++ mirror of the reference's MapOptimization member
functions, host/). -fmad=false is part of the numerics contract (csrc/common.cuh).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "lib")
ABI_LIB = os.path.join(LIB_DIR, "liblegoloam_b200.so")
HOST_LIB = os.path.join(LIB_DIR, "liblegoloam_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(srcs, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_abi(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [os.path.join(ROOT, "include", "legoloam_b200.h")]
    if not (force or _newer(srcs, ABI_LIB)):
        return ABI_LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cus = [s for s in srcs if s.endswith(".cu")]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", ABI_LIB] + cus
    subprocess.check_call(cmd, stdout=sys.stderr)  # (bench.py prints exactly one line on stdout)
    return ABI_LIB


def build_host(force=False):
    hdir = os.path.join(HERE, "host")
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir))] if os.path.isdir(hdir) else []
    cpps = [s for s in srcs if s.endswith(".cpp")]
    if not cpps:
        return None
    if not (force or _newer(srcs + [ABI_LIB], HOST_LIB)):
        return HOST_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", HOST_LIB] + cpps + \
          ["-L", LIB_DIR, "-llegoloam_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
    subprocess.check_call(cmd, stdout=sys.stderr)  # (bench.py prints exactly one line on stdout)
    return HOST_LIB


def build_all(force=False, verbose=False):
    a = build_abi(force, verbose)
    b = build_host(force)
    return a, b


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
