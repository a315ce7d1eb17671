"""Build the CUDA C-ABI library (pcramp_b200/csrc/libpcramp_gpu.so).

Everything is built in-tree with explicit nvcc / make invocations so that the resulting .so files
travel with a snapshot of the repository (they are git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pcramp_b200", "csrc")
LIB = os.path.join(CSRC, "libpcramp_gpu.so")
HOST_DIR = os.path.join(ROOT, "pcramp_b200", "host")
HOST_BIN = os.path.join(HOST_DIR, "pcramp_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # identity / coverage arithmetic must round like the reference's scalar x86 code: no FMA contraction,
    # IEEE division and square root (the defaults, stated here so nobody adds --use_fast_math)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared", "-lpthread",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "pcramp_gpu.h"))
    if not force and not _stale(LIB, srcs):
        return LIB
    units = [s for s in srcs if s.endswith(".cu")]
    extra = os.environ.get("PCRAMP_NVCC_EXTRA", "").split()   # tuning experiments only (e.g. -DTHERMO_MIN_BLOCKS=6)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["--threads", str(len(units))] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + units
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX wrapper is not a usable host compiler for nvcc
    env.pop("CC", None)
    subprocess.run(cmd, check=True, env=env)
    return LIB


def build_host(force=False):
    """the command-line host (pcramp_b200/host/pcramp_b200): plain g++, linked against the C-ABI library next to it"""
    src = os.path.join(HOST_DIR, "pcramp_b200_main.cpp")
    if not force and not _stale(HOST_BIN, [src, os.path.join(ROOT, "include", "pcramp_gpu.h"), LIB]):
        return HOST_BIN
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, src, "-L" + CSRC, "-lpcramp_gpu", "-lz",
                    "-Wl,-rpath,$ORIGIN/../csrc"], check=True)
    return HOST_BIN


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
    print("built", build_host(force="--force" in sys.argv))
