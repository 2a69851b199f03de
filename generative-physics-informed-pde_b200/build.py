"""Builds libgpde_b200.so (the C-ABI extension, include/gpde_b200.h) in-tree with nvcc for sm_100a.

    python generative-physics-informed-pde_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so this also is the CPU-side "does it build" check (__graft_entry__.build()).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpde_b200.so")
SOURCES = ["rom.cu", "vo.cu", "prolong.cu", "fom_cg.cu"]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
          [os.path.join(os.path.dirname(HERE), "include", "gpde_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build_library(force=False, verbose=False):
    """Compiles the translation units in parallel (one nvcc -c each), then links the shared library."""
    if not force and not is_stale():
        return LIB
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + compile_flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, proc in jobs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode, " ".join(cmd)))
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed (%d): %s" % (res.returncode, " ".join(cmd)))
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
