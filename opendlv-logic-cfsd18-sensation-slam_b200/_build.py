"""In-tree build of the CUDA library (sm_100a only).

`build()` compiles csrc/*.cu|*.cpp with nvcc into libslam_b200.so next to this file.  The .so is
git-ignored but travels to the GPU box with the working tree.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libslam_b200.so")
HOSTLIB = os.path.join(HERE, "libslam_b200_host.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]
# per-file extra flags: the association kernels must round products and sums separately, like the
# reference's x86-64 -O2 build (no FMA contraction), to keep gate decisions bit-exact.
EXTRA = {"assoc.cu": ["-fmad=false"]}
SOURCES = ["capi.cu", "graph.cu", "solver.cu", "assoc.cu", "symbolic.cpp", "tileplan.cpp"]
HOST_SOURCES = ["host/cone.cpp", "host/slam.cpp", "host/frame_assembler.cpp", "host/wgs84.cpp", "host/rec_reader.cpp", "host/slam_c.cpp"]


def _nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return p


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))]
    hd = os.path.join(CSRC, "host")
    if os.path.isdir(hd):
        hs += [os.path.join(hd, f) for f in os.listdir(hd) if f.endswith((".h", ".hpp"))]
    hs.append(os.path.join(HERE, "..", "include", "slam_b200.h"))
    return hs


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    # host-side mirror of the reference's Slam/Cone classes over the C ABI
    hobjs = []
    have_host = all(os.path.exists(os.path.join(CSRC, s)) for s in HOST_SOURCES)
    if have_host:
        for src in HOST_SOURCES:
            s = os.path.join(CSRC, src)
            o = os.path.join(OBJ, src.replace("/", "_") + ".o")
            hobjs.append(o)
            if force or _stale(o, [s] + hdrs):
                cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-Wall", "-Wextra", "-c", s, "-o", o]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.run(cmd, check=True)
        if force or _stale(HOSTLIB, hobjs + [LIB]):
            cmd = ["g++", "-shared", "-o", HOSTLIB] + hobjs + ["-L" + HERE, "-lslam_b200", "-Wl,-rpath,$ORIGIN"]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
