"""Build the drop-in library fastglobetrotter_b200/fastGLOBETROTTERCompanion.so for sm_100a.

nvcc cross-compiles without a GPU.  The library is built in-tree (git-ignored, but shipped to
the GPU box by gpurun) and links cudart statically, so it can be dyn.load()ed by R or ctypes
without torch.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "fastGLOBETROTTERCompanion.so")
SOURCES = ["kernels.cu", "engine.cu", "companion_api.cpp", "dist.cpp"]
HEADERS = ["kernels.cuh", "engine.h", "glibc_rand.h", "dist.h", os.path.join("..", "..", "include", "fgt_companion.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # bit-exact fp64: the reference binary has no FMA contraction
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default,-pthread",
    "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.rsplit(".", 1)[0] + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd)))
        objs.append(o)
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    subprocess.check_call(["nvcc", "-shared", "-cudart", "static", "-o", OUT, *objs, "-lz", "-lpthread", "-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
