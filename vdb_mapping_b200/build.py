"""Builds libvdbm_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python -m vdb_mapping_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with gpurun.
-fmad=false: the fp64 DDA must replay the reference's rounding sequence (no FMA contraction).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libvdbm_b200.so")
SOURCES = ["vdbm_kernels.cu", "vdbm_abi.cu", "vdbm_group.cu"]
HEADERS = [os.path.join(CSRC, "vdbm_device.cuh"), os.path.join(ROOT, "include", "vdbm_b200.h")]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or _stale(SO, objs):
        cmd = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
