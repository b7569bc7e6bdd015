"""Build recipe of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU).

``python -m kmer_mapper_b200._build`` or ``__graft_entry__.build()`` writes
``kmer_mapper_b200/libkmer_mapper_b200.so`` in-tree (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.environ.get("KMB_LIB_PATH") or os.path.join(PKG, "libkmer_mapper_b200.so")  # override: A/B builds only
SOURCES = [os.path.join(CSRC, "kmb_capi.cu"), os.path.join(CSRC, "kmb_reader.cpp")]
HEADERS = [os.path.join(CSRC, "kmb_kernels.cuh"), os.path.join(CSRC, "kmb_core.cuh"),
           os.path.join(os.path.dirname(PKG), "include", "kmer_mapper_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def up_to_date() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS if os.path.exists(p))


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and up_to_date():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB_PATH)
    cmd = [nvcc] + NVCC_FLAGS + SOURCES + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True))
