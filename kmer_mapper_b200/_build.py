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
SOURCES = [os.path.join(CSRC, "kmb_capi.cu"), os.path.join(CSRC, "kmb_reader.cpp"), os.path.join(CSRC, "kmb_hostpack.cpp"),
           os.path.join(CSRC, "kmb_gunzip.cpp"), os.path.join(CSRC, "kmb_inflate.cpp")]
HEADERS = [os.path.join(CSRC, "kmb_kernels.cuh"), os.path.join(CSRC, "kmb_core.cuh"), os.path.join(CSRC, "kmb_host.h"),
           os.path.join(CSRC, "kmb_textparse.cuh"), os.path.join(CSRC, "kmb_gzdev.cuh"),
           os.path.join(os.path.dirname(PKG), "include", "kmer_mapper_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]
LINK_FLAGS = ["-lz", "-ldl"]  # member-parallel gzip inflate of the chunk reader (kmb_gunzip.cpp)


def up_to_date() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS if os.path.exists(p))


BOUNDS_LIB_PATH = os.path.join(PKG, "libkmer_mapper_b200_bounds.so")  # debug build, see build_bounds_checked()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB_PATH)
    return nvcc


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and up_to_date():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + SOURCES + LINK_FLAGS + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


def build_bounds_checked(verbose: bool = True) -> str:
    """The same sources with -DKMB_BOUNDS_CHECKS: every device-side index is compared with the size of what it
    indexes (kmb_kernels.cuh, KMB_BOUND).  Not the product: run the GPU tests against it with
    ``KMB_LIB_PATH=kmer_mapper_b200/libkmer_mapper_b200_bounds.so python -m pytest tests -m gpu``; the session
    fails if any check fired (tests/conftest.py)."""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DKMB_BOUNDS_CHECKS"] + SOURCES + LINK_FLAGS + ["-o", BOUNDS_LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return BOUNDS_LIB_PATH


def build_variant(name: str, defines, verbose: bool = True) -> str:
    """An A/B build of the same sources with extra -D flags (tuning experiments: run with KMB_LIB_PATH=<result>)."""
    out = os.path.join(PKG, "libkmer_mapper_b200_%s.so" % name)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + SOURCES + LINK_FLAGS + ["-o", out]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    argv = sys.argv[1:]
    if "--bounds" in argv:
        print(build_bounds_checked())
    elif "--variant" in argv:
        i = argv.index("--variant")
        print(build_variant(argv[i + 1], argv[i + 2:]))
    else:
        print(build(force=True))
