"""Build the REAL reference lookup kernel as a checker (test infrastructure, not product code).

Compiles /root/reference/kmer_mapper/mapper.pyx -- unmodified, read where it lies, never copied
into this repository -- to ``oracle/_ref/kmer_mapper/mapper*.so`` with the reference's own optimisation level
(setup.py:9-15: -O3 -march=native; here -march=x86-64-v3 instead of native, because the .so is
built in the CPU container and executed on a different host, the GPU box).  Only build outputs (generated C, the .so and an
empty package marker) are written, all under the git-ignored ``oracle/_ref/``; the .so travels to
the GPU box with the gpurun snapshot, the reference tree does not.

Also compiles the plain-C restatement ``oracle/kmer_oracle.c`` to ``oracle/_build/libkmer_oracle.so``.

Usage:  python oracle/build_ref.py            (idempotent; skips what is up to date)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/kmer_mapper/mapper.pyx"
REF_OUT = os.path.join(HERE, "_ref")
C_OUT = os.path.join(HERE, "_build")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def ref_so_path() -> str:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(REF_OUT, "kmer_mapper", "mapper" + suffix)


def build_reference_mapper(verbose: bool = True) -> str | None:
    """Returns the path of the compiled reference module, or None when /root/reference is absent
    and no prebuilt copy exists (e.g. on the GPU box before a snapshot carrying it)."""
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None
    if _newer(so, REF_PYX):
        return so
    import numpy as np
    pkg = os.path.join(REF_OUT, "kmer_mapper")
    os.makedirs(pkg, exist_ok=True)
    with open(os.path.join(pkg, "__init__.py"), "w") as f:  # the reference's __init__.py is empty too
        f.write("")
    c_file = os.path.join(pkg, "mapper.c")
    cmd = [sys.executable, "-m", "cython", "-3", "--module-name", "kmer_mapper.mapper", REF_PYX, "-o", c_file]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    cc = ["gcc", "-shared", "-fPIC", "-O3", "-march=x86-64-v3", "-w",
          "-I", sysconfig.get_paths()["include"], "-I", np.get_include(),
          "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", c_file, "-o", so, "-lm"]
    if verbose:
        print(" ".join(cc))
    subprocess.check_call(cc)
    return so


def build_c_oracle(verbose: bool = True) -> str:
    src = os.path.join(HERE, "kmer_oracle.c")
    so = os.path.join(C_OUT, "libkmer_oracle.so")
    if _newer(so, src):
        return so
    os.makedirs(C_OUT, exist_ok=True)
    cc = ["gcc", "-shared", "-fPIC", "-O3", "-march=x86-64-v3", "-fopenmp", "-Wall", src, "-o", so]
    if verbose:
        print(" ".join(cc))
    subprocess.check_call(cc)
    return so


if __name__ == "__main__":
    print("reference mapper:", build_reference_mapper())
    print("C oracle:", build_c_oracle())
