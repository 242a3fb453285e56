"""Builds the CUDA engine in-tree: geneticscre_b200/libgcre_b200.so (sm_100a only, -lineinfo)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libgcre_b200.so")
SOURCES = [os.path.join(PKG, "csrc", "gcre_capi.cu"), os.path.join(PKG, "csrc", "host_pack.cpp")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC,-O3,-Wall,-pthread", "-shared", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include")]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(ROOT, "include", "gcre_b200.h")]
    csrc = os.path.join(PKG, "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libgcre_b200.so")
    with open(os.path.join(PKG, "csrc", ".ptxas.log"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
