"""Build libhalva_dpa.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhalva_dpa.so")
SOURCES = [os.path.join(CSRC, "dpa_api.cu"), os.path.join(CSRC, "dpa_host.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("dpa_ptx.cuh", "dpa_gemm.cuh", "dpa_kernels.cuh", "dpa_peer.cuh")] + \
    [os.path.join(os.path.dirname(HERE), "include", "halva_dpa.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "--use_fast_math" if False else "-DDPA_BUILD=1"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS if os.path.exists(p))


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    """`out` / `defines` produce experiment variants next to the default library (picked up with HALVA_DPA_LIB=...)."""
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else LIB, defines=defs))
