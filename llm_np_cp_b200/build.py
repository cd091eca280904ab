"""Builds libb200fwd.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200fwd.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "gemm_tcgen05.cuh", "gemm2_tcgen05.cuh", "gemm_sk_tcgen05.cuh", "attention_tcgen05.cuh", "elementwise.cuh",
           os.path.join("..", "..", "include", "b200fwd.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # B2_DEFINES="-DB2_DIAG -DATT_POLY=1" etc.: bring-up / A-B builds (tools/gpu_diag.py tracing needs -DB2_DIAG)
    defines = os.environ.get("B2_DEFINES", "").split()
    out = os.environ.get("B2_LIB_OUT", LIB)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-Xptxas", "-v" if verbose else "-warn-spills"] + defines + [
           "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libb200fwd.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
