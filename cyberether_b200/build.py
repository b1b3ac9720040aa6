"""In-tree build of libb200dsp.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["context.cu", "elementwise.cu", "fft.cu", "fir.cu", "fm.cu", "agc.cu"]
OUT = os.path.join(HERE, "libb200dsp.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    newest = max(os.path.getmtime(os.path.join(root, f)) for root, _, files in os.walk(CSRC) for f in files)
    header = os.path.join(os.path.dirname(HERE), "include", "b200dsp.h")
    return max(newest, os.path.getmtime(header)) > os.path.getmtime(OUT)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
