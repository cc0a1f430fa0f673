"""Build libheston_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["heston_b200.cu"]
HEADERS = ["kernels.cuh", "fft_smem.cuh", "heston_math.cuh", os.path.join("..", "..", "include", "heston_b200.h")]
TARGET = os.path.join(HERE, "libheston_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886,177",
]


def needs_build() -> bool:
    if not os.path.exists(TARGET):
        return True
    t = os.path.getmtime(TARGET)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return TARGET
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", TARGET, *SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)  # this image exports CC/CXX wrappers nvcc must not pick up
    env.pop("CXX", None)
    subprocess.run(cmd, cwd=HERE, check=True, env=env)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
