"""Build libheston_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# (source, extra flags): the SABR unit must not contract a*b+c into FMAs (see its header)
SOURCES = ["heston_b200.cu", "sabr_b200.cu"]
EXTRA = {"sabr_b200.cu": ["-fmad=false"]}
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
    env = dict(os.environ)
    env.pop("CC", None)  # this image exports CC/CXX wrappers nvcc must not pick up
    env.pop("CXX", None)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    objs = []
    for src in SOURCES:
        obj = os.path.join(HERE, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *compile_flags, *EXTRA.get(src, []), *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, src]
        subprocess.run(cmd, cwd=HERE, check=True, env=env)
        objs.append(obj)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", TARGET, *objs], cwd=HERE,
                   check=True, env=env)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
