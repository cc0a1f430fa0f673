"""Build libheston_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# (source, extra flags): the SABR unit must not contract a*b+c into FMAs (see its header)
SOURCES = ["heston_b200.cu", "sabr_b200.cu"]
EXTRA = {"sabr_b200.cu": ["-fmad=false"]}
# every file a translation unit includes: a stale library after a math-header edit would silently keep the old bits
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".inc", ".h"))) + [
    os.path.join("..", "..", "include", "heston_b200.h")]
TARGET = os.path.join(HERE, "libheston_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886,177",
]


STAMP = TARGET + ".sha256"  # content hash of what the library was built from (mtimes do not survive a snapshot)


def source_hash() -> str:
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS + [f"{k}:{v}" for k, v in sorted(EXTRA.items())]).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not (os.path.exists(TARGET) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


# every elision of the fused kernel off: no perturbed-class tail skip, no asymptotic / series stage B, no rigorous
# tail bound in the decimation path (run it with hb_plan_set_truncation(plan, 0) for "exact" underflow-only zeros)
NOELIDE_DEFINES = ["HB_TAIL=0", "HB_ASYM_DT=1e300", "HB_MID=0", "HB_BOUND_DECIM=0", "HB_ZERO_AWARE=0", "HB_F3_PERT=0"]


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A second build of the same sources with -D switches (kernel A/B experiments and bench.py's no-elision
    leg), as libheston_b200_<name>.so; select it with PDE_B200_LIB."""
    target = os.path.join(HERE, f"libheston_b200_{name}.so")
    stamp = target + ".sha256"
    want = source_hash() + " " + " ".join(defines)
    if os.path.exists(target) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return target
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]
    objs = []
    for src in SOURCES:
        obj = os.path.join(HERE, f"{os.path.splitext(src)[0]}_{name}.o")
        cmd = [nvcc, *compile_flags, *EXTRA.get(src, []), *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, src]
        subprocess.run(cmd, cwd=HERE, check=True, env=env)
        objs.append(obj)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target, *objs], cwd=HERE,
                   check=True, env=env)
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return target


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
    with open(STAMP, "w") as fh:
        fh.write(source_hash() + "\n")
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
