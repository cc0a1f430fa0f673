"""Compile tests/ext/heston_b200_bindings.cpp (the pybind11 binding of INTEGRATION.md section 2) against
include/heston_b200.h and link it to the in-tree libheston_b200.so.  Test scaffolding: the product never loads it.

    python tests/ext/build_ext.py     -> tests/ext/quant_cpp_b200<ext-suffix>  (git-ignored; travels to the GPU box)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "heston_b200_bindings.cpp")
TARGET = os.path.join(HERE, "quant_cpp_b200" + sysconfig.get_config_var("EXT_SUFFIX"))
LIBDIR = os.path.join(ROOT, "pde_b200", "csrc")


def build(force: bool = False) -> str:
    lib = os.path.join(LIBDIR, "libheston_b200.so")
    if not force and os.path.exists(TARGET) and os.path.getmtime(TARGET) >= max(os.path.getmtime(SRC),
                                                                                  os.path.getmtime(lib)):
        return TARGET
    import pybind11

    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{pybind11.get_include()}",
           f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(ROOT, 'include')}", SRC, "-o", TARGET,
           f"-L{LIBDIR}", "-lheston_b200", "-Wl,-rpath,$ORIGIN/../../pde_b200/csrc"]
    subprocess.run(cmd, check=True, env=env)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
