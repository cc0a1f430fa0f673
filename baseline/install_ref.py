"""Install the UNMODIFIED reference Python package where the GPU box can see it: baseline/_ref/ (git-ignored, so
no reference source enters the history; not gpurun-ignored, so it travels with the snapshot).

    python baseline/install_ref.py          (build container only: needs /root/reference)

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>` --
   the reference's setup.py never builds its C++ (setup.py:90-93), so this is its pure-Python layer:
   quant_trading.models, quant_trading.calibration, ... without quant_trading/cpp/quant_cpp*.so.
2. Its own acceptance tests for the binding boundary are copied next to it, unmodified
   (tests/python/test_cpp_bindings.py, tests/python/calibration/test_calibration.py, tests/python/conftest.py)
   -> baseline/_ref/_reftests/.

tests/test_gpu_reference_side.py then injects pde_b200.cpp.quant_cpp as quant_trading.cpp.quant_cpp and runs the
reference's wrapper, calibrator and those tests over it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")
DEST = os.path.join(HERE, "_ref")
TESTS = ["tests/python/test_cpp_bindings.py", "tests/python/calibration/test_calibration.py", "tests/python/conftest.py"]


def installed() -> bool:
    return os.path.isdir(os.path.join(DEST, "quant_trading", "calibration")) and os.path.isdir(
        os.path.join(DEST, "_reftests"))


def install(force: bool = False) -> bool:
    """-> True when baseline/_ref holds the package (installed now or before)."""
    if installed() and not force:
        return True
    if not os.path.isdir(REF):
        return False
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")  # the build writes egg-info into the source tree; /root/reference is read-only
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git"))
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--upgrade", "--target", DEST, src], check=True,
                       stdout=subprocess.DEVNULL)
    out = os.path.join(DEST, "_reftests")
    os.makedirs(out, exist_ok=True)
    for rel in TESTS:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(out, os.path.basename(rel)))
    return True


if __name__ == "__main__":
    print("installed" if install(force="--force" in sys.argv) else "reference not present")
