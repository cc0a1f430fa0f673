"""Put pde_b200's module object where the reference's UNMODIFIED Python layer looks for its pybind11 extension
(INTEGRATION.md section 1), for the reference-side boundary tests.

The reference package comes from baseline/_ref (baseline/install_ref.py: `pip install --target`; git-ignored,
travels to the GPU box).  Its root ``quant_trading/__init__.py`` imports the whole trading platform (database,
monitoring, ... -- SQLAlchemy and friends are absent in this image), so the root package is
registered as a bare namespace; every sub-package below it (``quant_trading.cpp``, ``.models``,
``.calibration``) is the reference's own file, executed unmodified.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SITE = os.path.join(ROOT, "baseline", "_ref")
REF_TESTS = os.path.join(REF_SITE, "_reftests")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SITE, "quant_trading", "calibration"))


def inject():
    """-> the injected module object (pde_b200.cpp.quant_cpp)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import pde_b200.cpp.quant_cpp as b200

    pkg = types.ModuleType("quant_trading")
    pkg.__path__ = [os.path.join(REF_SITE, "quant_trading")]
    sys.modules["quant_trading"] = pkg
    # what `from . import quant_cpp` in the reference's cpp/__init__.py (:21-23) resolves to
    sys.modules["quant_trading.cpp.quant_cpp"] = b200
    import quant_trading.cpp as qcpp  # the reference's own file

    assert qcpp.is_available() and qcpp.quant_cpp is b200, qcpp.get_import_error()
    return b200


def main(argv):
    """python tests/ref_inject.py <pytest args>: run pytest with the module injected (used as a subprocess so the
    reference package never leaks into this repository's own test session)."""
    import pytest

    inject()
    return pytest.main(list(argv))


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
