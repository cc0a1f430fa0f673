"""Counterpart of the reference's ``quant_trading/cpp/__init__.py``: exposes ``quant_cpp``."""
from . import quant_cpp  # noqa: F401

__all__ = ["quant_cpp"]


def is_available() -> bool:
    """True when the CUDA library can be loaded (it is built on demand)."""
    try:
        from .. import _lib

        _lib.load()
        return True
    except Exception:
        return False


def get_import_error() -> str:
    try:
        from .. import _lib

        _lib.load()
        return ""
    except Exception as e:  # pragma: no cover
        return str(e)
