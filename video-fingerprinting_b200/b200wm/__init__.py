"""b200wm: Python side of the B200 watermark hot path (ctypes over libb200wm.so)."""
from . import _lib, ops            # importing fails loudly if the shared library is missing
from ._lib import B200wmError, LIB_PATH

__all__ = ["ops", "B200wmError", "LIB_PATH"]
