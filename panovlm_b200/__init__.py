"""panovlm_b200 — B200-native (sm_100a) correspondence-and-residual hot path of PanoVLM.

The product is the C-ABI shared library ``libpanovlm_b200.so`` (``include/panovlm_b200.h``); this package is a
thin ctypes mirror of that ABI for tests, the benchmark and Python callers.  There is NO CPU fallback: creating a
:class:`Context` without the built library or without a CUDA device raises.
"""
from .api import (Context, PvbError, LineFrame, BlockList, lib_path, load_library, build_library,  # noqa: F401
                  P2PLANE_METER, P2PLANE_ANGLE, P2LINE_METER, P2LINE_ANGLE, PLANE2PLANE_GLOBAL, PLANE_IOU)

__all__ = ["Context", "PvbError", "LineFrame", "BlockList", "lib_path", "load_library", "build_library",
           "P2PLANE_METER", "P2PLANE_ANGLE", "P2LINE_METER", "P2LINE_ANGLE", "PLANE2PLANE_GLOBAL", "PLANE_IOU"]
