"""ctypes loader for libnmfb200.so (the C-ABI product library).

There is deliberately no fallback: if the CUDA library is missing or cannot be
loaded the import of anything that computes fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NMFB_LIB: path of an alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("NMFB_LIB") or os.path.join(_HERE, "libnmfb200.so")

_lib = None


class NmfbLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libnmfb200.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NmfbLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
    return _lib
