"""tfqmrgpu_b200 - host-side Python mirror of the B200-native tfQMR library (libtfQMRgpu.so).

The product is the C-ABI shared library built from ``tfqmrgpu_b200/csrc`` (hand-written CUDA for
sm_100a behind the reference's ``tfqmrgpu.h`` interface).  This package only binds it with ctypes
(``api``), generates/reads problems (``problems``) and drives RHS-block-column sharding across the GPUs of
one box (``sharded``).  There is no CPU fallback: loading fails if the CUDA library was not built.
"""
from . import _lib  # noqa: F401
from .api import BsrsvPlan, Handle, TfqmrError, bsrsv  # noqa: F401

__all__ = ["BsrsvPlan", "Handle", "TfqmrError", "bsrsv"]
