"""tudocomp_b200 — B200-native (sm_100a) TextDS + lzss_lcp + BWT hot path of tudocomp behind a C ABI.

The package binds `tudocomp_b200/libtdcgpu.so` (built by `__graft_entry__.build()` / `make -C tudocomp_b200/csrc`).
There is NO CPU fallback: importing works without the library (so that CPU-only tooling can import `synth`), but any
use of the compute API raises `RuntimeError` when the CUDA library is missing and `TdcGpuError` when no GPU is present.
"""
from __future__ import annotations

import os

from . import synth  # noqa: F401
from ._abi import BWT, FACTOR_DTYPE, ISA, LCP, PHI, PLCP, SA, Context, TdcGpuError, TdcGpuLib  # noqa: F401

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtdcgpu.so")
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> TdcGpuLib:
    """Bind the CUDA library.  Raises loudly if it has not been built — nothing else can serve the hot path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  tudocomp_b200 has no CPU fallback."
            )
        _lib = TdcGpuLib(_LIB_PATH)
    return _lib


from .textds import FactorBuffer, LZSSLCPCompressor, TextDS, bwt  # noqa: E402,F401
