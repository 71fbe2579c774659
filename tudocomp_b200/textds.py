"""Python mirror of the reference interface for the hot path (used by tests and bench; the production host side is
the C++ plugin in tudocomp_b200/plugin/ compiled against the reference headers).

Names and meaning follow the reference:
  TextDS            include/tudocomp/ds/TextDS.hpp            require_sa/isa/lcp/phi/plcp, size(), text
  LZSSLCPCompressor include/tudocomp/compressors/LZSSLCPCompressor.hpp   option `threshold` (default 3), factorize phase
  FactorBuffer      include/tudocomp/compressors/lzss/LZSSFactors.hpp    shortest_factor()/longest_factor()/is_sorted()
  bwt               include/tudocomp/ds/bwt.hpp:19-22 as used by BWTCompressor::compress
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi


def _ctx(device: int) -> _abi.Context:
    from . import load

    return _abi.Context(load(), device)


class TextDS:
    SA, ISA, LCP, PHI, PLCP = _abi.SA, _abi.ISA, _abi.LCP, _abi.PHI, _abi.PLCP

    def __init__(self, text, flags: int = 0, device: int = 0):
        text = np.ascontiguousarray(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text,
                                    dtype=np.uint8)
        if text.size == 0 or text[-1] != 0:
            # same condition and wording as TextDS::TextDS (ds/TextDS.hpp:132-138)
            raise ValueError("Input has no sentinel! Please make sure you declare the compressor calling this with "
                             "`m.needs_sentinel_terminator()` in its `meta()` function.")
        self._text = text
        self.ctx = _ctx(device)
        self.ctx.set_text(text)
        if flags:
            self.require(flags)

    def require(self, flags: int) -> None:
        self.ctx.build(flags)

    def _req(self, flag: int) -> np.ndarray:
        self.ctx.build(flag)
        return self.ctx.get(flag)

    def require_sa(self) -> np.ndarray:
        return self._req(_abi.SA)

    def require_isa(self) -> np.ndarray:
        return self._req(_abi.ISA)

    def require_lcp(self) -> np.ndarray:
        return self._req(_abi.LCP)

    def require_phi(self) -> np.ndarray:
        return self._req(_abi.PHI)

    def require_plcp(self) -> np.ndarray:
        return self._req(_abi.PLCP)

    def max_lcp(self) -> int:
        self.ctx.build(_abi.PLCP)
        return self.ctx.max_lcp()

    def size(self) -> int:
        return int(self._text.size)

    @property
    def text(self) -> np.ndarray:
        return self._text

    def close(self) -> None:
        self.ctx.close()


class FactorBuffer:
    def __init__(self, factors: np.ndarray, shortest: int, longest: int):
        self.factors = factors
        self._shortest, self._longest = shortest, longest

    def __len__(self) -> int:
        return int(self.factors.size)

    def size(self) -> int:
        return len(self)

    def empty(self) -> bool:
        return len(self) == 0

    def is_sorted(self) -> bool:
        return bool(np.all(np.diff(self.factors["pos"].astype(np.int64)) >= 0))

    def shortest_factor(self) -> int:
        return self._shortest

    def longest_factor(self) -> int:
        return self._longest

    def as_triples(self) -> np.ndarray:
        f = self.factors
        return np.stack([f["pos"], f["src"], f["len"]], axis=1) if len(self) else np.zeros((0, 3), np.uint32)


class LZSSLCPCompressor:
    """`lzss_lcp(threshold=3)`: the TextDS + Factorize phases on the GPU (the Encode phase stays with the coder)."""

    def __init__(self, threshold: int = 3, device: int = 0):
        if threshold < 1:
            raise ValueError("threshold must be >= 1")
        self.threshold, self.device = threshold, device

    def factorize(self, text, textds: Optional[TextDS] = None) -> FactorBuffer:
        own = textds is None
        t = textds or TextDS(text, device=self.device)
        try:
            z, mn, mx = t.ctx.factorize(self.threshold)
            return FactorBuffer(t.ctx.factors(z), mn, mx)
        finally:
            if own:
                t.close()


def bwt(text, device: int = 0) -> np.ndarray:
    """BWTCompressor::compress payload: BWT[i] = SA[i] ? T[SA[i]-1] : T[n-1] (n bytes, one 0)."""
    t = TextDS(text, device=device)
    try:
        t.ctx.build(_abi.SA | _abi.BWT)
        return t.ctx.get(_abi.BWT)
    finally:
        t.close()
