"""ctypes binding of the multi-GPU entry points (tdcgpu_dist_*, include/tdcgpu.h): one text sharded over the ranks of a
torch.distributed job (one process per GPU).  torch.distributed is only the plumbing that hands rank 0's NCCL id to the
other ranks; the data path is the library's own NCCL all-to-alls (tudocomp_b200/csrc/dist_textds.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import FACTOR_DTYPE, ISA, LCP, SA, TdcGpuLib

DIST_EXPORTS = [
    "tdcgpu_dist_unique_id", "tdcgpu_dist_create", "tdcgpu_dist_destroy", "tdcgpu_dist_set_text", "tdcgpu_dist_build",
    "tdcgpu_dist_shard_info", "tdcgpu_dist_get", "tdcgpu_dist_max_lcp", "tdcgpu_dist_lzss_lcp_factorize",
    "tdcgpu_dist_get_factors", "tdcgpu_dist_sync", "tdcgpu_dist_event_record", "tdcgpu_dist_event_elapsed_ms",
    "tdcgpu_dist_stats", "tdcgpu_dist_phase_count", "tdcgpu_dist_phase_name", "tdcgpu_dist_phase_ms",
]


def bind_dist(lib: TdcGpuLib) -> None:
    L = lib.lib
    if getattr(L, "_tdc_dist_bound", False):
        return
    L.tdcgpu_dist_destroy.argtypes = [C.c_void_p]
    L.tdcgpu_dist_destroy.restype = None
    L.tdcgpu_dist_set_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.tdcgpu_dist_build.argtypes = [C.c_void_p, C.c_uint32]
    L.tdcgpu_dist_shard_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.tdcgpu_dist_get.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    L.tdcgpu_dist_max_lcp.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.tdcgpu_dist_lzss_lcp_factorize.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.tdcgpu_dist_get_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.tdcgpu_dist_sync.argtypes = [C.c_void_p]
    L.tdcgpu_dist_event_record.argtypes = [C.c_void_p, C.c_int]
    L.tdcgpu_dist_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.tdcgpu_dist_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.tdcgpu_dist_phase_count.argtypes = [C.c_void_p]
    L.tdcgpu_dist_phase_name.argtypes = [C.c_void_p, C.c_int]
    L.tdcgpu_dist_phase_name.restype = C.c_char_p
    L.tdcgpu_dist_phase_ms.argtypes = [C.c_void_p, C.c_int]
    L.tdcgpu_dist_phase_ms.restype = C.c_float
    L._tdc_dist_bound = True


class DistContext:
    """One rank of a sharded text index.  `handle` is a tdcgpu_dist* created by `create_nccl` (product) or by the CPU
    simulator harness in tests/ (test infrastructure)."""

    def __init__(self, lib: TdcGpuLib, handle: C.c_void_p, rank: int, nranks: int):
        bind_dist(lib)
        self.lib, self._h, self.rank, self.nranks, self.n = lib, handle, rank, nranks, 0

    @classmethod
    def create_nccl(cls, lib: TdcGpuLib, device: int, dist=None) -> "DistContext":
        """Collective.  dist: an initialised torch.distributed module (None = single rank)."""
        bind_dist(lib)
        L = lib.lib
        L.tdcgpu_dist_unique_id.argtypes = [C.c_void_p]
        L.tdcgpu_dist_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        ident = np.zeros(128, np.uint8)
        if world > 1:
            import torch

            if rank == 0:
                lib.check(L.tdcgpu_dist_unique_id(C.c_void_p(ident.ctypes.data)))
            t = torch.from_numpy(ident).cuda(device)
            dist.broadcast(t, 0)
            ident = t.cpu().numpy()
        h = C.c_void_p()
        lib.check(L.tdcgpu_dist_create(device, rank, world, C.c_void_p(ident.ctypes.data), C.byref(h)))
        return cls(lib, h, rank, world)

    def close(self) -> None:
        if self._h:
            self.lib.lib.tdcgpu_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_text(self, text: np.ndarray) -> None:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        self.lib.check(self.lib.lib.tdcgpu_dist_set_text(self._h, C.c_void_p(text.ctypes.data), text.size, 0))
        self.n = int(text.size)

    def set_text_ptr(self, ptr: int, n: int, on_device: bool) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_set_text(self._h, C.c_void_p(ptr), n, 1 if on_device else 0))
        self.n = int(n)

    def build(self, flags: int = SA | ISA | LCP) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_build(self._h, flags))

    def shard_info(self) -> dict:
        buf = (C.c_uint64 * 4)()
        self.lib.check(self.lib.lib.tdcgpu_dist_shard_info(self._h, buf))
        return dict(zip(["slot_lo", "slot_cnt", "pos_lo", "pos_cnt"], [int(x) for x in buf]))

    def get(self, which: int) -> np.ndarray:
        info = self.shard_info()
        out = np.empty(info["pos_cnt"] if which == ISA else info["slot_cnt"], dtype=np.uint32)
        self.lib.check(self.lib.lib.tdcgpu_dist_get(self._h, which, C.c_void_p(out.ctypes.data), 0))
        return out

    def max_lcp(self) -> int:
        v = C.c_uint32()
        self.lib.check(self.lib.lib.tdcgpu_dist_max_lcp(self._h, C.byref(v)))
        return int(v.value)

    def factorize(self, threshold: int = 3):
        """-> (local factor count, total factor count, min len, max len)"""
        zl, zt, mn, mx = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint32()
        self.lib.check(self.lib.lib.tdcgpu_dist_lzss_lcp_factorize(self._h, threshold, C.byref(zl), C.byref(zt), C.byref(mn), C.byref(mx)))
        return int(zl.value), int(zt.value), int(mn.value), int(mx.value)

    def factors(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=FACTOR_DTYPE)
        self.lib.check(self.lib.lib.tdcgpu_dist_get_factors(self._h, C.c_void_p(out.ctypes.data), count, 0))
        return out

    def get_factors_into(self, host_ptr: int, cap: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_get_factors(self._h, C.c_void_p(host_ptr), cap, 0))

    def sync(self) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_sync(self._h))

    def event_record(self, slot: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self.lib.check(self.lib.lib.tdcgpu_dist_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def stats(self) -> dict:
        buf = (C.c_uint64 * 8)()
        self.lib.check(self.lib.lib.tdcgpu_dist_stats(self._h, buf))
        keys = ["rounds", "active_sum", "radix_passes", "radix_elems", "alphabet", "symbols_per_key", "capacity", "p2p"]
        return dict(zip(keys, [int(x) for x in buf]))

    def phases(self):
        L = self.lib.lib
        return [(L.tdcgpu_dist_phase_name(self._h, i).decode(), float(L.tdcgpu_dist_phase_ms(self._h, i)))
                for i in range(L.tdcgpu_dist_phase_count(self._h))]
