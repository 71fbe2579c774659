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
    "tdcgpu_dist_text_device_ptr", "tdcgpu_dist_factors_device_ptr",
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
    L.tdcgpu_dist_text_device_ptr.argtypes = [C.c_void_p]
    L.tdcgpu_dist_text_device_ptr.restype = C.c_void_p
    L.tdcgpu_dist_factors_device_ptr.argtypes = [C.c_void_p]
    L.tdcgpu_dist_factors_device_ptr.restype = C.c_void_p
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

    def get_into_device(self, which: int, dev_ptr: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_dist_get(self._h, which, C.c_void_p(dev_ptr), 1))

    def verify_full(self, threshold: int, zl: int, device: int, dist=None) -> dict:
        """Collective.  EVERY slot and EVERY parse position is checked on the device (tdcgpu_check_index /
        tdcgpu_check_factors, csrc/check.cu): the SA / LCP / ISA shards are gathered into full arrays on every rank
        (torch.distributed all_gather, plumbing), then each rank checks its own slot range and its own position range
        against them.  Returns the violation counters of this rank plus `ok` (all ranks, all zero)."""
        import torch

        from ._abi import Context

        dev = torch.device("cuda", device)
        world = dist.get_world_size() if dist is not None else 1
        info, n = self.shard_info(), self.n

        # all ranks take the same decision: the three gathered arrays (12 n bytes) must fit next to the shards
        free = torch.tensor([torch.cuda.mem_get_info(dev)[0]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(free, op=dist.ReduceOp.MIN)
        if int(free.item()) < 12 * n + (2 << 30):
            return {"ok": None, "skipped": f"full verification needs {12 * n >> 30} GiB for the gathered arrays, {int(free.item()) >> 30} GiB free"}

        def gather(which, cnt):
            cnts = torch.zeros(world, dtype=torch.int64, device=dev)
            cnts[self.rank] = cnt
            if world > 1:
                dist.all_reduce(cnts)
            cnts = [int(x) for x in cnts.tolist()]
            full = torch.empty(sum(cnts), dtype=torch.int32, device=dev)
            off = 0
            for r in range(world):  # every shard goes straight into its place: the owner writes it, then broadcasts the slice
                part = full[off:off + cnts[r]]
                if r == self.rank and cnt:
                    self.get_into_device(which, part.data_ptr())
                if world > 1 and cnts[r]:
                    dist.broadcast(part, src=r)
                off += cnts[r]
            return full

        sa = gather(SA, info["slot_cnt"])
        lcp = gather(LCP, info["slot_cnt"])
        isa = gather(ISA, info["pos_cnt"])
        torch.cuda.synchronize()  # the checkers run on their own stream
        assert sa.numel() == n and lcp.numel() == n and isa.numel() == n, "shards do not tile the arrays"
        L = self.lib.lib
        t_ptr = L.tdcgpu_dist_text_device_ptr(self._h)
        f_ptr = L.tdcgpu_dist_factors_device_ptr(self._h) or 0
        # a factor that starts in an earlier rank's range may cover the first positions of mine
        ends = torch.zeros(world, dtype=torch.int64, device=dev)
        if zl:
            fl = torch.empty(3 * zl, dtype=torch.int32, device=dev)
            self.lib.check(L.tdcgpu_dist_get_factors(self._h, C.c_void_p(fl.data_ptr()), zl, 1))
            last = [int(x) & 0xFFFFFFFF for x in fl[-3:].tolist()]
            ends[self.rank] = last[0] + last[2]
            del fl
        if world > 1:
            dist.all_reduce(ends)
        covered_to = max([0] + [int(x) for x in ends.tolist()[:self.rank]])
        pos_lo = max(info["pos_lo"], min(covered_to, info["pos_lo"] + info["pos_cnt"]))
        pos_cnt = info["pos_lo"] + info["pos_cnt"] - pos_lo
        with Context(self.lib, device) as chk:  # only its stream and scalar scratch are used
            res = chk.check_index_ptrs(t_ptr, n, sa.data_ptr(), isa.data_ptr(), lcp.data_ptr(), info["slot_lo"], info["slot_cnt"])
            res.update(chk.check_factors_ptrs(t_ptr, n, sa.data_ptr(), isa.data_ptr(), lcp.data_ptr(), f_ptr, zl, threshold, pos_lo, pos_cnt))
        bad = torch.tensor([sum(res.values())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(bad)
        res["ok"] = int(bad.item()) == 0
        res["checked_slots"], res["checked_positions"] = info["slot_cnt"], pos_cnt
        del sa, lcp, isa
        torch.cuda.empty_cache()
        return res

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
