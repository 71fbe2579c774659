"""ctypes binding of the C ABI declared in include/tdcgpu.h.

`TdcGpuLib(path)` binds one shared object.  The package (`tudocomp_b200/__init__.py`) only ever binds
`tudocomp_b200/libtdcgpu.so` — the nvcc-built CUDA library — and raises if it is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

SA, ISA, LCP, PHI, PLCP, BWT = 0x01, 0x02, 0x04, 0x08, 0x10, 0x100

EXPORTS = [
    "tdcgpu_last_error", "tdcgpu_device_count", "tdcgpu_create", "tdcgpu_destroy", "tdcgpu_set_text",
    "tdcgpu_textds_build", "tdcgpu_textds_get", "tdcgpu_textds_device_ptr", "tdcgpu_textds_max_lcp",
    "tdcgpu_lzss_lcp_factorize", "tdcgpu_lzss_lcp_get_factors", "tdcgpu_textds_build_host", "tdcgpu_bwt_host",
    "tdcgpu_phase_count", "tdcgpu_phase_name", "tdcgpu_phase_ms", "tdcgpu_sa_stats", "tdcgpu_sync",
    "tdcgpu_event_record", "tdcgpu_event_elapsed_ms", "tdcgpu_launch_count", "tdcgpu_profile_enable",
    "tdcgpu_profile_reset", "tdcgpu_profile_count", "tdcgpu_profile_entry",
    "tdcgpu_lzss_literal_histogram", "tdcgpu_lzss_encode", "tdcgpu_lzss_encode_get", "tdcgpu_textds_get_packed",
    "tdcgpu_mtf_encode", "tdcgpu_rle_encode", "tdcgpu_literal_encode_begin", "tdcgpu_literal_encode", "tdcgpu_literal_encode_get",
    "tdcgpu_lzss_encode_get_chunk", "tdcgpu_literal_encode_get_chunk", "tdcgpu_pinned_alloc", "tdcgpu_pinned_free", "tdcgpu_set_device", "tdcgpu_set_text_cached", "tdcgpu_set_len_bits", "tdcgpu_device_alloc", "tdcgpu_device_free", "tdcgpu_device_copy",
    "tdcgpu_sa_layout", "tdcgpu_check_index", "tdcgpu_check_factors", "tdcgpu_text_device_ptr", "tdcgpu_factors_device_ptr",
]

FACTOR_DTYPE = np.dtype([("pos", "<u4"), ("src", "<u4"), ("len", "<u4")])


class TdcGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tdcgpu error {code}: {msg}")
        self.code = code


class TdcGpuLib:
    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.tdcgpu_last_error.restype = C.c_char_p
        L.tdcgpu_device_count.restype = C.c_int
        L.tdcgpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.tdcgpu_destroy.argtypes = [C.c_void_p]
        L.tdcgpu_destroy.restype = None
        L.tdcgpu_set_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.tdcgpu_textds_build.argtypes = [C.c_void_p, C.c_uint32]
        L.tdcgpu_textds_get.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.tdcgpu_textds_get_packed.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int]
        L.tdcgpu_mtf_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.tdcgpu_rle_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        L.tdcgpu_literal_encode_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.tdcgpu_literal_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.POINTER(C.c_uint64)]
        L.tdcgpu_literal_encode_get.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.c_int]
        L.tdcgpu_textds_device_ptr.argtypes = [C.c_void_p, C.c_uint32]
        L.tdcgpu_textds_device_ptr.restype = C.c_void_p
        L.tdcgpu_textds_max_lcp.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.tdcgpu_lzss_lcp_factorize.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                                C.POINTER(C.c_uint32)]
        L.tdcgpu_lzss_lcp_get_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.tdcgpu_lzss_literal_histogram.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.tdcgpu_lzss_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.POINTER(C.c_uint64)]
        L.tdcgpu_lzss_encode_get.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.c_int]
        L.tdcgpu_textds_build_host.argtypes = [C.c_int, C.c_void_p, C.c_uint64] + [C.c_void_p] * 5 + [C.POINTER(C.c_uint32)]
        L.tdcgpu_bwt_host.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.tdcgpu_lzss_encode_get_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.tdcgpu_literal_encode_get_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.tdcgpu_pinned_alloc.argtypes = [C.c_uint64]
        L.tdcgpu_pinned_alloc.restype = C.c_void_p
        L.tdcgpu_pinned_free.argtypes = [C.c_void_p]
        L.tdcgpu_pinned_free.restype = None
        L.tdcgpu_check_index.argtypes = [C.c_void_p] + [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.tdcgpu_check_factors.argtypes = [C.c_void_p] + [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                         C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.tdcgpu_text_device_ptr.argtypes = [C.c_void_p]
        L.tdcgpu_text_device_ptr.restype = C.c_void_p
        L.tdcgpu_factors_device_ptr.argtypes = [C.c_void_p]
        L.tdcgpu_factors_device_ptr.restype = C.c_void_p
        L.tdcgpu_phase_count.argtypes = [C.c_void_p]
        L.tdcgpu_phase_name.argtypes = [C.c_void_p, C.c_int]
        L.tdcgpu_phase_name.restype = C.c_char_p
        L.tdcgpu_phase_ms.argtypes = [C.c_void_p, C.c_int]
        L.tdcgpu_phase_ms.restype = C.c_float
        L.tdcgpu_sa_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.tdcgpu_sa_layout.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.tdcgpu_sync.argtypes = [C.c_void_p]
        L.tdcgpu_event_record.argtypes = [C.c_void_p, C.c_int]
        L.tdcgpu_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.tdcgpu_launch_count.restype = C.c_uint64
        L.tdcgpu_profile_enable.argtypes = [C.c_int]
        L.tdcgpu_profile_enable.restype = None
        L.tdcgpu_profile_reset.restype = None
        L.tdcgpu_profile_entry.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double)]

    def check(self, rc: int) -> None:
        if rc < 0:
            raise TdcGpuError(rc, (self.lib.tdcgpu_last_error() or b"").decode(errors="replace"))

    def device_count(self) -> int:
        return int(self.lib.tdcgpu_device_count())

    def launch_count(self) -> int:
        return int(self.lib.tdcgpu_launch_count())

    def profile_enable(self, on: bool) -> None:
        self.lib.tdcgpu_profile_enable(1 if on else 0)

    def profile_reset(self) -> None:
        self.lib.tdcgpu_profile_reset()

    def profile(self) -> dict:
        """kernel name -> dict(launches, ms, bytes) aggregated since the last reset (resolves pending events)."""
        out = {}
        for i in range(self.lib.tdcgpu_profile_count()):
            name, n, ms, by = C.c_char_p(), C.c_uint64(), C.c_double(), C.c_double()
            self.check(self.lib.tdcgpu_profile_entry(i, C.byref(name), C.byref(n), C.byref(ms), C.byref(by)))
            out[name.value.decode()] = dict(launches=int(n.value), ms=float(ms.value), bytes=float(by.value))
        return out


def _host_ptr(a: np.ndarray) -> C.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class Context:
    """One device context: resident text, index arrays and factor list (tdcgpu_ctx)."""

    def __init__(self, lib: TdcGpuLib, device: int = 0):
        self.lib = lib
        self._h = C.c_void_p()
        lib.check(lib.lib.tdcgpu_create(device, C.byref(self._h)))
        self.n = 0

    def close(self) -> None:
        if self._h:
            self.lib.lib.tdcgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- text ------------------------------------------------------------------------------------------------------
    def set_text(self, text: np.ndarray) -> None:
        text = np.ascontiguousarray(text, dtype=np.uint8)
        self.lib.check(self.lib.lib.tdcgpu_set_text(self._h, _host_ptr(text), text.size, 0))
        self.n = int(text.size)

    def set_text_cached(self, text: np.ndarray) -> bool:
        """tdcgpu_set_text_cached: True if the resident text had exactly these bytes (built structures stay valid)."""
        text = np.ascontiguousarray(text, np.uint8)
        reused = C.c_int(0)
        self.lib.check(self.lib.lib.tdcgpu_set_text_cached(self._h, _host_ptr(text), C.c_uint64(text.size), C.byref(reused)))
        self.n = int(text.size)
        return bool(reused.value)

    def set_text_device(self, dev_ptr: int, n: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_set_text(self._h, C.c_void_p(dev_ptr), n, 1))
        self.n = int(n)

    # -- text DS ---------------------------------------------------------------------------------------------------
    def build(self, flags: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_textds_build(self._h, flags))

    def get(self, which: int) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint8 if which == BWT else np.uint32)
        self.lib.check(self.lib.lib.tdcgpu_textds_get(self._h, which, _host_ptr(out), 0))
        return out

    def get_packed(self, which: int, width: int) -> np.ndarray:
        """The array bit-packed to `width` bits per element (DynamicIntVector layout): uint64 words."""
        out = np.empty((self.n * width + 63) // 64, dtype=np.uint64)
        self.lib.check(self.lib.lib.tdcgpu_textds_get_packed(self._h, which, width, _host_ptr(out), out.size, 0))
        return out

    def get_into_device(self, which: int, dev_ptr: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_textds_get(self._h, which, C.c_void_p(dev_ptr), 1))

    def device_ptr(self, which: int) -> Optional[int]:
        return self.lib.lib.tdcgpu_textds_device_ptr(self._h, which)

    def max_lcp(self) -> int:
        v = C.c_uint32()
        self.lib.check(self.lib.lib.tdcgpu_textds_max_lcp(self._h, C.byref(v)))
        return int(v.value)

    # -- lzss_lcp --------------------------------------------------------------------------------------------------
    def factorize(self, threshold: int = 3):
        cnt, mn, mx = C.c_uint64(), C.c_uint32(), C.c_uint32()
        self.lib.check(self.lib.lib.tdcgpu_lzss_lcp_factorize(self._h, threshold, C.byref(cnt), C.byref(mn), C.byref(mx)))
        return int(cnt.value), int(mn.value), int(mx.value)

    def factors(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=FACTOR_DTYPE)
        self.lib.check(self.lib.lib.tdcgpu_lzss_lcp_get_factors(self._h, _host_ptr(out), count, 0))
        return out

    # -- lzss::encode_text on the device ---------------------------------------------------------------------------
    def literal_histogram(self):
        """(hist[256] of the literals outside factors, fdist_max) of the last factorize call."""
        hist = np.zeros(256, np.uint64)
        fd = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_lzss_literal_histogram(self._h, _host_ptr(hist), C.byref(fd)))
        return hist, int(fd.value)

    def encode(self, codes: np.ndarray, lens: np.ndarray, lead_bits: int = 0, lead_byte: int = 0) -> int:
        """Encode the factor list + literals as lzss::encode_text does; returns the stream length in bits."""
        codes = np.ascontiguousarray(codes, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint8)
        assert codes.size == 256 and lens.size == 256
        nbits = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_lzss_encode(self._h, _host_ptr(codes), _host_ptr(lens), lead_bits, lead_byte,
                                                       C.byref(nbits)))
        return int(nbits.value)

    def set_len_bits(self, len_field_bits: int) -> None:
        """Width of the archive's text-length field: 32 (default build of the reference) or 64 (its -DLEN_BITS=40 build)."""
        self.lib.check(self.lib.lib.tdcgpu_set_len_bits(self._h, C.c_uint32(len_field_bits)))

    def encoded(self, nbits: int, finalize: bool = True) -> np.ndarray:
        out = np.empty(nbits // 8 + 2, np.uint8)
        nb = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_lzss_encode_get(self._h, _host_ptr(out), out.size, 1 if finalize else 0,
                                                           C.byref(nb), 0))
        return out[:int(nb.value)]

    def encoded_chunks(self, chunk: int = 1 << 20, finalize: bool = True) -> np.ndarray:
        """The stream drained piece by piece (tdcgpu_lzss_encode_get_chunk), concatenated."""
        buf = np.empty(chunk, np.uint8)
        parts, off = [], 0
        while True:
            total, wr = C.c_uint64(), C.c_uint64()
            self.lib.check(self.lib.lib.tdcgpu_lzss_encode_get_chunk(self._h, off, _host_ptr(buf), chunk, 1 if finalize else 0, C.byref(total), C.byref(wr)))
            if wr.value == 0:
                break
            parts.append(buf[:wr.value].copy())
            off += wr.value
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)

    def encoded_into(self, host_ptr: int, cap: int, finalize: bool = True) -> int:
        nb = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_lzss_encode_get(self._h, C.c_void_p(host_ptr), cap, 1 if finalize else 0,
                                                           C.byref(nb), 0))
        return int(nb.value)

    # -- byte-stream stages behind the BWT -------------------------------------------------------------------------
    def mtf_encode(self, data: np.ndarray) -> np.ndarray:
        data = np.ascontiguousarray(data, np.uint8)
        out = np.empty(data.size, np.uint8)
        self.lib.check(self.lib.lib.tdcgpu_mtf_encode(self._h, _host_ptr(data), data.size, _host_ptr(out), 0))
        return out

    def rle_encode(self, data: np.ndarray, offset: int = 0) -> np.ndarray:
        data = np.ascontiguousarray(data, np.uint8)
        out = np.empty(12 * data.size + 32, np.uint8)
        m = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_rle_encode(self._h, _host_ptr(data), data.size, offset, _host_ptr(out), out.size, C.byref(m), 0))
        return out[:int(m.value)].copy()

    def literal_histogram_of(self, data: np.ndarray) -> np.ndarray:
        """LiteralEncoder step 1: stage `data` on the device, return its byte histogram."""
        data = np.ascontiguousarray(data, np.uint8)
        hist = np.zeros(256, np.uint64)
        self.lib.check(self.lib.lib.tdcgpu_literal_encode_begin(self._h, _host_ptr(data), data.size, 0, _host_ptr(hist)))
        return hist

    def literal_encode(self, codes: np.ndarray, lens: np.ndarray, lead_bits: int = 0, lead_byte: int = 0, finalize: bool = True) -> np.ndarray:
        """LiteralEncoder steps 2+3 on the staged data: the bit stream (with BitOStream's tail when finalize)."""
        codes = np.ascontiguousarray(codes, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint8)
        nbits = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_literal_encode(self._h, _host_ptr(codes), _host_ptr(lens), lead_bits, lead_byte, C.byref(nbits)))
        out = np.empty(int(nbits.value) // 8 + 2, np.uint8)
        nb = C.c_uint64()
        self.lib.check(self.lib.lib.tdcgpu_literal_encode_get(self._h, _host_ptr(out), out.size, 1 if finalize else 0, C.byref(nb), 0))
        return out[:int(nb.value)].copy()

    # -- device-side checkers (tdcgpu_check_*; not part of the product path) ----------------------------------------
    def check_index_ptrs(self, d_text: int, n: int, d_sa: int, d_isa: int, d_lcp: int, slot_lo: int, slot_cnt: int) -> dict:
        out = (C.c_uint64 * 4)()
        self.lib.check(self.lib.lib.tdcgpu_check_index(self._h, C.c_void_p(d_text), n, C.c_void_p(d_sa), C.c_void_p(d_isa),
                                                       C.c_void_p(d_lcp) if d_lcp else None, slot_lo, slot_cnt, out))
        return dict(zip(("isa_of_sa", "suffix_order", "lcp"), (int(x) for x in out[:3])))

    def check_factors_ptrs(self, d_text: int, n: int, d_sa: int, d_isa: int, d_lcp: int, d_factors: int, z: int, threshold: int,
                           pos_lo: int, pos_cnt: int) -> dict:
        out = (C.c_uint64 * 5)()
        self.lib.check(self.lib.lib.tdcgpu_check_factors(self._h, C.c_void_p(d_text), n, C.c_void_p(d_sa), C.c_void_p(d_isa), C.c_void_p(d_lcp),
                                                         C.c_void_p(d_factors) if d_factors else None, z, threshold, pos_lo, pos_cnt, out))
        return dict(zip(("malformed", "order", "factor_rule", "missed_factor", "undecided"), (int(x) for x in out)))

    def check(self, threshold: int = 0, z: int = 0) -> dict:
        """Full device-side check of the context's own SA / ISA / LCP (and, with threshold > 0, of its factor list)."""
        L = self.lib.lib
        t = L.tdcgpu_text_device_ptr(self._h)
        sa, isa, lcp = (self.device_ptr(w) for w in (SA, ISA, LCP))
        res = self.check_index_ptrs(t, self.n, sa, isa, lcp, 0, self.n)
        if threshold:
            res.update(self.check_factors_ptrs(t, self.n, sa, isa, lcp, L.tdcgpu_factors_device_ptr(self._h) or 0, z, threshold, 0, self.n))
        res["ok"] = all(v == 0 for v in res.values())
        return res

    # -- stats -----------------------------------------------------------------------------------------------------
    def phases(self):
        L = self.lib.lib
        return [(L.tdcgpu_phase_name(self._h, i).decode(), float(L.tdcgpu_phase_ms(self._h, i)))
                for i in range(L.tdcgpu_phase_count(self._h))]

    def sa_stats(self) -> dict:
        buf = (C.c_uint64 * 8)()
        self.lib.check(self.lib.lib.tdcgpu_sa_stats(self._h, buf))
        keys = ["rounds", "active_sum", "radix_passes", "radix_elems", "alphabet", "symbols_per_key", "lcp_route", "prefix_work"]
        out = dict(zip(keys, [int(x) for x in buf]))
        lay = (C.c_uint64 * 4)()
        self.lib.check(self.lib.lib.tdcgpu_sa_layout(self._h, lay))
        out.update(packed_records=int(lay[0]), key_bits=int(lay[1]), record_index_bits=int(lay[2]))
        return out

    def sync(self) -> None:
        self.lib.check(self.lib.lib.tdcgpu_sync(self._h))

    def event_record(self, slot: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self.lib.check(self.lib.lib.tdcgpu_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def get_factors_into(self, host_ptr: int, cap: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_lzss_lcp_get_factors(self._h, C.c_void_p(host_ptr), cap, 0))

    def set_text_host_ptr(self, host_ptr: int, n: int) -> None:
        self.lib.check(self.lib.lib.tdcgpu_set_text(self._h, C.c_void_p(host_ptr), n, 0))
        self.n = int(n)
