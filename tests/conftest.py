import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on the B200 box)")
    config.addinivalue_line("markers", "sim: runs the kernels in the CPU interpreter tests/sim/cusim.h (test infrastructure)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a box without one (plain `pytest tests/` in the build container) they are
    skipped instead of failing in tdcgpu_create.  With a device present nothing is skipped — a missing libtdcgpu.so is
    then an error, never a silent pass."""
    try:
        import tudocomp_b200 as tdc
        have_gpu = tdc.load().device_count() > 0
    except Exception:  # library not built: let the gpu tests fail loudly if there is a device, skip if there is none
        have_gpu = os.path.exists("/dev/nvidiactl")
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """tests-only binding of oracle/libtdcoracle.so (the C restatement)."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "libtdcoracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libtdcoracle.so"])
        self.lib = ctypes.CDLL(path)
        self.lib.tdcoracle_lzss_lcp_factorize.restype = ctypes.c_int64
        self.lib.tdcoracle_plcp.restype = ctypes.c_uint32
        self.lib.tdcoracle_lzss_encode.restype = ctypes.c_int64
        self.lib.tdcoracle_rle_encode.restype = ctypes.c_int64
        self.lib.tdcoracle_literal_encode.restype = ctypes.c_int64

    def textds(self, t):
        n = t.size
        sa, isa, lcp, phi, plcp = (np.zeros(n, np.uint32) for _ in range(5))
        mx = ctypes.c_uint32()
        rc = self.lib.tdcoracle_textds(_P(t), ctypes.c_uint32(n), _P(sa), _P(isa), _P(lcp), _P(phi), _P(plcp), ctypes.byref(mx))
        assert rc == 0, rc
        return dict(sa=sa, isa=isa, lcp=lcp, phi=phi, plcp=plcp, max_lcp=int(mx.value))

    def bwt(self, t, sa):
        out = np.zeros(t.size, np.uint8)
        self.lib.tdcoracle_bwt(_P(t), _P(sa), ctypes.c_uint32(t.size), _P(out))
        return out

    def factorize(self, ds, n, threshold):
        out = np.zeros((max(n, 1), 3), np.uint32)
        z = self.lib.tdcoracle_lzss_lcp_factorize(_P(ds["sa"]), _P(ds["isa"]), _P(ds["lcp"]), ctypes.c_uint32(n),
                                                  ctypes.c_uint32(threshold), _P(out), ctypes.c_uint64(max(n, 1)))
        assert z >= 0
        return out[:z].copy()

    def factor_stats(self, triples, n):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        tr = np.ascontiguousarray(triples, np.uint32)
        self.lib.tdcoracle_factor_stats(_P(tr), ctypes.c_uint64(len(tr)), ctypes.c_uint32(n), ctypes.byref(a),
                                        ctypes.byref(b), ctypes.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def literal_histogram(self, text, triples):
        tr = np.ascontiguousarray(triples, np.uint32)
        hist = np.zeros(256, np.uint64)
        self.lib.tdcoracle_literal_histogram(_P(text), ctypes.c_uint32(text.size), _P(tr), ctypes.c_uint64(len(tr)), _P(hist))
        return hist

    def encode(self, text, triples, codes, lens, lead_bits=0, lead_byte=0, finalize=True):
        """lzss::encode_text restated: returns (bytes, nbits)."""
        tr = np.ascontiguousarray(triples, np.uint32)
        codes = np.ascontiguousarray(codes, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint8)
        cap = 13 * text.size + 64
        out = np.zeros(cap, np.uint8)
        nbits = ctypes.c_uint64()
        m = self.lib.tdcoracle_lzss_encode(_P(text), ctypes.c_uint32(text.size), _P(tr), ctypes.c_uint64(len(tr)), _P(codes),
                                           _P(lens), ctypes.c_uint32(lead_bits), ctypes.c_uint8(lead_byte),
                                           ctypes.c_int(1 if finalize else 0), _P(out), ctypes.c_uint64(cap), ctypes.byref(nbits))
        assert m >= 0
        return out[:m].copy(), int(nbits.value)

    def set_len_field_bits(self, bits):
        self.lib.tdcoracle_set_len_field_bits(ctypes.c_uint32(bits))

    def mtf_encode(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(data.size, np.uint8)
        self.lib.tdcoracle_mtf_encode(_P(data), ctypes.c_uint64(data.size), _P(out))
        return out

    def rle_encode(self, data, offset=0):
        data = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(12 * data.size + 16, np.uint8)
        m = self.lib.tdcoracle_rle_encode(_P(data), ctypes.c_uint64(data.size), ctypes.c_uint64(offset), _P(out), ctypes.c_uint64(out.size))
        assert m >= 0
        return out[:m].copy()

    def literal_encode(self, data, codes, lens, lead_bits=0, lead_byte=0, finalize=True):
        data = np.ascontiguousarray(data, np.uint8)
        codes = np.ascontiguousarray(codes, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint8)
        out = np.zeros(8 * data.size + 64, np.uint8)
        nbits = ctypes.c_uint64()
        m = self.lib.tdcoracle_literal_encode(_P(data), ctypes.c_uint64(data.size), _P(codes), _P(lens), ctypes.c_uint32(lead_bits),
                                              ctypes.c_uint8(lead_byte), ctypes.c_int(1 if finalize else 0), _P(out),
                                              ctypes.c_uint64(out.size), ctypes.byref(nbits))
        assert m >= 0
        return out[:m].copy(), int(nbits.value)

    def decode(self, triples, text):
        tr = np.ascontiguousarray(triples, np.uint32)
        out = np.zeros(text.size, np.uint8)
        rc = self.lib.tdcoracle_lzss_decode(_P(tr), ctypes.c_uint64(len(tr)), _P(text), ctypes.c_uint32(text.size), _P(out))
        assert rc == 0
        return out


class Reference:
    """tests-only binding of oracle/_ref/libtdcref.so (the unmodified reference compiled by oracle/Makefile)."""

    def __init__(self, wide=False):
        # wide: the same wrapper over the reference's wide-index build (-DLEN_BITS=40, oracle/Makefile)
        path = os.path.join(ROOT, "oracle", "_ref", "libtdcref40.so" if wide else "libtdcref.so")
        if not os.path.exists(path):
            pytest.skip(os.path.basename(path) + " not built (needs /root/reference; run `make -C oracle ref`)")
        self.lib = ctypes.CDLL(path)
        for f in ("tdcref_lzss_lcp_factors", "tdcref_lzss_lcp_compress", "tdcref_lzss_lcp_decompress", "tdcref_escape",
                  "tdcref_bwt_compress", "tdcref_stream_stage"):
            getattr(self.lib, f).restype = ctypes.c_int64
        self.lib.tdcref_last_error.restype = ctypes.c_char_p

    def textds(self, t):
        n = t.size
        sa, isa, lcp, phi, plcp = (np.zeros(n, np.uint32) for _ in range(5))
        mx = ctypes.c_uint32()
        rc = self.lib.tdcref_textds(_P(t), ctypes.c_uint64(n), _P(sa), _P(isa), _P(lcp), _P(phi), _P(plcp), ctypes.byref(mx))
        assert rc == 0, self.lib.tdcref_last_error()
        return dict(sa=sa, isa=isa, lcp=lcp, phi=phi, plcp=plcp, max_lcp=int(mx.value))

    def bwt(self, t):
        out = np.zeros(t.size, np.uint8)
        assert self.lib.tdcref_bwt(_P(t), ctypes.c_uint64(t.size), _P(out)) == 0
        return out

    def factors(self, t, threshold):
        out = np.zeros((max(t.size, 1), 3), np.uint32)
        hdr = np.zeros(3, np.uint64)
        z = self.lib.tdcref_lzss_lcp_factors(_P(t), ctypes.c_uint64(t.size), ctypes.c_uint32(threshold), _P(out),
                                             ctypes.c_uint64(max(t.size, 1)), _P(hdr))
        assert z >= 0, self.lib.tdcref_last_error()
        return out[:z].copy(), tuple(int(x) for x in hdr)

    def compress(self, t, threshold, coder):
        cap = 16 * t.size + 65536
        out = np.zeros(cap, np.uint8)
        secs = ctypes.c_double()
        m = self.lib.tdcref_lzss_lcp_compress(_P(t), ctypes.c_uint64(t.size), ctypes.c_uint32(threshold), coder, _P(out),
                                              ctypes.c_uint64(cap), ctypes.byref(secs))
        assert 0 <= m <= cap, self.lib.tdcref_last_error()
        return out[:m].copy(), secs.value

    def literal_coder(self, coder, hist):
        """The coder's header bits and literal code words for a literal histogram, from the reference's own
        HuffmanCoder / BitOStream code: (header bytes incl. the partial last byte, header_bits, codes[256], lens[256])."""
        hist = np.ascontiguousarray(hist, np.uint64)
        head = np.zeros(4096, np.uint8)
        bits = ctypes.c_uint64()
        codes = np.zeros(256, np.uint64)
        lens = np.zeros(256, np.uint8)
        rc = self.lib.tdcref_literal_coder(coder, _P(hist), _P(head), ctypes.c_uint64(head.size), ctypes.byref(bits), _P(codes), _P(lens))
        assert rc == 0, self.lib.tdcref_last_error()
        nb = (int(bits.value) + 7) // 8
        return head[:nb].copy(), int(bits.value), codes, lens

    def stream_stage(self, stage, data, offset=0):
        """0 = mtf, 1 = rle(offset), 2 = encode(bit), 3 = encode(huff) through the reference's own Compressor classes"""
        data = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(12 * data.size + 4096, np.uint8)
        m = self.lib.tdcref_stream_stage(stage, _P(data) if data.size else None, ctypes.c_uint64(data.size), ctypes.c_uint64(offset),
                                         _P(out), ctypes.c_uint64(out.size), None)
        assert 0 <= m <= out.size, self.lib.tdcref_last_error()
        return out[:m].copy()

    def decompress(self, arc, coder, n):
        out = np.zeros(n + 16, np.uint8)
        m = self.lib.tdcref_lzss_lcp_decompress(_P(arc), ctypes.c_uint64(arc.size), coder, _P(out), ctypes.c_uint64(out.size))
        assert m >= 0, self.lib.tdcref_last_error()
        return out[:m].copy()

    def escape(self, raw: bytes):
        a = np.frombuffer(raw, np.uint8).copy() if raw else np.zeros(0, np.uint8)
        out = np.zeros(2 * a.size + 16, np.uint8)
        m = self.lib.tdcref_escape(_P(a) if a.size else None, ctypes.c_uint64(a.size), _P(out), ctypes.c_uint64(out.size))
        assert m >= 0
        return out[:m].copy()


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    return Reference()
