"""The stream stages behind the BWT in `bwt:mtf:rle:encode(huff)` (BASELINE config 3):
mtf_encode (MTFCompressor.hpp:46-56), rle_encode (RunLengthEncoder.hpp:15-31), LiteralEncoder (LiteralEncoder.hpp:23-32).

CPU: the oracle restatements against the committed outputs of the unmodified reference (tests/golden/stream_stage_vectors.npz)
and against oracle/_ref on fresh inputs; the kernels in the tests/sim interpreter.  GPU: the CUDA path through the C ABI
against the golden outputs, and against the oracle at 16-64 MiB.  Byte-exact."""
import os
import subprocess

import numpy as np
import pytest

from tudocomp_b200 import _abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "stream_stage_vectors.npz")
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _lead(head, bits):
    return bits % 8, (int(head[bits // 8]) if bits % 8 else 0)


def _huff(gold, name):
    return gold[f"{name}/huff_head"], int(gold[f"{name}/huff_head_bits"][0]), gold[f"{name}/huff_codes"], gold[f"{name}/huff_lens"]


def test_oracle_matches_golden_stage_outputs(oracle, gold):
    for name in gold["names"]:
        d = gold[f"{name}/in"]
        assert np.array_equal(oracle.mtf_encode(d), gold[f"{name}/mtf"]), name
        for off in (0, 1, 300):
            assert np.array_equal(oracle.rle_encode(d, off), gold[f"{name}/rle{off}"]), (name, off)
        body, _ = oracle.literal_encode(d, np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8))
        assert np.array_equal(body, gold[f"{name}/enc_bit"]), name
        head, hb, codes, lens = _huff(gold, name)
        lb, lbyte = _lead(head, hb)
        body, _ = oracle.literal_encode(d, codes, lens, lb, lbyte)
        assert np.array_equal(np.concatenate([head[:hb // 8], body]), gold[f"{name}/enc_huff"]), name
    assert np.array_equal(oracle.mtf_encode(gold["chain/bwt"]), gold["chain/mtf"])
    assert np.array_equal(oracle.rle_encode(gold["chain/mtf"], 0), gold["chain/rle"])


def test_oracle_matches_reference_on_fresh_inputs(oracle, reference):
    rng = np.random.default_rng(77)
    for name, d in (("random", rng.integers(0, 256, 30000, dtype=np.uint8)), ("dna", synth.dna(30000, 5)[:-1]),
                    ("runs", np.repeat(rng.integers(0, 256, 500, dtype=np.uint8), rng.integers(1, 200, 500)))):
        assert np.array_equal(oracle.mtf_encode(d), reference.stream_stage(0, d)), name
        assert np.array_equal(oracle.rle_encode(d, 7), reference.stream_stage(1, d, 7)), name


def _device_checks(lib, oracle, gold, names, device=0):
    with _abi.Context(lib, device) as c:
        for name in names:
            d = gold[f"{name}/in"]
            assert np.array_equal(c.mtf_encode(d), gold[f"{name}/mtf"]), name
            for off in (0, 1, 300):
                assert np.array_equal(c.rle_encode(d, off), gold[f"{name}/rle{off}"]), (name, off)
            assert np.array_equal(c.literal_histogram_of(d), np.bincount(d, minlength=256).astype(np.uint64)), name
            assert np.array_equal(c.literal_encode(np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)), gold[f"{name}/enc_bit"]), name
            head, hb, codes, lens = _huff(gold, name)
            lb, lbyte = _lead(head, hb)
            c.literal_histogram_of(d)
            body = c.literal_encode(codes, lens, lb, lbyte)
            assert np.array_equal(np.concatenate([head[:hb // 8], body]), gold[f"{name}/enc_huff"]), name
        assert np.array_equal(c.mtf_encode(gold["chain/bwt"]), gold["chain/mtf"])
        assert np.array_equal(c.rle_encode(gold["chain/mtf"], 0), gold["chain/rle"])


@pytest.mark.sim
def test_sim_stream_stages(oracle, gold):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    lib = _abi.TdcGpuLib(SIM)
    small = [n for n in gold["names"] if gold[f"{n}/in"].size <= 5000] + ["two_symbols", "runs_mixed"]
    _device_checks(lib, oracle, gold, small)
    with _abi.Context(lib) as c:
        with pytest.raises(_abi.TdcGpuError):
            c.literal_encode(np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8))  # nothing staged


def _device_resident_pipeline(lib, oracle, data, offset=0):
    """mtf -> rle -> encode(bit) with the bytes between the stages in DEVICE memory (the C-ABI calls of the plugin's
    GpuChainCompressor): host in -> device (TDCGPU_BUF_OUT_DEVICE), device -> device, device -> coded stream on the host."""
    import ctypes as C
    L = lib.lib
    L.tdcgpu_device_alloc.restype = C.c_void_p
    L.tdcgpu_device_alloc.argtypes = [C.c_void_p, C.c_uint64]
    L.tdcgpu_device_free.argtypes = [C.c_void_p, C.c_void_p]
    L.tdcgpu_device_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.tdcgpu_mtf_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.tdcgpu_rle_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    L.tdcgpu_literal_encode_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    data = np.ascontiguousarray(data, np.uint8)
    n = data.size
    worst = 12 * n + 32
    codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)
    with _abi.Context(lib) as c:
        h = c._h
        d_m, d_r = L.tdcgpu_device_alloc(h, n), L.tdcgpu_device_alloc(h, worst)
        assert d_m and d_r
        try:
            lib.check(L.tdcgpu_mtf_encode(h, data.ctypes.data, n, d_m, 2))        # TDCGPU_BUF_OUT_DEVICE
            produced = C.c_uint64()
            lib.check(L.tdcgpu_rle_encode(h, d_m, n, offset, d_r, worst, C.byref(produced), 1))  # TDCGPU_BUF_DEVICE
            back = np.empty(produced.value, np.uint8)
            lib.check(L.tdcgpu_device_copy(h, back.ctypes.data, d_r, produced.value, 1))
            want_r = oracle.rle_encode(oracle.mtf_encode(data), offset)
            assert np.array_equal(back, want_r)
            m_host = np.empty(n, np.uint8)
            lib.check(L.tdcgpu_mtf_encode(h, d_m, n, m_host.ctypes.data, 3))      # TDCGPU_BUF_IN_DEVICE (mtf of the mtf output)
            assert np.array_equal(m_host, oracle.mtf_encode(oracle.mtf_encode(data)))
            hist = np.zeros(256, np.uint64)
            lib.check(L.tdcgpu_literal_encode_begin(h, d_r, produced.value, 1, hist.ctypes.data))
            assert np.array_equal(hist, np.bincount(want_r, minlength=256).astype(np.uint64))
            got = c.literal_encode(codes, lens)
            assert np.array_equal(got, oracle.literal_encode(want_r, codes, lens)[0])
            with pytest.raises(_abi.TdcGpuError):
                lib.check(L.tdcgpu_mtf_encode(h, data.ctypes.data, n, d_m, 7))    # not a TDCGPU_BUF_* value
        finally:
            L.tdcgpu_device_free(h, d_m)
            L.tdcgpu_device_free(None, d_r)


@pytest.mark.sim
def test_sim_device_resident_stage_pipeline(oracle):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    lib = _abi.TdcGpuLib(SIM)
    rng = np.random.default_rng(5)
    for data, off in ((np.repeat(rng.integers(0, 256, 300, dtype=np.uint8), rng.integers(1, 9, 300)), 0),
                      (rng.integers(97, 101, 4000, dtype=np.uint8), 3), (np.full(1, 65, np.uint8), 0)):
        _device_resident_pipeline(lib, oracle, data, off)


@pytest.mark.gpu
def test_gpu_device_resident_stage_pipeline(oracle):
    import tudocomp_b200 as tdc
    rng = np.random.default_rng(6)
    for data, off in ((np.repeat(rng.integers(0, 256, 1 << 14, dtype=np.uint8), rng.integers(1, 300, 1 << 14)), 0),
                      (rng.integers(97, 101, 1 << 22, dtype=np.uint8), 5)):
        _device_resident_pipeline(tdc.load(), oracle, data, off)


@pytest.mark.gpu
def test_gpu_stream_stages_golden(oracle, gold):
    import tudocomp_b200 as tdc
    _device_checks(tdc.load(), oracle, gold, list(gold["names"]))


@pytest.mark.gpu
def test_gpu_stream_stages_large(oracle):
    """16-64 MiB: a real BWT (from the GPU) through mtf -> rle -> encode against the oracle; random bytes (MTF worst case)."""
    import tudocomp_b200 as tdc
    lib = tdc.load()
    rng = np.random.default_rng(9)
    with _abi.Context(lib, 0) as c:
        t = synth.repetitive(1 << 24, 3, block=1 << 16, p=0.01)
        c.set_text(t)
        c.build(tdc.SA | tdc.BWT)
        inputs = {"bwt_repetitive_16m": c.get(tdc.BWT), "random_16m": rng.integers(0, 256, 1 << 24, dtype=np.uint8),
                  "runs_64m": np.repeat(rng.integers(0, 256, 1 << 16, dtype=np.uint8), 1024)}
        for name, d in inputs.items():
            m = c.mtf_encode(d)
            assert np.array_equal(m, oracle.mtf_encode(d)), name
            r = c.rle_encode(m, 0)
            assert np.array_equal(r, oracle.rle_encode(m, 0)), name
            hist = c.literal_histogram_of(r)
            assert np.array_equal(hist, np.bincount(r, minlength=256).astype(np.uint64)), name
            codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)
            want, _ = oracle.literal_encode(r, codes, lens)
            assert np.array_equal(c.literal_encode(codes, lens), want), name


@pytest.mark.sim
def test_sim_device_pointer_variants(oracle):
    """The on_device / to_device flavours of the entry points (device pointers in, device pointers out).  In the interpreter
    build device memory is host memory, so the same buffers serve as "device" pointers: this checks the staging logic
    (no copy-in, caller-owned output, tail bytes written through the device path), not the CUDA copies themselves."""
    import ctypes as C

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    lib = _abi.TdcGpuLib(SIM)
    L = lib.lib
    rng = np.random.default_rng(31)
    d = np.repeat(rng.integers(0, 200, 700, dtype=np.uint8), rng.integers(1, 9, 700))
    with _abi.Context(lib) as c:
        h = c._h
        # mtf: device in, device out
        out = np.zeros(d.size + 64, np.uint8)
        lib.check(L.tdcgpu_mtf_encode(h, d.ctypes.data, d.size, out.ctypes.data, 1))
        assert np.array_equal(out[:d.size], oracle.mtf_encode(d))
        # rle: a device output buffer must hold the worst case
        want = oracle.rle_encode(d, 5)
        small = np.zeros(want.size + 8, np.uint8)
        m = C.c_uint64()
        assert L.tdcgpu_rle_encode(h, d.ctypes.data, d.size, 5, small.ctypes.data, small.size, C.byref(m), 1) == -5
        big = np.zeros(3 * d.size + 64, np.uint8)
        lib.check(L.tdcgpu_rle_encode(h, d.ctypes.data, d.size, 5, big.ctypes.data, big.size, C.byref(m), 1))
        assert np.array_equal(big[:m.value], want)
        # literal encoder: device input, stream fetched to a "device" buffer with and without the tail
        hist = np.zeros(256, np.uint64)
        lib.check(L.tdcgpu_literal_encode_begin(h, d.ctypes.data, d.size, 1, hist.ctypes.data))
        assert np.array_equal(hist, np.bincount(d, minlength=256).astype(np.uint64))
        codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 7, np.uint8)  # 7-bit words: the stream ends mid-byte
        nbits = C.c_uint64()
        lib.check(L.tdcgpu_literal_encode(h, codes.ctypes.data, lens.ctypes.data, 3, 0xA0, C.byref(nbits)))
        want, wbits = oracle.literal_encode(d, codes & np.uint64(0x7F), lens, 3, 0xA0)
        assert nbits.value == wbits
        for fin in (1, 0):
            buf = np.zeros(want.size + 8, np.uint8)
            nb = C.c_uint64()
            lib.check(L.tdcgpu_literal_encode_get(h, buf.ctypes.data, buf.size, fin, C.byref(nb), 1))
            ref_bytes = want if fin else oracle.literal_encode(d, codes & np.uint64(0x7F), lens, 3, 0xA0, finalize=False)[0]
            assert np.array_equal(buf[:nb.value], ref_bytes), fin
        # text index: packed array and encoded lzss stream into "device" buffers
        t = synth.markov_text(3000, 8)
        c.set_text(t)
        c.build(_abi.SA)
        w = int(t.size).bit_length()
        host = c.get_packed(_abi.SA, w)
        dev = np.zeros(host.size, np.uint64)
        lib.check(L.tdcgpu_textds_get_packed(h, _abi.SA, w, dev.ctypes.data, dev.size, 1))
        assert np.array_equal(dev, host)
        c.factorize(3)
        c.literal_histogram()
        nbits = c.encode(np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8))
        host = c.encoded(nbits)
        dev = np.zeros(host.size + 4, np.uint8)
        nb = C.c_uint64()
        lib.check(L.tdcgpu_lzss_encode_get(h, dev.ctypes.data, dev.size, 1, C.byref(nb), 1))
        assert np.array_equal(dev[:nb.value], host)
