"""Seeded differential fuzz of the kernels added behind the factoriser (device-side encode_text, bit-packed arrays, mtf /
rle / literal encoder) in the tests/sim interpreter: random sizes around the tile / chunk / word boundaries, random
alphabets (incl. bytes >= 0x80 in runs), random thresholds, offsets, coders and widths — against the oracle, and every
few iterations the oracle against the unmodified reference (oracle/_ref).  CPU only; ~40 s (20 seeded iterations)."""
import os
import subprocess

import numpy as np
import pytest

from tudocomp_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")
SIZES = [1, 2, 3, 15, 16, 17, 31, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1000, 1024, 2047, 2048, 2049, 3000, 4097, 7000]

pytestmark = pytest.mark.sim


def _input(rng, n):
    kind = int(rng.integers(0, 5))
    if kind == 0:
        d = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == 1:
        d = rng.integers(0, int(rng.integers(1, 5)) + 1, n, dtype=np.uint8) + np.uint8(rng.choice([0, 100, 200, 250]))
    elif kind == 2:
        d = np.repeat(rng.integers(0, 256, n, dtype=np.uint8), rng.integers(1, int(rng.integers(2, 40)), n))[:n]
    elif kind == 3:
        d = np.repeat(rng.integers(120, 136, n, dtype=np.uint8), rng.integers(1, 400, n))[:n]
    else:
        d = np.full(n, rng.integers(0, 256), np.uint8)
    if d.size >= 2 and d[-1] == 255 and d[-2] == 255:
        d[-1] = 7  # the reference's rle_encode spins forever on a trailing 0xFF 0xFF (oracle/tdc_oracle.c)
    return np.ascontiguousarray(d)


def test_fuzz_new_kernels_against_oracle_and_reference(oracle, reference):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    lib = _abi.TdcGpuLib(SIM)
    rng = np.random.default_rng(20261017)
    with _abi.Context(lib) as c:
        for it in range(20):
            d = _input(rng, int(rng.choice(SIZES)))
            off = int(rng.choice([0, 1, 127, 128, 5000, 1 << 20]))
            tag = (it, d.size, off)
            # stream stages
            assert np.array_equal(c.mtf_encode(d), oracle.mtf_encode(d)), ("mtf",) + tag
            want = oracle.rle_encode(d, off)
            assert np.array_equal(c.rle_encode(d, off), want), ("rle",) + tag
            hist = c.literal_histogram_of(d)
            coder = int(rng.integers(0, 2))
            head, hb, codes, lens = reference.literal_coder(coder, hist)
            lb, lbyte = hb % 8, (int(head[hb // 8]) if hb % 8 else 0)
            lit, _ = oracle.literal_encode(d, codes, lens, lb, lbyte)
            assert np.array_equal(c.literal_encode(codes, lens, lb, lbyte), lit), ("literal",) + tag
            if it % 6 == 0:
                assert np.array_equal(oracle.mtf_encode(d), reference.stream_stage(0, d)), ("mtf oracle/ref",) + tag
                assert np.array_equal(want, reference.stream_stage(1, d, off)), ("rle oracle/ref",) + tag
                assert np.array_equal(np.concatenate([head[:hb // 8], lit]), reference.stream_stage(2 + coder, d)), ("literal oracle/ref",) + tag
            # lzss_lcp: factor list, literal histogram, archive body, packed SA
            t = np.concatenate([np.where(d == 0, 1, d), np.zeros(1, np.uint8)]).astype(np.uint8)
            thr = int(rng.choice([1, 2, 3, 5, 30]))
            ds = oracle.textds(t)
            fw = oracle.factorize(ds, t.size, thr)
            c.set_text(t)
            z, _, _ = c.factorize(thr)
            f = c.factors(z)
            got = np.stack([f["pos"], f["src"], f["len"]], 1) if z else np.zeros((0, 3), np.uint32)
            assert np.array_equal(got, fw), ("factors", thr) + tag
            hist, _ = c.literal_histogram()
            assert np.array_equal(hist, oracle.literal_histogram(t, fw)), ("lzss histogram", thr) + tag
            head, hb, codes, lens = reference.literal_coder(coder, hist)
            lb, lbyte = hb % 8, (int(head[hb // 8]) if hb % 8 else 0)
            body, nbits = oracle.encode(t, fw, codes, lens, lb, lbyte)
            assert c.encode(codes, lens, lb, lbyte) == nbits, ("lzss bits", thr) + tag
            assert np.array_equal(c.encoded(nbits), body), ("lzss encode", thr, coder) + tag
            if it % 6 == 0:
                arc, _ = reference.compress(t, thr, coder)
                assert np.array_equal(np.concatenate([head[:hb // 8], body]), arc), ("lzss oracle/ref", thr, coder) + tag
            w = int(rng.integers(1, 33))
            a = c.get(_abi.SA)
            bits = (((a.astype(np.uint64) & np.uint64((1 << w) - 1))[:, None] >> np.arange(w, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8).ravel()
            bits = np.concatenate([bits, np.zeros((-bits.size) % 64, np.uint8)])
            assert np.array_equal(c.get_packed(_abi.SA, w), np.packbits(bits, bitorder="little").view(np.uint64)), ("pack", w) + tag
