"""lzss::encode_text (compressors/lzss/LZSSCoding.hpp:18-92) on the device — SURVEY §8(f) row 1.

CPU (`-m "not gpu"`): the C restatement in oracle/tdc_oracle.c against the committed archives of the unmodified reference
(tests/golden) and against oracle/_ref on fresh inputs; the kernels' logic in the tests/sim interpreter.
GPU (`-m gpu`): the CUDA path through the C ABI against the golden archives, against oracle/_ref at 1 MiB, and at 8-16 MiB
against the oracle plus the reference's own decoder (archive -> text round trip).  Byte-exact everywhere."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from inputs import generator_strings, roundtrip_batch, small_synthetic
from tudocomp_b200 import _abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")
BIT, HUFF = 0, 1


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def bit_coder():
    """BitCoder: no header, every literal in 8 bits (coders/BitCoder.hpp, Coder.hpp:63-66 with LiteralRange)."""
    return np.zeros(0, np.uint8), 0, np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)


def splice(head, head_bits, body):
    """coder header (whole bytes) + device/oracle stream (which starts with the header's partial byte)"""
    return np.concatenate([head[:head_bits // 8], body])


def lead(head, head_bits):
    return head_bits % 8, (int(head[head_bits // 8]) if head_bits % 8 else 0)


def same_archive(got, want_arc_or_sha):
    if want_arc_or_sha.size == 32 and got.size != 32:  # large golden cases store the sha256 only
        return hashlib.sha256(got.tobytes()).digest() == want_arc_or_sha.tobytes()
    return np.array_equal(got, want_arc_or_sha)


def gold_coder(gold, name, coder):
    if coder == BIT:
        return bit_coder()
    return (gold[f"{name}/huff_head"], int(gold[f"{name}/huff_head_bits"][0]), gold[f"{name}/huff_codes"], gold[f"{name}/huff_lens"])


# ---------------------------------------------------------------------------------------------------------------------
# oracle
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_encode_matches_golden_archives(oracle, gold):
    for name in gold["names"]:
        t, f3 = gold[f"{name}/text"], gold[f"{name}/factors3"]
        assert np.array_equal(oracle.literal_histogram(t, f3), gold[f"{name}/lit_hist3"]), name
        for coder, cname in ((BIT, "bit"), (HUFF, "huff")):
            head, hb, codes, lens = gold_coder(gold, name, coder)
            lb, lbyte = lead(head, hb)
            body, _ = oracle.encode(t, f3, codes, lens, lb, lbyte)
            assert same_archive(splice(head, hb, body), gold[f"{name}/arc_{cname}"]), (name, cname)


def test_oracle_encode_matches_reference_on_fresh_inputs(oracle, reference):
    cases = [("dna", synth.dna(40000, 51)), ("markov", synth.markov_text(40000, 52)),
             ("rep", synth.repetitive(40000, 53, block=555, p=0.02)),
             ("random_bytes", synth.with_sentinel(np.random.default_rng(54).integers(1, 255, 30000, dtype=np.uint8))),
             ("one_symbol", synth.with_sentinel(np.full(5000, 66, np.uint8))), ("sentinel_only", np.zeros(1, np.uint8))]
    for name, t in cases:
        for thr in (2, 3, 40):
            f, _ = reference.factors(t, thr)
            hist = oracle.literal_histogram(t, f)
            for coder in (BIT, HUFF):
                head, hb, codes, lens = reference.literal_coder(coder, hist)
                lb, lbyte = lead(head, hb)
                body, _ = oracle.encode(t, f, codes, lens, lb, lbyte)
                arc, _ = reference.compress(t, thr, coder)
                assert np.array_equal(splice(head, hb, body), arc), (name, thr, coder)


def _wide_cases(size=30000):
    return [("dna", synth.dna(size, 61)), ("markov", synth.markov_text(size, 62)), ("rep", synth.repetitive(size, 63, block=400, p=0.02)),
            ("no_factor", synth.with_sentinel(np.arange(1, 200, dtype=np.uint8))), ("sentinel_only", np.zeros(1, np.uint8)),
            ("banana", synth.with_sentinel(np.frombuffer(b"banana", np.uint8)))]


def _wide_format_check(encode_body, size=30000, thresholds=(2, 5)):
    """The LEN_BITS=40 archive format (text length in 64 bits): `encode_body(t, thr, f, codes, lens,
    lead_bits, lead_byte)` must reproduce the archive of the reference compiled with -DLEN_BITS=40 (oracle/_ref/libtdcref40.so)."""
    from conftest import Reference
    wide, narrow = Reference(wide=True), Reference()
    for name, t in _wide_cases(size):
        for thr in thresholds:
            f, _ = wide.factors(t, thr)
            assert np.array_equal(f, narrow.factors(t, thr)[0]), name  # same factors, only the header differs
            from conftest import Oracle
            hist = Oracle().literal_histogram(t, f)
            for coder in (BIT, HUFF):
                head, hb, codes, lens = wide.literal_coder(coder, hist)
                lb, lbyte = lead(head, hb)
                arc, _ = wide.compress(t, thr, coder)
                assert arc.size == narrow.compress(t, thr, coder)[0].size + 4, name  # the 32 extra bits of the length field
                got = splice(head, hb, encode_body(t, thr, f, codes, lens, lb, lbyte))
                assert np.array_equal(got, arc), (name, thr, coder)


def test_oracle_encode_wide_index_format(oracle):
    def body(t, thr, f, codes, lens, lb, lbyte):
        oracle.set_len_field_bits(64)
        try:
            return oracle.encode(t, f, codes, lens, lb, lbyte)[0]
        finally:
            oracle.set_len_field_bits(32)
    _wide_format_check(body)


def _device_wide_body(lib, device=0):
    def body(t, thr, f, codes, lens, lb, lbyte):
        with _abi.Context(lib, device) as c:
            c.set_text(t)
            c.factorize(thr)
            c.literal_histogram()
            c.set_len_bits(64)
            nbits = c.encode(codes, lens, lb, lbyte)
            return c.encoded(nbits)
    return body


# ---------------------------------------------------------------------------------------------------------------------
# kernels in the CPU interpreter
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def simlib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    return _abi.TdcGpuLib(SIM)


def _device_vs_oracle(lib, oracle, name, t, thr, coder_of, check_archive=None, device=0):
    ds = oracle.textds(t)
    want = oracle.factorize(ds, t.size, thr)
    with _abi.Context(lib, device) as c:
        c.set_text(t)
        c.factorize(thr)
        hist, fdist = c.literal_histogram()
        assert np.array_equal(hist, oracle.literal_histogram(t, want)), (name, thr)
        assert fdist == oracle.factor_stats(want, t.size)[2], (name, thr)
        for coder in (BIT, HUFF):
            head, hb, codes, lens = coder_of(coder, hist)
            lb, lbyte = lead(head, hb)
            body, nbits = oracle.encode(t, want, codes, lens, lb, lbyte)
            assert c.encode(codes, lens, lb, lbyte) == nbits, (name, thr, coder)
            got = c.encoded(nbits)
            assert np.array_equal(got, body), (name, thr, coder)
            for chunk in (7, 4096):  # the same stream drained in pieces, tail bytes split across pieces included
                if got.size <= 40000 or chunk > 7:
                    assert np.array_equal(c.encoded_chunks(chunk), body), (name, thr, coder, chunk)
            raw = c.encoded(nbits, finalize=False)
            assert raw.size == (nbits + 7) // 8 and np.array_equal(raw[:nbits // 8], body[:nbits // 8])
            if check_archive:
                check_archive(coder, splice(head, hb, got))


@pytest.mark.sim
def test_sim_encode_reference_strings(simlib, oracle, reference):
    def coder_of(coder, hist):
        return reference.literal_coder(coder, hist)
    for name, t in roundtrip_batch():
        for thr in (2,):
            _device_vs_oracle(simlib, oracle, name, t, thr, coder_of,
                              lambda coder, arc: np.testing.assert_array_equal(arc, reference.compress(t, thr, coder)[0]))
    for name, t in generator_strings(7):
        _device_vs_oracle(simlib, oracle, name, t, 1, coder_of)
    for name, t in small_synthetic():
        if t.size <= 21000:
            _device_vs_oracle(simlib, oracle, name, t, 3, coder_of,
                              lambda coder, arc: np.testing.assert_array_equal(arc, reference.compress(t, 3, coder)[0]))


@pytest.mark.sim
def test_sim_encode_wide_index_format(simlib):
    _wide_format_check(_device_wide_body(simlib), size=4000, thresholds=(3,))  # (the interpreter is slow: small texts)


@pytest.mark.sim
def test_sim_encode_long_codes_and_lead_bits(simlib, oracle):
    """Code words longer than 32 bits (deep Huffman trees), every lead-bit offset, state errors."""
    rng = np.random.default_rng(3)
    t = synth.with_sentinel(rng.integers(1, 40, 6000, dtype=np.uint8))
    lens = rng.integers(1, 65, 256).astype(np.uint8)
    codes = rng.integers(0, 1 << 62, 256, dtype=np.uint64) >> (np.uint64(64) - lens.astype(np.uint64)).clip(0, 63)
    for lb in range(8):
        _device_vs_oracle(simlib, oracle, f"lead{lb}", t, 3, lambda coder, hist: (np.array([0xA5], np.uint8) if lb else np.zeros(0, np.uint8), lb, codes, lens))
    with _abi.Context(simlib) as c:
        c.set_text(t)
        with pytest.raises(_abi.TdcGpuError):
            c.literal_histogram()  # no factor list yet
        c.factorize(3)
        with pytest.raises(_abi.TdcGpuError):
            c.encoded(100)  # nothing encoded yet


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpulib():
    import tudocomp_b200 as tdc
    return tdc.load()


@pytest.mark.gpu
def test_gpu_encode_golden_archives(gpulib, gold):
    for name in gold["names"]:
        t = gold[f"{name}/text"]
        with _abi.Context(gpulib) as c:
            c.set_text(t)
            c.factorize(3)
            hist, _ = c.literal_histogram()
            assert np.array_equal(hist, gold[f"{name}/lit_hist3"]), name
            for coder, cname in ((BIT, "bit"), (HUFF, "huff")):
                head, hb, codes, lens = gold_coder(gold, name, coder)
                lb, lbyte = lead(head, hb)
                nbits = c.encode(codes, lens, lb, lbyte)
                assert same_archive(splice(head, hb, c.encoded(nbits)), gold[f"{name}/arc_{cname}"]), (name, cname)


@pytest.mark.gpu
def test_gpu_encode_wide_index_format(gpulib):
    _wide_format_check(_device_wide_body(gpulib))


@pytest.mark.gpu
def test_gpu_encode_vs_reference_1m(gpulib, oracle, reference):
    for name, t in (("dna_1m", synth.dna(1 << 20, 31)), ("markov_1m", synth.markov_text(1 << 20, 32)),
                    ("rep_1m", synth.repetitive(1 << 20, 33, block=3000, p=0.01))):
        for thr in (3, 5):
            with _abi.Context(gpulib) as c:
                c.set_text(t)
                c.factorize(thr)
                hist, _ = c.literal_histogram()
                for coder in (BIT, HUFF):
                    head, hb, codes, lens = reference.literal_coder(coder, hist)
                    lb, lbyte = lead(head, hb)
                    nbits = c.encode(codes, lens, lb, lbyte)
                    arc, _ = reference.compress(t, thr, coder)
                    assert np.array_equal(splice(head, hb, c.encoded(nbits)), arc), (name, thr, coder)


@pytest.mark.gpu
def test_gpu_encode_large_roundtrip(gpulib, oracle, reference):
    """8-16 MiB: device stream == oracle stream on the device's own factor list, and the reference's decoder turns the
    archive back into the text (lzss::decode_text, LZSSCoding.hpp:94-140)."""
    for name, t in (("dna_16m", synth.dna(1 << 24, 41)), ("markov_8m", synth.markov_text(1 << 23, 42)),
                    ("repetitive_8m", synth.repetitive(1 << 23, 43, block=1 << 16, p=0.01))):
        with _abi.Context(gpulib) as c:
            c.set_text(t)
            z, _, _ = c.factorize(3)
            f = c.factors(z)
            tr = np.stack([f["pos"], f["src"], f["len"]], 1)
            hist, fdist = c.literal_histogram()
            assert np.array_equal(hist, oracle.literal_histogram(t, tr)), name
            assert fdist == oracle.factor_stats(tr, t.size)[2]
            for coder in (BIT, HUFF):
                head, hb, codes, lens = reference.literal_coder(coder, hist)
                lb, lbyte = lead(head, hb)
                nbits = c.encode(codes, lens, lb, lbyte)
                body, onbits = oracle.encode(t, tr, codes, lens, lb, lbyte)
                assert nbits == onbits
                got = c.encoded(nbits)
                assert np.array_equal(got, body), (name, coder)
                assert np.array_equal(reference.decompress(splice(head, hb, got), coder, t.size), t), (name, coder)
