"""Bit-exactness at the sizes BASELINE.json's configs name (`-m gpu`).

The unmodified reference was run offline on exactly these seeded texts (tests/golden/make_size_hashes.py, 15-25 CPU
minutes per 2^30 B job) and left SHA-256 fingerprints of SA, ISA, LCP, BWT, the lzss_lcp(threshold=3) factor list and the
raw `bit` / `huff` archives in tests/golden/size_hashes.json.  Here the CUDA path computes the same objects through the
C ABI and must reproduce every fingerprint: 100 MB Markov (config 1), 2^30 B DNA (config 2 / the bench headline),
2^30 B repetitive (config 3), plus 2^26 / 2^28 B cases that are cheap enough for every run.

TDC_SIZE_CASES=dna_2p26,markov_2p26 restricts the cases (development runs on a metered GPU box)."""
import hashlib
import json
import os

import numpy as np
import pytest

import tudocomp_b200 as tdc
from tudocomp_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
HASHES = json.load(open(os.path.join(HERE, "golden", "size_hashes.json")))
_sel = [c for c in os.environ.get("TDC_SIZE_CASES", "").split(",") if c]
CASES = [c for c in sorted(HASHES, key=lambda k: HASHES[k]["n"]) if not _sel or c in _sel]
BIT, HUFF = 0, 1


def sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


def gen(case):
    kind, n_body, seed = HASHES[case]["generator"]
    return {"dna": synth.dna, "markov": synth.markov_text, "repetitive": synth.repetitive}[kind](n_body, seed)


@pytest.fixture(scope="module")
def ctx():
    c = tdc.Context(tdc.load(), 0)
    yield c
    c.close()


@pytest.mark.parametrize("case", CASES)
def test_reference_fingerprints_at_config_sizes(ctx, case):
    want = HASHES[case]
    t = gen(case)
    assert t.size == want["n"] and sha(t) == want["text_sha256"], "the synthetic generator is not reproducible on this box"
    ctx.set_text(t)
    if "index" in want:
        w = want["index"]
        ctx.build(tdc.SA | tdc.ISA | tdc.LCP | tdc.BWT)
        for name, flag in (("sa", tdc.SA), ("isa", tdc.ISA), ("lcp", tdc.LCP), ("bwt", tdc.BWT)):
            a = ctx.get(flag)
            assert sha(a) == w[name], (case, name)
            del a
        assert ctx.max_lcp() == w["max_lcp"], case
    if "factors" not in want:
        return
    w = want["factors"]
    z, mn, mx = ctx.factorize(w["threshold"])
    assert z == w["z"], (case, z, w["z"])
    assert (mn, mx) == (w["flen_min"] & 0xFFFFFFFF, w["flen_max"]), case
    f = ctx.factors(z)
    assert sha(f) == w["factors"], case
    del f
    hist, fdist = ctx.literal_histogram()
    assert fdist == w["fdist_max"], case
    for coder, cname in ((BIT, "bit"), (HUFF, "huff")):
        if cname not in want:
            continue
        if coder == BIT:  # BitCoder: no header, every literal in 8 bits (Coder.hpp:63-66)
            head, hb, codes, lens = np.zeros(0, np.uint8), 0, np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)
        else:  # the Huffman table is the reference's own code (tie order), from the device's literal histogram
            from conftest import Reference
            head, hb, codes, lens = Reference().literal_coder(HUFF, hist)
        lb, lbyte = hb % 8, (int(head[hb // 8]) if hb % 8 else 0)
        nbits = ctx.encode(codes, lens, lb, lbyte)
        body = ctx.encoded(nbits)
        h = hashlib.sha256()
        h.update(head[:hb // 8].tobytes())
        h.update(memoryview(body).cast("B"))
        assert head[:hb // 8].size + body.size == want[cname]["archive_len"], (case, cname)
        assert h.hexdigest() == want[cname]["archive"], (case, cname)
