"""CPU tests: pin the oracle (C restatement) against known answers, the committed golden vectors generated from the
unmodified reference, and — when oracle/_ref is present — the reference itself on fresh seeded inputs."""
import os

import numpy as np
import pytest

from inputs import all_small_cases, small_synthetic
from tudocomp_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _t(b):
    return np.frombuffer(b, np.uint8).copy()


def test_known_answers_from_survey(oracle):
    # recorded from the reference
    ds = oracle.textds(_t(b"abcdebcdeabc\0"))
    assert ds["sa"].tolist() == [12, 9, 0, 10, 5, 1, 11, 6, 2, 7, 3, 8, 4]
    assert ds["lcp"].tolist() == [0, 0, 3, 0, 2, 4, 0, 1, 3, 0, 2, 0, 1]
    assert ds["isa"].tolist() == [2, 5, 8, 10, 12, 4, 7, 9, 11, 1, 3, 6, 0]
    assert ds["phi"].tolist() == [9, 5, 6, 7, 8, 10, 11, 2, 3, 12, 0, 1, 4]
    assert ds["plcp"].tolist() == [3, 4, 3, 2, 1, 2, 1, 0, 0, 0, 0, 0, 4]  # PLCP[n-1] keeps Phi[n-1]
    assert bytes(oracle.bwt(_t(b"abcdebcdeabc\0"), ds["sa"])) == b"ce\0aeabbbccdd"
    assert oracle.factorize(ds, 13, 2).tolist() == [[5, 1, 4], [9, 0, 3]]
    ds = oracle.textds(_t(b"banana\0"))
    assert ds["sa"].tolist() == [6, 5, 3, 1, 0, 4, 2]
    assert ds["lcp"].tolist() == [0, 0, 1, 3, 0, 0, 2]
    assert bytes(oracle.bwt(_t(b"banana\0"), ds["sa"])) == b"annb\0aa"
    assert oracle.factorize(ds, 7, 3).tolist() == [[3, 1, 3]]


def test_escape_golden_bytes():
    # test/tudocomp_tests.cpp:528-533
    assert bytes(synth.escape_with_sentinel(b"\x00\x01\xff\xfe\x00")) == b"\xff\xfe\x01\xff\xff\xfe\xff\xfe\x00"


def test_oracle_matches_golden_vectors(oracle, gold):
    for name in gold["names"]:
        t = gold[f"{name}/text"]
        ds = oracle.textds(t)
        for k in ("sa", "isa", "lcp", "phi", "plcp"):
            assert np.array_equal(ds[k], gold[f"{name}/{k}"]), (name, k)
        assert ds["max_lcp"] == int(gold[f"{name}/max_lcp"][0]), name
        assert np.array_equal(oracle.bwt(t, ds["sa"]), gold[f"{name}/bwt"]), name
        for thr in (1, 2, 3, 5):
            f = oracle.factorize(ds, t.size, thr)
            assert np.array_equal(f, gold[f"{name}/factors{thr}"]), (name, thr)
            hdr = tuple(int(x) for x in gold[f"{name}/hdr{thr}"])
            mn, mx, dist = oracle.factor_stats(f, t.size)
            # the reference truncates INDEX_MAX (no factors) when encoding; compare on the same 32-bit footing
            assert (mn, mx, dist) == (hdr[0] & 0xFFFFFFFF, hdr[1], hdr[2]), (name, thr)
            assert np.array_equal(oracle.decode(f, t), t), (name, thr)


def test_golden_inputs_are_the_generated_ones(gold):
    # the fixtures must correspond to the seeded generators the GPU tests use
    cases = dict(all_small_cases())
    for name, t in cases.items():
        assert np.array_equal(gold[f"{name}/text"], t), name


def test_oracle_matches_reference_on_fresh_inputs(oracle, reference):
    cases = [("dna", synth.dna(50000, 11)), ("markov", synth.markov_text(50000, 12)),
             ("rep", synth.repetitive(50000, 13, block=777, p=0.02))]
    for name, t in cases:
        ds_o, ds_r = oracle.textds(t), reference.textds(t)
        for k in ("sa", "isa", "lcp", "phi", "plcp"):
            assert np.array_equal(ds_o[k], ds_r[k]), (name, k)
        assert ds_o["max_lcp"] == ds_r["max_lcp"]
        for thr in (2, 3, 7):
            f_r, _ = reference.factors(t, thr)
            assert np.array_equal(oracle.factorize(ds_o, t.size, thr), f_r), (name, thr)


def test_reference_archives_roundtrip(reference):
    # the reference's own round trip (test/matrix_tests.cpp) through our wrapper: pins the wrapper, not the oracle
    for name, t in small_synthetic():
        for coder in (0, 1, 2):
            arc, _ = reference.compress(t, 3, coder)
            assert np.array_equal(reference.decompress(arc, coder, t.size), t), (name, coder)
        break
