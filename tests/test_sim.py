"""CPU tests: the CUDA kernels' LOGIC, executed by the test-only interpreter tests/sim/cusim.h, against the oracle.
This is how kernels are debugged without a GPU in the build container; it proves nothing about the real device
(memory ordering, occupancy), which the `-m gpu` tests cover."""
import os
import subprocess

import numpy as np
import pytest

from inputs import generator_strings, roundtrip_batch, small_synthetic
from tudocomp_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")

pytestmark = pytest.mark.sim


@pytest.fixture(scope="module")
def simlib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    return _abi.TdcGpuLib(SIM)


def _check(simlib, oracle, name, t, thresholds=(3,)):
    ds = oracle.textds(t)
    with _abi.Context(simlib) as c:
        c.set_text(t)
        c.build(_abi.SA | _abi.ISA | _abi.LCP | _abi.PHI | _abi.PLCP | _abi.BWT)
        for k, fl in (("sa", _abi.SA), ("isa", _abi.ISA), ("lcp", _abi.LCP), ("phi", _abi.PHI), ("plcp", _abi.PLCP)):
            assert np.array_equal(c.get(fl), ds[k]), (name, k)
        assert c.max_lcp() == ds["max_lcp"], name
        assert np.array_equal(c.get(_abi.BWT), oracle.bwt(t, ds["sa"])), name
        for thr in thresholds:
            z, mn, mx = c.factorize(thr)
            f = c.factors(z)
            want = oracle.factorize(ds, t.size, thr)
            got = np.stack([f["pos"], f["src"], f["len"]], 1) if z else np.zeros((0, 3), np.uint32)
            assert np.array_equal(got, want), (name, thr)
            wmn, wmx, _ = oracle.factor_stats(want, t.size)
            assert (mn, mx) == (wmn, wmx), (name, thr)


def test_sim_reference_test_strings(simlib, oracle):
    for name, t in roundtrip_batch():
        _check(simlib, oracle, name, t, thresholds=(1, 2, 3))


def test_sim_generator_strings(simlib, oracle):
    for name, t in generator_strings(9):
        _check(simlib, oracle, name, t, thresholds=(2, 3))


def test_sim_synthetic_multi_tile(simlib, oracle):
    for name, t in small_synthetic():
        if t.size > 21000:
            continue  # keep the CPU suite short; the larger cases run on the GPU
        _check(simlib, oracle, name, t, thresholds=(3, 5))


def test_sim_lcp_only_routes(simlib, oracle):
    """LCP without Phi/PLCP: short-prefix texts take the direct comparison route, repetitive ones the Phi route."""
    from tudocomp_b200 import synth
    cases = [("dna", synth.dna(12000, 3), 1), ("markov", synth.markov_text(9000, 4), 1),
             ("repetitive", synth.repetitive(12000, 5, block=300, p=0.005), 2),
             ("run_a", synth.with_sentinel(np.full(3000, 97, np.uint8)), 2),
             ("long_repeat_in_random", synth.with_sentinel(np.concatenate([np.random.default_rng(1).integers(97, 101, 4000, dtype=np.uint8)] * 2 + [np.random.default_rng(2).integers(97, 101, 9000, dtype=np.uint8)])), None)]
    for name, t, route in cases:
        ds = oracle.textds(t)
        with _abi.Context(simlib) as c:
            c.set_text(t)
            c.build(_abi.SA | _abi.ISA | _abi.LCP)
            assert np.array_equal(c.get(_abi.LCP), ds["lcp"]), name
            assert c.max_lcp() == ds["max_lcp"], name
            if route is not None:
                assert c.sa_stats()["lcp_route"] == route, (name, c.sa_stats())
            z, mn, mx = c.factorize(3)
            f = c.factors(z)
            got = np.stack([f["pos"], f["src"], f["len"]], 1) if z else np.zeros((0, 3), np.uint32)
            assert np.array_equal(got, oracle.factorize(ds, t.size, 3)), name


def test_sim_key_layouts(simlib, oracle):
    """Initial-key layouts: power-of-two alphabets (length field) and others (sentinel code 0), texts ending in runs of
    the smallest symbol (the case the length field exists for), and forced short keys (many doubling rounds)."""
    from tudocomp_b200 import synth
    rng = np.random.default_rng(77)
    cases = []
    for sigma in (1, 2, 3, 4, 8, 100, 255):
        body = rng.integers(1, sigma + 1, 1500, dtype=np.uint16).astype(np.uint8)
        cases.append((f"sigma{sigma}", synth.with_sentinel(body)))
        tail = np.concatenate([body[:500], np.full(40, 1, np.uint8)])  # ...AAAA$ : padded keys would tie
        cases.append((f"sigma{sigma}_tailrun", synth.with_sentinel(tail)))
    try:
        for ksym in (None, "1"):
            if ksym is None:
                os.environ.pop("TDCGPU_SA_SYMBOLS", None)
            else:
                os.environ["TDCGPU_SA_SYMBOLS"] = ksym
            for name, t in cases:
                _check(simlib, oracle, f"{name}/k={ksym}", t, thresholds=(2,))
                with _abi.Context(simlib) as c:  # LCP alone: seeded from the keys + direct comparison (or Phi route)
                    c.set_text(t)
                    c.build(_abi.SA | _abi.LCP)
                    ds = oracle.textds(t)
                    assert np.array_equal(c.get(_abi.LCP), ds["lcp"]), (name, ksym, c.sa_stats())
                    assert c.max_lcp() == ds["max_lcp"], (name, ksym)
    finally:
        os.environ.pop("TDCGPU_SA_SYMBOLS", None)


def test_sim_set_text_cached(simlib, oracle):
    """tdcgpu_set_text_cached (the per-provider classes' entry point): structures survive exactly when the resident text has
    the same bytes; a different text of the SAME length must invalidate them."""
    from tudocomp_b200 import synth
    t1 = synth.markov_text(3000, 11)
    t2 = t1.copy()
    t2[1500] = t2[1500] % 26 + 97 if t2[1500] != 97 else 98  # one byte differs, same length
    assert t2[1500] != t1[1500]
    t3 = synth.dna(777, 12)
    with _abi.Context(simlib) as c:
        assert c.set_text_cached(t1) is False
        c.build(_abi.SA | _abi.LCP)
        assert np.array_equal(c.get(_abi.SA), oracle.textds(t1)["sa"])
        assert c.set_text_cached(t1.copy()) is True  # same bytes at another address: nothing rebuilt
        assert c.device_ptr(_abi.SA) is not None
        c.build(_abi.SA | _abi.ISA)
        assert np.array_equal(c.get(_abi.ISA), oracle.textds(t1)["isa"])
        assert c.set_text_cached(t2) is False  # same length, one byte differs
        assert c.device_ptr(_abi.SA) is None
        c.build(_abi.SA | _abi.ISA | _abi.LCP)
        ds2 = oracle.textds(t2)
        assert np.array_equal(c.get(_abi.SA), ds2["sa"]) and np.array_equal(c.get(_abi.LCP), ds2["lcp"])
        assert c.set_text_cached(t3) is False  # another length
        c.build(_abi.SA)
        assert np.array_equal(c.get(_abi.SA), oracle.textds(t3)["sa"])
        for tail in (15, 16, 17):  # the comparison's 16-byte vector part and its tail
            a = synth.markov_text(160 + tail, 13)
            b = a.copy()
            b[-2] = 98 if b[-2] != 98 else 99
            assert c.set_text_cached(a) is False and c.set_text_cached(a.copy()) is True and c.set_text_cached(b) is False


def _pack_reference(a, w):
    """DynamicIntVector layout (ds/BitPackingVector.hpp:62-98): element i at bits [i*w, (i+1)*w), 64-bit LE words."""
    m = np.uint64((1 << w) - 1)
    bits = (((a.astype(np.uint64) & m)[:, None] >> np.arange(w, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8).ravel()
    bits = np.concatenate([bits, np.zeros((-bits.size) % 64, np.uint8)])
    return np.packbits(bits, bitorder="little").view(np.uint64)


def test_sim_bit_packed_arrays(simlib):
    """tdcgpu_textds_get_packed: every array at several widths, incl. truncating ones (the stale PLCP[n-1])."""
    from tudocomp_b200 import synth
    for name, t in (("markov", synth.markov_text(5000, 1)), ("run", synth.with_sentinel(np.full(777, 97, np.uint8))),
                    ("sentinel_only", np.zeros(1, np.uint8))):
        with _abi.Context(simlib) as c:
            c.set_text(t)
            c.build(_abi.SA | _abi.ISA | _abi.LCP | _abi.PLCP | _abi.PHI)
            for which in (_abi.SA, _abi.ISA, _abi.LCP, _abi.PLCP, _abi.PHI):
                a = c.get(which)
                for w in (1, 7, 13, 17, 31, 32, 33, 40, 64, int(t.size).bit_length()):  # > 32: widened (wide-index builds of the caller)
                    assert np.array_equal(c.get_packed(which, w), _pack_reference(a, w)), (name, which, w)
            with pytest.raises(_abi.TdcGpuError):
                c.get_packed(_abi.SA, 65)
