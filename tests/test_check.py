"""The device-side checkers (tudocomp_b200/csrc/check.cu) must accept correct results and flag every kind of damage:
CPU — kernels in the interpreter (device pointers are host pointers there, so arrays can be damaged in place);
GPU — the same at sizes with millions of slots.  Also the ADVICE round-1 fixes that are visible at the C ABI:
sentinel position, get_factors before factorize, literal re-encode."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tudocomp_b200 import _abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")


@pytest.fixture(scope="module")
def simlib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    return _abi.TdcGpuLib(SIM)


def _view(ptr, n, dtype=np.uint32):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32 if dtype == np.uint32 else C.c_uint8)), shape=(n,))


CASES = [("banana", synth.with_sentinel(np.frombuffer(b"banana", np.uint8))),
         ("dna_5k", synth.dna(5000, 7)), ("markov_6k", synth.markov_text(6000, 8)),
         ("repetitive_8k", synth.repetitive(8000, 9, block=500, p=0.02)),
         ("run_a_3000", synth.with_sentinel(np.full(3000, 97, np.uint8))),
         ("fib18", synth.escape_with_sentinel(synth.fib_word(18)))]


@pytest.mark.sim
def test_sim_checkers_accept_correct_results_and_flag_damage(simlib):
    for name, t in CASES:
        for thr in (1, 3):
            with _abi.Context(simlib) as c:
                c.set_text(t)
                c.build(_abi.SA | _abi.ISA | _abi.LCP)
                z, _, _ = c.factorize(thr)
                res = c.check(thr, z)
                assert res["ok"], (name, thr, res)
                if t.size < 100:
                    continue
                n = t.size
                sa, isa, lcp = (_view(c.device_ptr(w), n) for w in (_abi.SA, _abi.ISA, _abi.LCP))
                # swap two neighbouring suffixes (and keep ISA consistent): only the ORDER criterion can notice
                i = n // 2
                sa[i], sa[i + 1] = sa[i + 1], sa[i]
                isa[sa[i]], isa[sa[i + 1]] = i, i + 1
                assert c.check()["suffix_order"] > 0, name
                sa[i], sa[i + 1] = sa[i + 1], sa[i]
                isa[sa[i]], isa[sa[i + 1]] = i, i + 1
                isa[5] ^= 1
                assert c.check()["isa_of_sa"] > 0, name
                isa[5] ^= 1
                lcp[n // 3] += 1
                assert c.check()["lcp"] > 0, name
                lcp[n // 3] -= 1
                assert c.check(thr, z)["ok"], name
                if z:
                    f = np.ctypeslib.as_array(C.cast(simlib.lib.tdcgpu_factors_device_ptr(c._h), C.POINTER(C.c_uint32)), shape=(z, 3))
                    k = z // 2
                    f[k, 2] -= 1  # a shorter factor: rule violation at its start (or malformed when it falls below the threshold)
                    r = c.check(thr, z)
                    assert r["factor_rule"] + r["malformed"] > 0, (name, r)
                    f[k, 2] += 1
                    if f[k, 1] > 0:
                        f[k, 1] -= 1  # another source
                        assert c.check(thr, z)["factor_rule"] > 0, name
                        f[k, 1] += 1
                    # drop one factor: its positions become literals that admit a factor
                    keep = np.delete(f.copy(), k, axis=0)
                    f[: z - 1] = keep
                    assert c.check(thr, z - 1)["missed_factor"] > 0, name


@pytest.mark.sim
def test_sim_abi_state_and_sentinel_errors(simlib):
    with _abi.Context(simlib) as c:
        bad = np.frombuffer(b"AB\x00CD", np.uint8).copy()          # exactly one 0, but not at the end
        with pytest.raises(_abi.TdcGpuError) as e:
            c.set_text(bad)
        assert e.value.code == -3
        two = np.frombuffer(b"AB\x00CD\x00", np.uint8).copy()      # ends in 0 but holds another one: the builder rejects it
        c.set_text(two)
        with pytest.raises(_abi.TdcGpuError) as e:
            c.build(_abi.SA)
        assert e.value.code == -3
        t = synth.dna(3000, 5)
        c.set_text(t)
        with pytest.raises(_abi.TdcGpuError) as e:                 # no factor list yet
            c.factors(0)
        assert e.value.code == -6
        # re-encoding the same staged literal input with another table must work (scratch is rewound)
        data = np.random.default_rng(1).integers(0, 256, 300000, dtype=np.uint8)
        c.literal_histogram_of(data)
        codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)
        a = c.literal_encode(codes, lens)
        b = c.literal_encode(codes, lens)
        assert np.array_equal(a, b)
        lens2 = np.full(256, 9, np.uint8)
        assert c.literal_encode(codes, lens2).size > a.size


@pytest.mark.gpu
def test_gpu_checkers_on_millions_of_slots():
    import tudocomp_b200 as tdc
    lib = tdc.load()
    for name, t in (("dna_8m", synth.dna(1 << 23, 3)), ("markov_4m", synth.markov_text(1 << 22, 4)),
                    ("repetitive_4m", synth.repetitive(1 << 22, 5, block=1 << 14, p=0.01))):
        with tdc.Context(lib, 0) as c:
            c.set_text(t)
            c.build(tdc.SA | tdc.ISA | tdc.LCP)
            z, _, _ = c.factorize(3)
            res = c.check(3, z)
            assert res["ok"], (name, res)
            # damage on the device: overwrite one LCP value / one SA pair through the C ABI's own device pointers
            import torch
            lcp = c.get(tdc.LCP)
            lcp[12345] += 1
            dl = torch.from_numpy(lcp.view(np.int32)).cuda()
            sa_p, isa_p = c.device_ptr(tdc.SA), c.device_ptr(tdc.ISA)
            r = c.check_index_ptrs(lib.lib.tdcgpu_text_device_ptr(c._h), t.size, sa_p, isa_p, dl.data_ptr(), 0, t.size)
            assert r["lcp"] == 1 and r["suffix_order"] == 0 and r["isa_of_sa"] == 0, (name, r)
