"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI (ctypes over tudocomp_b200/libtdcgpu.so),
against the oracle on seeded inputs, against the committed golden vectors of the unmodified reference, and — at sizes
the CPU oracle cannot reach in seconds — through size-independent properties (Burkhardt–Kärkkäinen SA check, ISA
inverse, Kasai-free LCP spot checks, factor decode round trip, greedy-parse spot checks).
Bit-exact everywhere: all results are integers/bytes."""
import os

import numpy as np
import pytest

from inputs import all_small_cases
import tudocomp_b200 as tdc
from tudocomp_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")
ALL = tdc.SA | tdc.ISA | tdc.LCP | tdc.PHI | tdc.PLCP | tdc.BWT


@pytest.fixture(scope="module")
def ctx():
    c = tdc.Context(tdc.load(), 0)
    yield c
    c.close()


def _triples(f):
    return np.stack([f["pos"], f["src"], f["len"]], 1) if f.size else np.zeros((0, 3), np.uint32)


def test_golden_vectors_of_the_reference(ctx):
    gold = np.load(GOLD)
    for name in gold["names"]:
        t = gold[f"{name}/text"]
        ctx.set_text(t)
        ctx.build(ALL)
        for k, fl in (("sa", tdc.SA), ("isa", tdc.ISA), ("lcp", tdc.LCP), ("phi", tdc.PHI), ("plcp", tdc.PLCP), ("bwt", tdc.BWT)):
            assert np.array_equal(ctx.get(fl), gold[f"{name}/{k}"]), (name, k)
        assert ctx.max_lcp() == int(gold[f"{name}/max_lcp"][0]), name
        for thr in (1, 2, 3, 5):
            z, mn, mx = ctx.factorize(thr)
            assert np.array_equal(_triples(ctx.factors(z)), gold[f"{name}/factors{thr}"]), (name, thr)
            hdr = gold[f"{name}/hdr{thr}"]
            assert (mn, mx) == (int(hdr[0]) & 0xFFFFFFFF, int(hdr[1])), (name, thr)


def test_oracle_parity_on_seeded_inputs(ctx, oracle):
    cases = list(all_small_cases()) + [
        ("dna_300k", synth.dna(300000, 21)), ("markov_300k", synth.markov_text(300000, 22)),
        ("repetitive_300k", synth.repetitive(300000, 23, block=5000, p=0.01)),
        ("run_a_100k", synth.with_sentinel(np.full(100000, 97, np.uint8))),
        ("fib25", synth.escape_with_sentinel(synth.fib_word(25))),
        ("two_symbols", synth.with_sentinel(np.random.default_rng(5).integers(1, 3, 200000, dtype=np.uint8))),
        ("full_alphabet", synth.with_sentinel(np.random.default_rng(6).integers(1, 255, 200000, dtype=np.uint8))),
    ]
    for name, t in cases:
        ds = oracle.textds(t)
        ctx.set_text(t)
        ctx.build(ALL)
        for k, fl in (("sa", tdc.SA), ("isa", tdc.ISA), ("lcp", tdc.LCP), ("phi", tdc.PHI), ("plcp", tdc.PLCP)):
            assert np.array_equal(ctx.get(fl), ds[k]), (name, k)
        assert ctx.max_lcp() == ds["max_lcp"], name
        assert np.array_equal(ctx.get(tdc.BWT), oracle.bwt(t, ds["sa"])), name
        for thr in (2, 3, 8):
            z, mn, mx = ctx.factorize(thr)
            want = oracle.factorize(ds, t.size, thr)
            assert np.array_equal(_triples(ctx.factors(z)), want), (name, thr)
            wmn, wmx, _ = oracle.factor_stats(want, t.size)
            assert (mn, mx) == (wmn, wmx), (name, thr)


def test_reference_parity_when_ref_library_travelled(ctx, reference):
    for name, t in (("dna_1m", synth.dna(1 << 20, 31)), ("markov_1m", synth.markov_text(1 << 20, 32))):
        ds = reference.textds(t)
        ctx.set_text(t)
        ctx.build(ALL)
        for k, fl in (("sa", tdc.SA), ("isa", tdc.ISA), ("lcp", tdc.LCP), ("phi", tdc.PHI), ("plcp", tdc.PLCP)):
            assert np.array_equal(ctx.get(fl), ds[k]), (name, k)
        f_ref, hdr = reference.factors(t, 3)
        z, mn, mx = ctx.factorize(3)
        assert np.array_equal(_triples(ctx.factors(z)), f_ref), name
        assert (mn, mx) == (hdr[0], hdr[1])


def _check_properties(t, ctx, rng, thr=3, samples=2000):
    n = t.size
    ctx.set_text(t)
    ctx.build(tdc.SA | tdc.ISA | tdc.LCP | tdc.BWT)
    sa, isa, lcp, bwt = ctx.get(tdc.SA), ctx.get(tdc.ISA), ctx.get(tdc.LCP), ctx.get(tdc.BWT)
    # permutation + inverse (ds_tests.cpp:71-99)
    assert sa[0] == n - 1
    assert np.array_equal(isa[sa], np.arange(n, dtype=np.uint32))
    # Burkhardt–Kärkkäinen: (T[SA[i-1]], ISA[SA[i-1]+1]) < (T[SA[i]], ISA[SA[i]+1]) for all i >= 1
    a, b = sa[:-1].astype(np.int64), sa[1:].astype(np.int64)
    ra = np.where(a + 1 < n, isa[np.minimum(a + 1, n - 1)], 0).astype(np.int64)
    rb = np.where(b + 1 < n, isa[np.minimum(b + 1, n - 1)], 0).astype(np.int64)
    ca, cb = t[a].astype(np.int64), t[b].astype(np.int64)
    assert np.all((ca < cb) | ((ca == cb) & (ra < rb)))
    # BWT definition (bwt.hpp:19-22)
    assert np.array_equal(bwt, np.where(sa == 0, t[n - 1], t[(sa.astype(np.int64) - 1) % n]))
    # LCP spot checks by direct comparison (ds_tests.cpp:101-112)
    assert lcp[0] == 0
    tb = t.tobytes()
    for i in rng.integers(1, n, size=samples):
        x, y, l = int(sa[i - 1]), int(sa[i]), int(lcp[i])
        assert tb[x:x + l] == tb[y:y + l] and tb[x + l] != tb[y + l]
    # factors: decode round trip (lzss::decode_text semantics) + greedy/threshold invariants
    z, mn, mx = ctx.factorize(thr)
    f = ctx.factors(z)
    pos, src, ln = f["pos"].astype(np.int64), f["src"].astype(np.int64), f["len"].astype(np.int64)
    assert np.all(ln >= thr) and np.all(src < pos) and np.all(pos[1:] >= pos[:-1] + ln[:-1]) and pos[-1] + ln[-1] <= n - 1
    assert (mn, mx) == (ln.min(), ln.max())
    for k in rng.integers(0, z, size=samples):
        p, s, l = int(pos[k]), int(src[k]), int(ln[k])
        assert tb[p:p + l] == tb[s:s + l] and tb[p + l] != tb[s + l]  # maximal copy from an earlier suffix
    out = np.array(t)
    covered = np.zeros(n + 1, np.int64)
    np.add.at(covered, pos, 1)
    np.add.at(covered, pos + ln, -1)
    lit = np.cumsum(covered[:-1]) == 0
    out[~lit] = 0
    # non-overlapping factors can be replayed vectorised in rounds; overlapping ones (src+len > pos) byte-wise
    for k in range(z):
        p, s, l = int(pos[k]), int(src[k]), int(ln[k])
        if s + l <= p:
            out[p:p + l] = out[s:s + l]
        else:
            for j in range(l):
                out[p + j] = out[s + j]
    assert np.array_equal(out, t)
    return z


def test_properties_at_sizes_beyond_the_cpu_oracle(ctx):
    rng = np.random.default_rng(7)
    for name, t in (("dna_16m", synth.dna(1 << 24, 41)), ("markov_8m", synth.markov_text(1 << 23, 42)),
                    ("repetitive_8m", synth.repetitive(1 << 23, 43, block=1 << 16, p=0.01))):
        z = _check_properties(t, ctx, rng)
        assert z > 0, name


def test_error_behaviour(ctx):
    lib = tdc.load()
    t = np.frombuffer(b"ban\0ana\0", np.uint8).copy()  # a second 0 violates the sentinel contract
    ctx.set_text(t)
    with pytest.raises(tdc.TdcGpuError) as e:
        ctx.build(tdc.SA)
    assert e.value.code == -3
    ctx.set_text(np.frombuffer(b"banana\0", np.uint8).copy())
    with pytest.raises(tdc.TdcGpuError):
        ctx.factorize(0)
    with pytest.raises(tdc.TdcGpuError):
        ctx.get(tdc.LCP)  # not built yet
    assert ctx.factorize(3)[0] == 1


def test_one_shot_host_entry_points():
    import ctypes as C
    lib = tdc.load()
    t = synth.markov_text(100000, 9)
    n = t.size
    sa, isa, lcp = (np.zeros(n, np.uint32) for _ in range(3))
    mx = C.c_uint32()
    rc = lib.lib.tdcgpu_textds_build_host(0, t.ctypes.data, n, sa.ctypes.data, isa.ctypes.data, lcp.ctypes.data, None, None, C.byref(mx))
    assert rc == 0, lib.lib.tdcgpu_last_error()
    assert np.array_equal(isa[sa], np.arange(n, dtype=np.uint32)) and int(lcp.max()) == mx.value
    out = np.zeros(n, np.uint8)
    assert lib.lib.tdcgpu_bwt_host(0, t.ctypes.data, n, out.ctypes.data) == 0
    assert np.array_equal(out, np.where(sa == 0, t[n - 1], t[(sa.astype(np.int64) - 1) % n]))
    assert np.array_equal(tdc.bwt(t), out)


def test_bit_packed_arrays(ctx):
    """tdcgpu_textds_get_packed against the DynamicIntVector layout computed with numpy (ds/BitPackingVector.hpp:62-98)."""
    def pack(a, w):
        m = np.uint64((1 << w) - 1)
        bits = (((a.astype(np.uint64) & m)[:, None] >> np.arange(w, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8).ravel()
        bits = np.concatenate([bits, np.zeros((-bits.size) % 64, np.uint8)])
        return np.packbits(bits, bitorder="little").view(np.uint64)

    t = synth.markov_text(1 << 20, 61)
    ctx.set_text(t)
    ctx.build(ALL)
    for fl in (tdc.SA, tdc.ISA, tdc.LCP, tdc.PLCP, tdc.PHI):
        a = ctx.get(fl)
        for w in (5, 21, int(t.size).bit_length(), 32):
            assert np.array_equal(ctx.get_packed(fl, w), pack(a, w)), (fl, w)
